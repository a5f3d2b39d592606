// Balloon-Windkessel BOLD + streaming zero-phase band-pass + decimation.
//
// Replaces BOLDModel.Sim (call site netwWilsonCowanPlastic.py:144) and the cut / filtfilt /
// [::BOLD_downsamp] of simBOLD (netwWilsonCowanPlastic.py:145-156).
//
// filtfilt needs the whole forward output reversed; only every ds-th output is kept, so the
// backward pass is evaluated from per-chunk summaries instead:
//   H(z) = b0 + sum_i rho_i z^-1 / (1 - p_i z^-1)       (parallel form, 2 conjugate pole pairs)
//   forward : s_i <- p_i s_i + u,  w = b0 u + 2 Re(sum rho_i s_i)        (state before u)
//   backward: Z_i(m) = p_i Z_i(m+1) + w[m],  y[m] = b0 w[m] + 2 Re(sum rho_i Z_i(m+1))
//   chunk   : Z(m0) = p^L Z(m0+L) + sum_k p^k w[m0+k]                    (summary c_j)
// with SciPy's defaults reproduced exactly: odd extension by 15 samples on both sides and
// steady-state (lfilter_zi) start-up of both passes.  All filter arithmetic is float64.
#pragma once
#include "common.cuh"

namespace nrem {

constexpr int kPad = 15;

struct FiltCoef {
    double b0;
    double Pr[2], Pi[2];     // one pole of each conjugate pair
    double Rr[2], Ri[2];     // its residue
    double Qr[2], Qi[2];     // 1 / (1 - p)
    double PLr[2], PLi[2];   // p^ds
    double PEr[2], PEi[2];   // p^(length of the last chunk)
    int64_t Tf;              // samples after the cut
    int64_t ds;              // decimation
    int64_t J;               // ceil(Tf / ds) outputs
    int64_t M;               // Tf + 2*kPad forward outputs
    const double* ptab;      // [ds + 2*kPad][4]  (Re p0^k, Im p0^k, Re p1^k, Im p1^k)
};

// Per-thread persistent storage, structure-of-arrays with stride nth (one slot per (sim,node)).
struct FiltScratch {
    double* fs;     // [4][nth]   forward state
    double* cs;     // [4][nth]   running chunk summary
    double* head;   // [16][nth]  first 16 samples after the cut
    double* tail;   // [16][nth]  last 16 samples
    double* summ;   // [J][4][nth]
    double* wdec;   // [J][nth]
    double* wlast;  // [nth]
    int64_t nth;
};

struct FiltRun {
    double s[4], c[4];
    int k, j;           // position inside the current chunk / chunk index (valid when m >= 16); all fit in 32 bits
    int kd, jd;         // (m - 15) % ds and (m - 15) / ds
};

__device__ __forceinline__ void filt_counters(FiltRun& r, const FiltCoef& f, int64_t m) {
    // m = index of the NEXT forward output (ext coordinates)
    if (m >= 16) {
        int64_t j = (m - 16) / f.ds;
        if (j > f.J - 1) j = f.J - 1;
        r.j = (int)j;
        r.k = (int)(m - 16 - j * f.ds);
    } else {
        r.j = 0; r.k = (int)(m - 16);    // negative until m reaches 16
    }
    if (m >= 15) { r.jd = (int)((m - 15) / f.ds); r.kd = (int)((m - 15) % f.ds); }
    else { r.jd = 0; r.kd = (int)(m - 15); }
}

__device__ __forceinline__ void filt_push(FiltRun& r, const FiltCoef& f, const FiltScratch& S, int64_t slot, double u, int64_t m) {
    const double w = f.b0 * u + 2.0 * ((f.Rr[0] * r.s[0] - f.Ri[0] * r.s[1]) + (f.Rr[1] * r.s[2] - f.Ri[1] * r.s[3]));
    {
        const double a0 = f.Pr[0] * r.s[0] - f.Pi[0] * r.s[1] + u;
        const double a1 = f.Pr[0] * r.s[1] + f.Pi[0] * r.s[0];
        const double a2 = f.Pr[1] * r.s[2] - f.Pi[1] * r.s[3] + u;
        const double a3 = f.Pr[1] * r.s[3] + f.Pi[1] * r.s[2];
        r.s[0] = a0; r.s[1] = a1; r.s[2] = a2; r.s[3] = a3;
    }
    const int ds = (int)f.ds, Jm1 = (int)f.J - 1;
    if (r.kd == 0 && r.jd <= Jm1) S.wdec[(int64_t)r.jd * S.nth + slot] = w;
    if (r.k >= 0) {
        if (r.k == 0) { r.c[0] = r.c[1] = r.c[2] = r.c[3] = 0.0; }
        const double2* pt = reinterpret_cast<const double2*>(f.ptab) + 2 * r.k;     // the table is 32-byte aligned per row
        const double2 p01 = __ldg(pt), p23 = __ldg(pt + 1);
        r.c[0] = fma(p01.x, w, r.c[0]);
        r.c[1] = fma(p01.y, w, r.c[1]);
        r.c[2] = fma(p23.x, w, r.c[2]);
        r.c[3] = fma(p23.y, w, r.c[3]);
        const bool last_of_chunk = (r.j < Jm1) ? (r.k == ds - 1) : (m == f.M - 1);
        if (last_of_chunk) {
#pragma unroll
            for (int q = 0; q < 4; ++q) S.summ[((int64_t)r.j * 4 + q) * S.nth + slot] = r.c[q];
        }
    }
    if (m == f.M - 1) S.wlast[slot] = w;
    // advance the uniform counters
    ++r.k;
    if (r.k == ds && r.j < Jm1) { r.k = 0; ++r.j; }
    ++r.kd;
    if (r.kd == ds) { r.kd = 0; ++r.jd; }
}

// filt_push for samples strictly inside the signal (16 <= n < Tf - 16): no head/tail/end-of-signal checks.
__device__ __forceinline__ void filt_push_steady(FiltRun& r, const FiltCoef& f, const FiltScratch& S, int64_t slot, double u, int ds, int Jm1) {
    const double w = f.b0 * u + 2.0 * ((f.Rr[0] * r.s[0] - f.Ri[0] * r.s[1]) + (f.Rr[1] * r.s[2] - f.Ri[1] * r.s[3]));
    {
        const double a0 = f.Pr[0] * r.s[0] - f.Pi[0] * r.s[1] + u;
        const double a1 = f.Pr[0] * r.s[1] + f.Pi[0] * r.s[0];
        const double a2 = f.Pr[1] * r.s[2] - f.Pi[1] * r.s[3] + u;
        const double a3 = f.Pr[1] * r.s[3] + f.Pi[1] * r.s[2];
        r.s[0] = a0; r.s[1] = a1; r.s[2] = a2; r.s[3] = a3;
    }
    if (r.kd == 0) S.wdec[(int64_t)r.jd * S.nth + slot] = w;
    if (r.k == 0) { r.c[0] = r.c[1] = r.c[2] = r.c[3] = 0.0; }
    const double2* pt = reinterpret_cast<const double2*>(f.ptab) + 2 * r.k;
    const double2 p01 = __ldg(pt), p23 = __ldg(pt + 1);
    r.c[0] = fma(p01.x, w, r.c[0]);
    r.c[1] = fma(p01.y, w, r.c[1]);
    r.c[2] = fma(p23.x, w, r.c[2]);
    r.c[3] = fma(p23.y, w, r.c[3]);
    ++r.k;
    if (r.k == ds && r.j < Jm1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) S.summ[((int64_t)r.j * 4 + q) * S.nth + slot] = r.c[q];
        r.k = 0; ++r.j;
    }
    if (++r.kd == ds) { r.kd = 0; ++r.jd; }
}

// Feeds sample n (index after the cut) with value x; handles both odd extensions.
__device__ __forceinline__ void filt_feed(FiltRun& r, const FiltCoef& f, const FiltScratch& S, int64_t slot, double x, int64_t n) {
    if (n < 16) {
        S.head[n * S.nth + slot] = x;
        if (n == 15) {
            const double x0 = S.head[slot];
            const double e0 = 2.0 * x0 - x;                // ext[0] = 2 x0 - x15
#pragma unroll
            for (int q = 0; q < 2; ++q) {                   // steady state for a constant input e0 (lfilter_zi * ext[0])
                r.s[2 * q] = f.Qr[q] * e0;
                r.s[2 * q + 1] = f.Qi[q] * e0;
            }
            filt_counters(r, f, 0);
            for (int e = 0; e < kPad; ++e) filt_push(r, f, S, slot, 2.0 * x0 - S.head[(15 - e) * S.nth + slot], e);
            for (int e = 0; e < 16; ++e) filt_push(r, f, S, slot, S.head[e * S.nth + slot], kPad + e);
        }
    } else {
        filt_push(r, f, S, slot, x, n + kPad);
    }
    if (n >= f.Tf - 16) S.tail[(n - (f.Tf - 16)) * S.nth + slot] = x;
    if (n == f.Tf - 1) {
        for (int e = 0; e < kPad; ++e) filt_push(r, f, S, slot, 2.0 * x - S.tail[(14 - e) * S.nth + slot], f.Tf + kPad + e);
    }
}

__device__ __forceinline__ void filt_load(FiltRun& r, const FiltCoef& f, const FiltScratch& S, int64_t slot, int64_t n_next) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { r.s[q] = S.fs[q * S.nth + slot]; r.c[q] = S.cs[q * S.nth + slot]; }
    filt_counters(r, f, n_next >= 16 ? n_next + kPad : 0);
}
__device__ __forceinline__ void filt_store(const FiltRun& r, const FiltScratch& S, int64_t slot) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { S.fs[q * S.nth + slot] = r.s[q]; S.cs[q * S.nth + slot] = r.c[q]; }
}

// ---- Balloon-Windkessel ------------------------------------------------------------------------
// State (s, f, v, q) from (0.1, 1, 1, 1); constants of SURVEY.md section 8c / oracle/bold_oracle.py.
template <typename T> struct BW;
template <> struct BW<double> {
    double s, f, v, q;
    __device__ __forceinline__ void init() { s = 0.1; f = 1.0; v = 1.0; q = 1.0; }
    __device__ __forceinline__ double step(double x, double dt) {
        const double kappa = 1.0 / 0.65, gamma = 1.0 / 0.41, tau = 0.98, E0 = 0.4, V0 = 0.04, TE = 0.04;
        const double k1 = 4.3 * 40.3 * E0 * TE, k2 = 25.0 * E0 * TE, k3 = 1.0, ia = 1.0 / 0.32;
        const double out = V0 * (k1 * (1 - q) + k2 * (1 - q / v) + k3 * (1 - v));
        const double va = pow(v, ia);
        const double ds = x - kappa * s - gamma * (f - 1);
        const double df = s;
        const double dv = (f - va) / tau;
        const double dq = (f * (1 - pow(1 - E0, 1 / f)) / E0 - q * va / v) / tau;
        s += dt * ds; f += dt * df; v += dt * dv; q += dt * dq;
        return out;
    }
};
template <> struct BW<float> {
    float s, f, v, q;
    __device__ __forceinline__ void init() { s = 0.1f; f = 1.0f; v = 1.0f; q = 1.0f; }
    __device__ __forceinline__ double step(float x, float dt) {
        const float kappa = 1.0f / 0.65f, gamma = 1.0f / 0.41f, itau = 1.0f / 0.98f, E0 = 0.4f, V0 = 0.04f, TE = 0.04f;
        const float k1 = 4.3f * 40.3f * E0 * TE, k2 = 25.0f * E0 * TE, k3 = 1.0f, ia = 1.0f / 0.32f;
        const float lg06 = -0.7369655941662062f;          // log2(1 - E0)
        const float iv = rcpf(v);
        // (1-q), (1-v) are exact float32 differences; the sum is formed in float32 and widened once for the float64 filter
        const double out = (double)(V0 * fmaf(k1, 1.0f - q, fmaf(k2, 1.0f - q * iv, k3 * (1.0f - v))));
        const float va = ex2f(ia * lg2f(v));
        const float ds = x - kappa * s - gamma * (f - 1.0f);
        const float df = s;
        const float dv = (f - va) * itau;
        const float dq = (f * (1.0f - ex2f(lg06 * rcpf(f))) * (1.0f / E0) - q * va * iv) * itau;
        s += dt * ds; f += dt * df; v += dt * dv; q += dt * dq;
        return out;
    }
};

// ---- kernels -----------------------------------------------------------------------------------

// Stage API: bold_sim on [B,T,N] float64.  Thread = (b, node), node fastest.
__global__ void bold_sim_f64_kernel(const double* rE, int B, int64_t T, int N, double dt, double* out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)B * N) return;
    const int64_t b = tid / N, n = tid % N;
    const double* in = rE + b * T * N + n;
    double* o = out + b * T * N + n;
    BW<double> bw;
    bw.init();
    for (int64_t t = 0; t < T; ++t) o[t * N] = bw.step(in[t * N], dt);
}

// Stage API: forward pass of the zero-phase filter over bold [B,T,N]; slot = b*N + node.
__global__ void filt_forward_f64_kernel(const double* bold, int B, int64_t T, int N, int64_t Neq, FiltCoef f, FiltScratch S) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (int64_t)B * N) return;
    const int64_t b = slot / N, n = slot % N;
    const double* in = bold + b * T * N + n;
    FiltRun r;
    r.s[0] = r.s[1] = r.s[2] = r.s[3] = 0.0;
    r.c[0] = r.c[1] = r.c[2] = r.c[3] = 0.0;
    filt_counters(r, f, 0);
    for (int64_t t = Neq; t < T; ++t) filt_feed(r, f, S, slot, in[t * N], t - Neq);
}

// Backward recursion over the chunk summaries.  out index = b*out_sb + j*out_sj + n*out_sn,
// slot -> (b, n) by slot_b_fast: slot = n*Bs + b (sweep layout) or b*N + n (stage layout).
__global__ void filt_backward_kernel(FiltCoef f, FiltScratch S, int64_t nslots, int N, int64_t Bs, int slot_b_fast,
                                     double* out, int64_t out_sb, int64_t out_sj, int64_t out_sn, int64_t Bvalid) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    int64_t b, n;
    if (slot_b_fast) { n = slot / Bs; b = slot % Bs; } else { b = slot / N; n = slot % N; }
    if (b >= Bvalid) return;
    const double wl = S.wlast[slot];
    double Z[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) { Z[2 * q] = f.Qr[q] * wl; Z[2 * q + 1] = f.Qi[q] * wl; }
    for (int64_t j = f.J - 1; j >= 0; --j) {
        const bool last = (j == f.J - 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double pr = last ? f.PEr[q] : f.PLr[q], pi = last ? f.PEi[q] : f.PLi[q];
            const double zr = pr * Z[2 * q] - pi * Z[2 * q + 1] + S.summ[(j * 4 + 2 * q) * S.nth + slot];
            const double zi = pr * Z[2 * q + 1] + pi * Z[2 * q] + S.summ[(j * 4 + 2 * q + 1) * S.nth + slot];
            Z[2 * q] = zr; Z[2 * q + 1] = zi;
        }
        const double y = f.b0 * S.wdec[j * S.nth + slot] + 2.0 * ((f.Rr[0] * Z[0] - f.Ri[0] * Z[1]) + (f.Rr[1] * Z[2] - f.Ri[1] * Z[3]));
        out[b * out_sb + j * out_sj + n * out_sn] = y;
    }
}

// Sweep path: consume `rows` float32 E samples [rows][N][Bs] (simulation fastest) of the simulations
// [sim0, sim0 + nsim), advance the Balloon-Windkessel state and the forward filter.  slot = node*Bs + sim.
// Occupancy: the per-sample recursions are one dependent chain per thread (~460 clk per sample), so the kernel lives on resident
// warps.  64 registers (8 CTAs per SM, a 4-deep load prefetch; 16 bytes of spill) instead of 90 (5 CTAs, 8-deep): +2.6 % on the whole
// sweep (profiles/r02_kernel_variants.md).
#ifndef NREM_K2_MINB
#define NREM_K2_MINB 8
#endif
#ifndef NREM_K2_PF
#define NREM_K2_PF 4
#endif
// STEADY (decided by the host, launch_bold_chunk): the whole chunk lies strictly inside the filtered signal (no cut, no odd extension,
// not the end), so the per-sample code has no position checks; WR: also write the series-major sample ring of the Welch kernel.
template <typename BT, bool STEADY, bool WR>
__global__ void __launch_bounds__(128, NREM_K2_MINB) bold_filter_chunk_kernel(const float* Ebuf, int rows, int64_t row_base, int N, int64_t Bs, int64_t sim0, int64_t nsim,
                                         int64_t Neq, BT dt, BT* bw_state /*[4][nth]*/, FiltCoef f, FiltScratch S, float* wring, int wL) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * nsim) return;
    const int64_t slot = (t / nsim) * Bs + sim0 + (t % nsim);
    BW<BT> bw;
    if (row_base == 0) {
        bw.init();
    } else {
        bw.s = bw_state[slot]; bw.f = bw_state[S.nth + slot]; bw.v = bw_state[2 * S.nth + slot]; bw.q = bw_state[3 * S.nth + slot];
    }
    FiltRun r;
    filt_load(r, f, S, slot, row_base - Neq);
    const float* in = Ebuf + slot;
    const int64_t stride = (int64_t)N * Bs;
    // optional series-major copy of the samples for the Welch kernel: wring[slot][sample mod wL], 16-byte stores
    float* wr = WR ? wring + slot * (int64_t)wL : nullptr;
    const int lead = (int)((4 - (row_base & 3)) & 3), full_end = lead + ((rows - lead) & ~3);
    int wpos = WR ? (int)(row_base % wL) : 0;                       // ring position of the current sample
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
    const int ds = (int)f.ds, Jm1 = (int)f.J - 1;
    // The recursions below are sequential in the sample index, the LOADS are not: without the explicit block prefetch every sample
    // pays one HBM/L2 round trip (250 samples x ~0.8 us made this kernel 5.6 % of a sweep).  Block b + 1 is requested while block b
    // is processed.
    constexpr int PF = NREM_K2_PF;
    float xnext[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) xnext[j] = j < rows ? __ldcs(in + (int64_t)j * stride) : 0.f;
    for (int r0 = 0; r0 < rows; r0 += PF) {
      float xcur[PF];
#pragma unroll
      for (int j = 0; j < PF; ++j) xcur[j] = xnext[j];
#pragma unroll
      for (int j = 0; j < PF; ++j) xnext[j] = r0 + PF + j < rows ? __ldcs(in + (int64_t)(r0 + PF + j) * stride) : 0.f;
      // Lean block: none of the PF samples writes a decimated output, starts or ends a summary chunk (the counters are the same in
      // every thread, so this is one uniform branch per block; with ds = 1000 all but one or two blocks of a launch are lean).  Same
      // arithmetic in the same order as filt_push_steady.
      if (STEADY && !WR && r0 + PF <= rows && r.kd >= 1 && r.kd + PF <= ds && r.k >= 1 && (r.k + PF < ds || r.j >= Jm1)) {
        const double2* pt = reinterpret_cast<const double2*>(f.ptab) + 2 * r.k;
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const double u = bw.step((BT)xcur[j], dt);
            const double w = f.b0 * u + 2.0 * ((f.Rr[0] * r.s[0] - f.Ri[0] * r.s[1]) + (f.Rr[1] * r.s[2] - f.Ri[1] * r.s[3]));
            const double a0 = f.Pr[0] * r.s[0] - f.Pi[0] * r.s[1] + u;
            const double a1 = f.Pr[0] * r.s[1] + f.Pi[0] * r.s[0];
            const double a2 = f.Pr[1] * r.s[2] - f.Pi[1] * r.s[3] + u;
            const double a3 = f.Pr[1] * r.s[3] + f.Pi[1] * r.s[2];
            r.s[0] = a0; r.s[1] = a1; r.s[2] = a2; r.s[3] = a3;
            const double2 p01 = __ldg(pt + 2 * j), p23 = __ldg(pt + 2 * j + 1);
            r.c[0] = fma(p01.x, w, r.c[0]);
            r.c[1] = fma(p01.y, w, r.c[1]);
            r.c[2] = fma(p23.x, w, r.c[2]);
            r.c[3] = fma(p23.y, w, r.c[3]);
        }
        r.k += PF;
        r.kd += PF;
        if (r.kd == ds) { r.kd = 0; ++r.jd; }
        continue;
      }
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int rr = r0 + j;
        if (rr >= rows) break;
        const float xe = xcur[j];
        if (WR) {
            if (rr < lead || rr >= full_end) wr[wpos] = xe;
            else {
                const int ph = (rr - lead) & 3;
                if (ph == 0) q0 = xe; else if (ph == 1) q1 = xe; else if (ph == 2) q2 = xe;
                else *reinterpret_cast<float4*>(wr + wpos - 3) = make_float4(q0, q1, q2, xe);
            }
            if (++wpos == wL) wpos = 0;
        }
        const double y = bw.step((BT)xe, dt);
        if (STEADY) {
            filt_push_steady(r, f, S, slot, y, ds, Jm1);
        } else {
            const int64_t ts = row_base + rr;
            if (ts >= Neq && ts - Neq < f.Tf) filt_feed(r, f, S, slot, y, ts - Neq);
        }
      }
    }
    bw_state[slot] = bw.s; bw_state[S.nth + slot] = bw.f; bw_state[2 * S.nth + slot] = bw.v; bw_state[3 * S.nth + slot] = bw.q;
    filt_store(r, S, slot);
}

}  // namespace nrem
