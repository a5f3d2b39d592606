// Welch power spectrum of the stored E samples and its peak frequency.
//
// Replaces  freqs, fftPow = signal.welch(E_t.T, fs=1/dt, nperseg=4000); meanpow = fftPow.mean(axis=0);
//           peakfreq = freqs[argmax(meanpow)]                                  (whole_sweep_both.py:90-95)
// with SciPy's defaults: periodic Hann window, 50 % overlap, constant detrend, one-sided density scaling, mean over
// segments.  One launch handles one segment (L samples) of every (simulation, node) series:
//   * a CTA owns 4 simulations and loops over the nodes; a series is packed as L/2 complex numbers
//     z[n] = x[2n] + i x[2n+1], transformed by a shared-memory Stockham autosort FFT (radices 4/2/5; L/2 = 2000 =
//     4*4*5*5*5 for the reference), and un-packed to the L/2+1 one-sided bins;
//   * |X|^2 is accumulated over nodes in shared memory and added once per launch to P[sim][L/2+1] (float32), so the
//     E samples are read exactly once more and nothing but the accumulator goes back to HBM.
// The samples come from a series-major ring  wring[node][sim][sample mod L]  that the BOLD/filter kernel fills while it
// consumes the integrator's row-major chunk buffer, so every series is one contiguous, coalesced read.
#pragma once
#include "common.cuh"

namespace nrem {

constexpr int kWelchSims = 4;
constexpr int kWelchThreads = 1024;        // 256 threads per series: the passes are latency-bound, occupancy is what pays
constexpr int kWelchMaxStages = 8;

struct WelchPlan {
    int L;                          // nperseg (even)
    int M;                          // L / 2 = product of the radices
    int nstages;
    int radix[kWelchMaxStages];
    const float* window;            // [L] periodic Hann
    const float2* tw;               // [M]  exp(-2 pi i k / M)
    const float2* tw2;              // [M + 1] exp(-2 pi i k / L) for the real-FFT un-packing
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// One Stockham autosort pass of radix R over M complex points: in -> out.  P = product of the previous radices
// (compile-time when PC > 0, so that i % p becomes a multiply-shift).  tw[k] = exp(-2 pi i k / M) in shared memory.
template <int R, int PC>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out, int M, int prt, const float2* tw,
                                              int lane, int nlanes) {
    const int p = PC > 0 ? PC : prt;
    const int t = M / R;
    const int twstep = M / (p * R);
    for (int i = lane; i < t; i += nlanes) {
        const int k = i % p;
        const int j = (i - k) * R + k;
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            u[r] = in[i + r * t];
            if (r > 0) u[r] = cmul(u[r], tw[r * k * twstep]);          // r*k*twstep < M: no wrap
        }
        if (R == 2) {
            out[j] = cadd(u[0], u[1]);
            out[j + p] = csub(u[0], u[1]);
        } else if (R == 4) {
            const float2 a = cadd(u[0], u[2]), b = csub(u[0], u[2]), c = cadd(u[1], u[3]), d = csub(u[1], u[3]);
            out[j] = cadd(a, c);
            out[j + p] = make_float2(b.x + d.y, b.y - d.x);            // b - i d
            out[j + 2 * p] = csub(a, c);
            out[j + 3 * p] = make_float2(b.x - d.y, b.y + d.x);        // b + i d
        } else {                                                        // radix 5 (forward transform)
            const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;      // cos(2 pi/5), cos(4 pi/5)
            const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;       // sin(2 pi/5), sin(4 pi/5)
            const float2 a1 = cadd(u[1], u[4]), b1 = csub(u[1], u[4]), a2 = cadd(u[2], u[3]), b2 = csub(u[2], u[3]);
            out[j] = cadd(u[0], cadd(a1, a2));
            const float2 m1 = make_float2(u[0].x + c1 * a1.x + c2 * a2.x, u[0].y + c1 * a1.y + c2 * a2.y);
            const float2 m2 = make_float2(u[0].x + c2 * a1.x + c1 * a2.x, u[0].y + c2 * a1.y + c1 * a2.y);
            const float2 n1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
            const float2 n2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
            // X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2   (-i z = (z.y, -z.x))
            out[j + p] = make_float2(m1.x + n1.y, m1.y - n1.x);
            out[j + 4 * p] = make_float2(m1.x - n1.y, m1.y + n1.x);
            out[j + 2 * p] = make_float2(m2.x + n2.y, m2.y - n2.x);
            out[j + 3 * p] = make_float2(m2.x - n2.y, m2.y + n2.x);
        }
    }
}

// FFT of one series in shared memory; returns the buffer holding the result.
// The series of a CTA are independent: each has its own named barrier (ids 1..kWelchSims), so that the warps of one series never wait
// for another series' pass.
__device__ __forceinline__ void series_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// `a` holds the output of the first pass (radix 4, done by the caller while loading) when FIRST_DONE, else the input.
template <bool FIRST_DONE>
__device__ __forceinline__ float2* fft_series(float2* a, float2* b, const WelchPlan& W, const float2* tw, int lane, int nlanes, bool active, int bar_id) {
    const int M = W.M;
    if (M == 2000) {                      // the reference's nperseg = 4000: radices 4,4,5,5,5 with compile-time strides
        if (!FIRST_DONE) {
            if (active) stockham_pass<4, 1>(a, b, M, 1, tw, lane, nlanes);
            series_sync(bar_id, nlanes);
        } else {
            float2* t = a; a = b; b = t;  // the first pass already wrote into what we call b below
        }
        if (active) stockham_pass<4, 4>(b, a, M, 4, tw, lane, nlanes);
        series_sync(bar_id, nlanes);
        if (active) stockham_pass<5, 16>(a, b, M, 16, tw, lane, nlanes);
        series_sync(bar_id, nlanes);
        if (active) stockham_pass<5, 80>(b, a, M, 80, tw, lane, nlanes);
        series_sync(bar_id, nlanes);
        if (active) stockham_pass<5, 400>(a, b, M, 400, tw, lane, nlanes);
        series_sync(bar_id, nlanes);
        return b;
    }
    int p = FIRST_DONE ? W.radix[0] : 1;
    for (int st = FIRST_DONE ? 1 : 0; st < W.nstages; ++st) {
        const int R = W.radix[st];
        if (active) {
            if (R == 4) stockham_pass<4, 0>(a, b, M, p, tw, lane, nlanes);
            else if (R == 2) stockham_pass<2, 0>(a, b, M, p, tw, lane, nlanes);
            else stockham_pass<5, 0>(a, b, M, p, tw, lane, nlanes);
        }
        p *= R;
        float2* t = a; a = b; b = t;
        series_sync(bar_id, nlanes);
    }
    return a;
}

// Series-major sample ring: wring[(node * Bs + sim) * L + (sample mod L)] (written by the BOLD/filter kernel).
// dynamic smem: 2 * S * M float2 (ping-pong) + M float2 (twiddles) + S * (M + 1) float (power accumulator) + 40 floats
__global__ void __launch_bounds__(kWelchThreads) welch_segment_kernel(const float* wring, int start, int N, int64_t Bs, int64_t sim0,
                                                                     int64_t nsim, WelchPlan W, float* P, float scale) {
    extern __shared__ __align__(16) unsigned char smw[];
    const int M = W.M, L = W.L;
    float2* buf0 = reinterpret_cast<float2*>(smw);
    float2* buf1 = buf0 + kWelchSims * M;
    float2* tw = buf1 + kWelchSims * M;
    float* pacc = reinterpret_cast<float*>(tw + M);
    float* red = pacc + kWelchSims * (M + 1);
    const int tid = threadIdx.x;
    constexpr int NL = kWelchThreads / kWelchSims;                       // threads per series
    constexpr int WPS = NL / 32;                                         // warps per series
    const int sub = tid / NL, lane = tid % NL;
    const int64_t s_base = sim0 + (int64_t)blockIdx.x * kWelchSims;
    const int ns = (int)min((int64_t)kWelchSims, sim0 + nsim - s_base);
    const bool active = sub < ns;
    for (int k = tid; k < kWelchSims * (M + 1); k += kWelchThreads) pacc[k] = 0.f;
    for (int k = tid; k < M; k += kWelchThreads) tw[k] = W.tw[k];
    const int half = start / 2;                                          // start is even: pairs never straddle the wrap
    const bool fuse4 = W.radix[0] == 4;                                  // first Stockham pass (radix 4, no twiddles) fused with the load
    const float2* w2 = reinterpret_cast<const float2*>(W.window);
    __syncthreads();                                                      // pacc and the twiddle table are set up by all threads
    for (int node = 0; node < N; ++node) {
        series_sync(1 + sub, NL);                                                  // the previous node's un-packing has read buf0/buf1
        // Load + Hann window (+ first radix-4 pass).  detrend='constant' is applied in the frequency domain: the periodic Hann
        // window's DFT is L/2 at k = 0, -L/4 at k = +-1 and 0 elsewhere, so removing the mean only changes bins 0 and 1.
        float sum = 0.f;
        if (active) {
            const float2* src = reinterpret_cast<const float2*>(wring + ((int64_t)node * Bs + s_base + sub) * L);
            if (fuse4) {
                const int t = M / 4;
                float2* out = buf1 + sub * M;
                for (int i = lane; i < t; i += NL) {
                    float2 u[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int n = i + r * t;
                        int m = n + half;
                        if (m >= M) m -= M;
                        const float2 v = src[m];
                        const float2 w = w2[n];
                        sum += v.x + v.y;
                        u[r] = make_float2(v.x * w.x, v.y * w.y);
                    }
                    const float2 a = cadd(u[0], u[2]), bq = csub(u[0], u[2]), c = cadd(u[1], u[3]), d = csub(u[1], u[3]);
                    float4* o4 = reinterpret_cast<float4*>(out + 4 * i);   // out[4 i + r], r = 0..3: two 16-byte stores
                    const float2 x0 = cadd(a, c), x1 = make_float2(bq.x + d.y, bq.y - d.x);
                    const float2 x2 = csub(a, c), x3 = make_float2(bq.x - d.y, bq.y + d.x);
                    o4[0] = make_float4(x0.x, x0.y, x1.x, x1.y);
                    o4[1] = make_float4(x2.x, x2.y, x3.x, x3.y);
                }
            } else {
                for (int n = lane; n < M; n += NL) {
                    int m = n + half;
                    if (m >= M) m -= M;
                    const float2 v = src[m];
                    const float2 w = w2[n];
                    sum += v.x + v.y;
                    buf0[sub * M + n] = make_float2(v.x * w.x, v.y * w.y);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((tid & 31) == 0) red[tid >> 5] = sum;                         // read after the FFT's barriers
        series_sync(1 + sub, NL);
        const float2* Z = fuse4 ? fft_series<true>(buf1 + sub * M, buf0 + sub * M, W, tw, lane, NL, active, 1 + sub)
                                : fft_series<false>(buf0 + sub * M, buf1 + sub * M, W, tw, lane, NL, active, 1 + sub);
        float msum = 0.f;
#pragma unroll
        for (int w = 0; w < WPS; ++w) msum += red[sub * WPS + w];
        const float mean = msum / (float)L;
        // un-pack the real FFT: X[k] = (Z[k] + conj(Z[M-k]))/2 - i w^k (Z[k] - conj(Z[M-k]))/2,  w = exp(-2 pi i / L)
        if (active) {
            for (int k = lane; k <= M; k += NL) {
                const float2 zk = Z[k == M ? 0 : k], zm = Z[k == 0 ? 0 : M - k];
                const float2 ev = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                const float2 od = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
                const float2 t = cmul(W.tw2[k], od);                      // -i * t = (t.y, -t.x)
                float xr = ev.x + t.y;
                const float xi = ev.y - t.x;
                if (k == 0) xr -= mean * (0.5f * (float)L);               // - mean * DFT(window)[0]
                if (k == 1) xr += mean * (0.25f * (float)L);              // - mean * DFT(window)[1]
                pacc[sub * (M + 1) + k] += (xr * xr + xi * xi) * ((k == 0 || k == M) ? 1.0f : 2.0f);
            }
        }
    }
    __syncthreads();
    if (active) {
        float* dst = P + (s_base + sub) * (int64_t)(M + 1);
        for (int k = lane; k <= M; k += NL) dst[k] += pacc[sub * (M + 1) + k] * scale;
    }
}

// peak[b] = df * argmax_k P[b][k]  (first maximum, like np.where(meanpow == meanpow.max())[0][0])
__global__ void welch_peak_kernel(const float* P, int nbins, double df, double* out, int64_t out_stride) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* p = P + (int64_t)b * nbins;
    float best = -1.f;
    int idx = 0x7fffffff;
    for (int k = tid; k < nbins; k += blockDim.x) {
        const float v = p[k];
        if (v > best || (v == best && k < idx)) { best = v; idx = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v = __shfl_xor_sync(0xffffffffu, best, o);
        const int i = __shfl_xor_sync(0xffffffffu, idx, o);
        if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
    if ((tid & 31) == 0) { bv[tid >> 5] = best; bi[tid >> 5] = idx; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
        out[(int64_t)b * out_stride] = df * (double)idx;
    }
}

}  // namespace nrem
