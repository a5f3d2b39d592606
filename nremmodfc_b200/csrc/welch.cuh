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
// The samples come from a ring of `ring_rows` rows laid out [row][node][sim] (the integrator's E buffer).
#pragma once
#include "common.cuh"

namespace nrem {

constexpr int kWelchSims = 4;
constexpr int kWelchThreads = 256;
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

// One Stockham pass of radix R over a series of M complex points: in -> out.  p = product of the previous radices.
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out, int M, int p, const float2* tw,
                                              int lane, int nlanes) {
    const int t = M / R;
    const int twstep = M / (p * R);
    for (int i = lane; i < t; i += nlanes) {
        const int k = i % p;
        const int j = (i - k) * R + k;
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            u[r] = in[i + r * t];
            if (r > 0) u[r] = cmul(u[r], tw[(r * k * twstep) % M]);
        }
        if (R == 2) {
            out[j] = make_float2(u[0].x + u[1].x, u[0].y + u[1].y);
            out[j + p] = make_float2(u[0].x - u[1].x, u[0].y - u[1].y);
        } else if (R == 4) {
            const float2 a = make_float2(u[0].x + u[2].x, u[0].y + u[2].y), b = make_float2(u[0].x - u[2].x, u[0].y - u[2].y);
            const float2 c = make_float2(u[1].x + u[3].x, u[1].y + u[3].y), d = make_float2(u[1].x - u[3].x, u[1].y - u[3].y);
            out[j] = make_float2(a.x + c.x, a.y + c.y);
            out[j + p] = make_float2(b.x + d.y, b.y - d.x);            // b - i d
            out[j + 2 * p] = make_float2(a.x - c.x, a.y - c.y);
            out[j + 3 * p] = make_float2(b.x - d.y, b.y + d.x);        // b + i d
        } else {                                                        // generic small DFT (R = 5)
#pragma unroll
            for (int q = 0; q < R; ++q) {
                float2 acc = u[0];
#pragma unroll
                for (int r = 1; r < R; ++r) acc = make_float2(acc.x + cmul(u[r], tw[((q * r) % R) * (M / R)]).x,
                                                               acc.y + cmul(u[r], tw[((q * r) % R) * (M / R)]).y);
                out[j + q * p] = acc;
            }
        }
    }
}

// dynamic smem: 2 * S * M float2 (ping-pong) + S * (M + 1) float (power accumulator) + 40 floats
__global__ void __launch_bounds__(kWelchThreads) welch_segment_kernel(const float* Er, int64_t ring_rows, int64_t start_row, int N, int64_t Bs,
                                                                     int64_t sim0, int64_t nsim, WelchPlan W, float* P, float scale) {
    extern __shared__ __align__(16) unsigned char smw[];
    const int M = W.M, L = W.L;
    float2* buf0 = reinterpret_cast<float2*>(smw);
    float2* buf1 = buf0 + kWelchSims * M;
    float* pacc = reinterpret_cast<float*>(buf1 + kWelchSims * M);
    float* red = pacc + kWelchSims * (M + 1);
    const int tid = threadIdx.x;
    const int64_t s_base = sim0 + (int64_t)blockIdx.x * kWelchSims;
    const int ns = (int)min((int64_t)kWelchSims, sim0 + nsim - s_base);
    for (int k = tid; k < kWelchSims * (M + 1); k += kWelchThreads) pacc[k] = 0.f;
    const int sub = tid / (kWelchThreads / kWelchSims), lane = tid % (kWelchThreads / kWelchSims);   // 64 threads per series
    constexpr int NL = kWelchThreads / kWelchSims;
    const int64_t row_stride = (int64_t)N * Bs;
    for (int node = 0; node < N; ++node) {
        __syncthreads();
        // load (packed even/odd), accumulate the segment mean
        float sum = 0.f;
        if (sub < ns) {
            const float* src = Er + (int64_t)node * Bs + s_base + sub;
            for (int n = lane; n < M; n += NL) {
                const int64_t r0 = (start_row + 2 * n) % ring_rows, r1 = (start_row + 2 * n + 1) % ring_rows;
                const float a = src[r0 * row_stride], b = src[r1 * row_stride];
                buf0[sub * M + n] = make_float2(a, b);
                sum += a + b;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((tid & 31) == 0) red[tid >> 5] = sum;
        __syncthreads();
        const float mean = (red[2 * sub] + red[2 * sub + 1]) / (float)L;           // 64 threads = 2 warps per series
        if (sub < ns) {
            for (int n = lane; n < M; n += NL) {                                    // detrend='constant', Hann window
                float2 v = buf0[sub * M + n];
                v.x = (v.x - mean) * W.window[2 * n];
                v.y = (v.y - mean) * W.window[2 * n + 1];
                buf0[sub * M + n] = v;
            }
        }
        __syncthreads();
        // Stockham passes
        float2* a = buf0 + sub * M;
        float2* b = buf1 + sub * M;
        int p = 1;
        for (int st = 0; st < W.nstages; ++st) {
            const int R = W.radix[st];
            if (sub < ns) {
                if (R == 4) stockham_pass<4>(a, b, M, p, W.tw, lane, NL);
                else if (R == 2) stockham_pass<2>(a, b, M, p, W.tw, lane, NL);
                else stockham_pass<5>(a, b, M, p, W.tw, lane, NL);
            }
            p *= R;
            float2* t = a; a = b; b = t;
            __syncthreads();
        }
        // un-pack the real FFT: X[k] = (Z[k] + conj(Z[M-k]))/2 - i w^k (Z[k] - conj(Z[M-k]))/2,  w = exp(-2 pi i / L)
        if (sub < ns) {
            for (int k = lane; k <= M; k += NL) {
                const float2 zk = a[k % M], zm = a[(M - k) % M];
                const float2 ev = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                const float2 od = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
                const float2 w = W.tw2[k];
                const float2 t = cmul(w, od);                                       // -i * t = (t.y, -t.x)
                const float xr = ev.x + t.y, xi = ev.y - t.x;
                const float pw = (xr * xr + xi * xi) * ((k == 0 || k == M) ? 1.0f : 2.0f);
                pacc[sub * (M + 1) + k] += pw;
            }
        }
    }
    __syncthreads();
    if (sub < ns) {
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
