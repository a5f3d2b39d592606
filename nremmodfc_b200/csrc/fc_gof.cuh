// FC = Pearson correlation over time, and goodness of fit against K empirical matrices.
//
// Replaces np.corrcoef(BOLD.T) (whole_sweep_both.py:81), utils.get_all_metrics
// (utils.py:42-50, incl. new_metric utils.py:28-31 and scikit-image's structural_similarity
// at its defaults, call site utils.py:48) and sFC.mean() (whole_sweep_both.py:94).
// One CTA per simulation; all reductions are warp-shuffle + one shared-memory hop, float64.
#pragma once
#include "common.cuh"

namespace nrem {

constexpr int kFcThreads = 512;
constexpr int kFcMaxPairs = 17;     // ceil(N(N+1)/2 / 512) for N <= 128
constexpr int kFcTile = 32;         // time rows per shared-memory tile

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in every thread.  red: >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// bold [B][J][N] -> fc [B][N][N].   dynamic smem: (2*N + kFcTile*N) doubles
__global__ void __launch_bounds__(kFcThreads) fc_f64_kernel(const double* bold, int64_t J, int N, double* fc) {
    extern __shared__ double smf[];
    double* mean = smf;
    double* sd = smf + N;
    double* tile = smf + 2 * N;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* x = bold + (size_t)b * J * N;
    if (tid < N) {
        double m = 0.0;
        for (int64_t t = 0; t < J; ++t) m += x[t * N + tid];
        mean[tid] = m / (double)J;
    }
    const int npairs = N * (N + 1) / 2;
    int pi[kFcMaxPairs], pj[kFcMaxPairs];
    double acc[kFcMaxPairs];
#pragma unroll
    for (int q = 0; q < kFcMaxPairs; ++q) {
        acc[q] = 0.0;
        int p = tid + q * kFcThreads;
        int i = 0;
        if (p < npairs) {
            // row-major upper triangle incl. diagonal: row i holds N - i entries
            int rem = p;
            while (rem >= N - i) { rem -= N - i; ++i; }
            pi[q] = i; pj[q] = i + rem;
        } else {
            pi[q] = -1; pj[q] = 0;
        }
    }
    __syncthreads();
    for (int64_t t0 = 0; t0 < J; t0 += kFcTile) {
        const int rows = (int)min((int64_t)kFcTile, J - t0);
        for (int k = tid; k < rows * N; k += kFcThreads) tile[k] = x[t0 * N + k] - mean[k % N];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kFcMaxPairs; ++q) {
            if (pi[q] >= 0) {
                double a = acc[q];
                for (int r = 0; r < rows; ++r) a = fma(tile[r * N + pi[q]], tile[r * N + pj[q]], a);
                acc[q] = a;
            }
        }
        __syncthreads();
    }
    const double inv = 1.0 / (double)(J - 1);              // np.cov: c *= 1/(J-1)
#pragma unroll
    for (int q = 0; q < kFcMaxPairs; ++q) {
        acc[q] *= inv;
        if (pi[q] >= 0 && pi[q] == pj[q]) sd[pi[q]] = sqrt(acc[q]);
    }
    __syncthreads();
    double* o = fc + (size_t)b * N * N;
#pragma unroll
    for (int q = 0; q < kFcMaxPairs; ++q) {
        if (pi[q] >= 0) {
            double c = acc[q] / sd[pi[q]];                 // np.corrcoef: c /= stddev[:,None]; c /= stddev[None,:]
            c = c / sd[pj[q]];
            c = fmin(1.0, fmax(-1.0, c));                  // np.clip(c.real, -1, 1)
            o[pi[q] * N + pj[q]] = c;
            o[pj[q] * N + pi[q]] = c;
        }
    }
}

// Any N: one CTA per 32 x 32 tile of the FC matrix of one simulation (grid (N/32, N/32, B), 256 threads, 2 x 2 entries per
// thread).  Means, centred 32-sample tiles and the diagonal c_ii / c_jj are computed in the CTA with the same operation order as
// the small kernel (sum over time ascending, fma), so both give the same numbers.
__global__ void __launch_bounds__(256) fc_big_f64_kernel(const double* bold, int64_t J, int N, double* fc) {
    __shared__ double mi[32], mj[32], si[32], sj[32];
    __shared__ double Xi[32][33], Xj[32][33];
    const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32, tid = threadIdx.x;
    const double* x = bold + (size_t)b * J * N;
    if (tid < 64) {
        const int node = (tid < 32 ? i0 : j0 - 32) + tid;
        double m = 0.0;
        if (node < N) for (int64_t t = 0; t < J; ++t) m += x[t * N + node];
        (tid < 32 ? mi[tid] : mj[tid - 32]) = m / (double)J;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, dacc = 0.0;
    for (int64_t t0 = 0; t0 < J; t0 += 32) {
        const int rows = (int)min((int64_t)32, J - t0);
        for (int k = tid; k < 32 * 32; k += 256) {
            const int r = k >> 5, c = k & 31;
            Xi[r][c] = (r < rows && i0 + c < N) ? x[(t0 + r) * N + i0 + c] - mi[c] : 0.0;
            Xj[r][c] = (r < rows && j0 + c < N) ? x[(t0 + r) * N + j0 + c] - mj[c] : 0.0;
        }
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            const double a0 = Xi[r][2 * ty], a1 = Xi[r][2 * ty + 1], b0 = Xj[r][2 * tx], b1 = Xj[r][2 * tx + 1];
            acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
            if (tid < 32) dacc = fma(Xi[r][tid], Xi[r][tid], dacc);
            else if (tid < 64) dacc = fma(Xj[r][tid - 32], Xj[r][tid - 32], dacc);
        }
        __syncthreads();
    }
    const double inv = 1.0 / (double)(J - 1);              // np.cov: c *= 1/(J-1)
    if (tid < 32) si[tid] = sqrt(dacc * inv);
    else if (tid < 64) sj[tid - 32] = sqrt(dacc * inv);
    __syncthreads();
    double* o = fc + (size_t)b * N * N;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = i0 + 2 * ty + u, j = j0 + 2 * tx + v;
            if (i < N && j < N) {
                // np.corrcoef: c /= stddev[:,None]; c /= stddev[None,:] -- in the order of the upper triangle on both sides of the
                // diagonal, so that the matrix is exactly symmetric (as the small kernel's mirrored store makes it)
                const double s1 = i <= j ? si[2 * ty + u] : sj[2 * tx + v], s2 = i <= j ? sj[2 * tx + v] : si[2 * ty + u];
                double c = acc[u][v] * inv / s1;
                c = c / s2;
                o[(size_t)i * N + j] = fmin(1.0, fmax(-1.0, c));
            }
        }
}

// The four metrics of utils.get_all_metrics for ONE (simulated, empirical) pair; S and Em are N x N row-major in shared or
// global memory.  Called by every thread of the CTA; thread 0 writes o[0..3] = (corr, euc, ssim, new_metric).
__device__ __forceinline__ void gof_one_target(const double* __restrict__ S, const double* __restrict__ Em, int N, double data_range,
                                               double* red, double* o) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int NN = N * N;
    const double P = 0.5 * (double)N * (double)(N - 1);
    const int W = N - 6;                                      // 7x7 windows fully inside: (N-6)^2
    const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
    const double cov_norm = 49.0 / 48.0;                      // sample covariance (use_sample_covariance=True)
    // strict upper triangle, two-pass Pearson (np.corrcoef of the two flattened vectors)
    double ss = 0.0, se = 0.0;
    for (int k = tid; k < NN; k += nt) {
        const int i = k / N, j = k % N;
        if (j > i) { ss += S[k]; se += Em[k]; }
    }
    const double ms = block_sum(ss, red) / P;
    const double me = block_sum(se, red) / P;
    double css = 0.0, cee = 0.0, cse = 0.0, d2 = 0.0;
    for (int k = tid; k < NN; k += nt) {
        const int i = k / N, j = k % N;
        if (j > i) {
            const double a = S[k] - ms, c = Em[k] - me, d = Em[k] - S[k];
            css = fma(a, a, css); cee = fma(c, c, cee); cse = fma(a, c, cse); d2 = fma(d, d, d2);
        }
    }
    css = block_sum(css, red); cee = block_sum(cee, red); cse = block_sum(cse, red); d2 = block_sum(d2, red);
    // SSIM, uniform 7x7 window, mean over the (N-6)^2 interior positions
    double sacc = 0.0;
    for (int wdx = tid; wdx < W * W; wdx += nt) {
        const int r0 = wdx / W, c0 = wdx % W;
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int r = 0; r < 7; ++r) {
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const double xv = S[(r0 + r) * N + c0 + c], yv = Em[(r0 + r) * N + c0 + c];
                sx += xv; sy += yv;
                sxx = fma(xv, xv, sxx); syy = fma(yv, yv, syy); sxy = fma(xv, yv, sxy);
            }
        }
        const double ux = sx / 49.0, uy = sy / 49.0, uxx = sxx / 49.0, uyy = syy / 49.0, uxy = sxy / 49.0;
        const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
        sacc += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
    sacc = block_sum(sacc, red);
    if (tid == 0) {
        const double corr = (cse / (P - 1)) / sqrt(css / (P - 1)) / sqrt(cee / (P - 1));
        o[0] = corr;
        o[1] = sqrt(d2);
        o[2] = sacc / (double)((double)W * (double)W);
        o[3] = (1.0 - corr) + (ms - me) * (ms - me);
    }
}

// fc [B][N][N], emp [K][N][N] -> gof [B][K][4] = (corr, euc, ssim, new_metric), meanfc [B].
// dynamic smem: (2*N*N + 40) doubles with emp_smem (N <= 118), else (N*N + 40) and the target is read from global memory
__global__ void __launch_bounds__(256) gof_f64_kernel(const double* fc, const double* emp, int K, int N, double data_range,
                                                      double* gof, double* meanfc, int emp_smem) {
    extern __shared__ double smg[];
    double* S = smg;
    double* Ems = smg + N * N;
    double* red = smg + (emp_smem ? 2 : 1) * N * N;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int NN = N * N;
    for (int k = tid; k < NN; k += nt) S[k] = fc[(size_t)b * NN + k];
    __syncthreads();
    {
        double s = 0.0;
        for (int k = tid; k < NN; k += nt) s += S[k];
        s = block_sum(s, red);
        if (tid == 0 && meanfc) meanfc[b] = s / (double)NN;
    }
    for (int kt = 0; kt < K; ++kt) {
        __syncthreads();
        const double* Em = emp + (size_t)kt * NN;
        if (emp_smem) {
            for (int k = tid; k < NN; k += nt) Ems[k] = emp[(size_t)kt * NN + k];
            Em = Ems;
        }
        __syncthreads();
        gof_one_target(S, Em, N, data_range, red, gof + ((size_t)b * K + kt) * 4);
    }
}

// Any N (other parcellations, BASELINE configs[4]): FC and the target stay in global memory and are read through L1/L2.
// grid (B, K): one CTA per (simulation, target); meanfc is written by the CTAs of target 0.  smem: 40 doubles.
__global__ void __launch_bounds__(1024) gof_big_f64_kernel(const double* fc, const double* emp, int K, int N, double data_range,
                                                           double* gof, double* meanfc) {
    __shared__ double red[40];
    const int b = blockIdx.x, kt = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const size_t NN = (size_t)N * N;
    const double* S = fc + (size_t)b * NN;
    if (kt == 0 && meanfc) {
        double s = 0.0;
        for (size_t k = tid; k < NN; k += nt) s += S[k];
        s = block_sum(s, red);
        if (tid == 0) meanfc[b] = s / (double)NN;
    }
    gof_one_target(S, emp + (size_t)kt * NN, N, data_range, red, gof + ((size_t)b * K + kt) * 4);
}

// Kuramoto order parameter of the Hilbert phases (utils.py:34-40: hilbert -> angle -> |mean exp(i angle)| -> mean, std).
// bold [B][J][N]; g[J] = imag(ifft(h)) is the circular Hilbert kernel of scipy.signal.hilbert for length J, so that
// imag(analytic)[t] = sum_s g[(t - s) mod J] x[s].  One CTA per simulation, one thread per time point.
// dynamic smem: (J + 40) doubles.  out [B][2] = (sync, meta).
__global__ void kuramoto_f64_kernel(const double* bold, const double* g, int J, int N, double* out) {
    extern __shared__ double smk[];
    double* gs = smk;
    double* red = smk + J;
    const int b = blockIdx.x, t = threadIdx.x;
    for (int k = t; k < J; k += blockDim.x) gs[k] = g[k];
    __syncthreads();
    const double* x = bold + (size_t)b * J * N;
    double kur = 0.0;
    if (t < J) {
        double cr = 0.0, ci = 0.0;
        for (int n = 0; n < N; ++n) {
            double im = 0.0;
            int idx = t;                                   // (t - s) mod J for s = 0
            for (int s_ = 0; s_ < J; ++s_) {
                im = fma(gs[idx], x[(size_t)s_ * N + n], im);
                idx = idx == 0 ? J - 1 : idx - 1;
            }
            const double re = x[(size_t)t * N + n];
            const double mag = sqrt(re * re + im * im);
            if (mag > 0.0) { cr += re / mag; ci += im / mag; } else { cr += 1.0; }     // np.angle(0) = 0
        }
        kur = sqrt(cr * cr + ci * ci) / (double)N;
    }
    const double mean = block_sum(t < J ? kur : 0.0, red) / (double)J;
    const double d = t < J ? kur - mean : 0.0;
    const double var = block_sum(d * d, red) / (double)J;                              // np.std: population
    if (t == 0) { out[2 * b] = mean; out[2 * b + 1] = sqrt(var); }
}

}  // namespace nrem
