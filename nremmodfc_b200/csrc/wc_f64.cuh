// Float64 Wilson-Cowan integrator with the reference's exact call shape: replaces
// run() (netwWilsonCowanPlastic.py:86-137) and wilsonCowan() (:77-83).
//
// One CTA per simulation, one thread per node; E is exchanged through shared memory once per
// Euler step.  This is the parity path (injected reference noise, 1e-6 trajectory tolerance)
// and the backing of the single-run compatibility API; the throughput path is wc_batch.cuh.
#pragma once
#include "philox.cuh"

namespace nrem {

struct WcF64Args {
    nrem_wc_params p;
    const double* CM;
    const double* G;
    const double* sg;
    const uint64_t* streams;
    const double* noise;
    const double* node_par;     // NULL, or [NREM_NODE_PARAMS][N] per-node overrides of the scalar parameters (include/nremfc.h)
    int noise_batch;
    int64_t nrec;
    double* Y;
    double* fin;
};

__device__ __forceinline__ double sigm(double x, double sigma, double mu) { return 1.0 / (1.0 + exp(-(x - mu) * sigma)); }

// one standard-normal draw of the Philox stream for (step, node i): the float64 form of philox.cuh's normals4f
__device__ __forceinline__ double philox_normal_f64(uint32_t step, int i, uint64_t strm, uint32_t k0, uint32_t k1) {
    const Philox4 r = philox4x32(step, (uint32_t)(i >> 2), (uint32_t)strm, (uint32_t)(strm >> 32), k0, k1);
    const bool second = (i & 2) != 0;
    const double rad = sqrt(-2.0 * log(u23d(second ? r.z : r.x)));
    const double ang = 6.283185307179586476925286766559 * (u23d(second ? r.w : r.y) - 0.5);
    double sn, cs;
    sincos(ang, &sn, &cs);
    return rad * ((i & 1) ? sn : cs);
}

// dynamic smem: Es[2][N], nzs[2][N], then (CM_SMEM) CMt[N][N] with CMt[j*N+i] = CM[i][j]; otherwise A.CM already IS the transpose
// (global memory, coalesced over i).
// NOISE_WARPS: the block has a second set of round_up(N, 32) threads that only draw the Philox noise of step t + 1 into shared memory
// while the first set integrates step t.  A float64 step is one long dependent chain (log, sqrt, sincos, two exp, divisions: the
// time loop of ONE simulation cannot be parallelised), and the noise is the half of it that does not depend on the state; same
// arithmetic, bit-identical results, the single run of BASELINE configs[0] takes about half the time.
template <bool CM_SMEM, bool NOISE_WARPS>
__global__ void __launch_bounds__(CM_SMEM ? 512 : 1024) wc_run_f64_kernel(const WcF64Args A) {
    extern __shared__ double sm64[];
    const nrem_wc_params& p = A.p;
    const int N = p.nnodes;
    const int NT = NOISE_WARPS ? (int)(blockDim.x >> 1) : (int)blockDim.x;
    const bool producer = NOISE_WARPS && (int)threadIdx.x >= NT;
    const int i = producer ? (int)threadIdx.x - NT : (int)threadIdx.x;
    const int b = blockIdx.x;
    double* Es = sm64;
    double* nzs = sm64 + 2 * N;
    double* CMt = sm64 + 4 * N;
    if (CM_SMEM) {
        for (int k = threadIdx.x; k < N * N; k += blockDim.x) {
            const int r = k / N, c = k % N;
            CMt[c * N + r] = A.CM[k];
        }
    }
    const bool live = i < N && !producer;
    double E = p.E0, I = p.I0, a = p.a_ie_0;
    // "Any of them can be redefined as a vector of length nnodes" (netwWilsonCowanPlastic.py:21)
    auto npar = [&](int k, double scalar) { return (A.node_par && live) ? A.node_par[(size_t)k * N + i] : scalar; };
    const double a_ee = npar(0, p.a_ee), a_ei = npar(1, p.a_ei), a_ii = npar(2, p.a_ii), tauE = npar(3, p.tauE), tauI = npar(4, p.tauI);
    const double Pn = npar(5, p.P), rhoE = npar(6, p.rhoE), rE = npar(7, p.rE), rI = npar(8, p.rI), mu = npar(9, p.mu), sigmaI = npar(10, p.sigmaI);
    a = npar(11, p.a_ie_0);                                   // a_ie = a_ie_0*np.ones(N) (netwWilsonCowanPlastic.py:93) broadcasts a vector
    const double G = live ? A.G[(size_t)b * N + i] : 0.0;
    const double sg = live ? A.sg[(size_t)b * N + i] : 1.0;
    const uint64_t strm = A.streams ? A.streams[b] : (uint64_t)b;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const int64_t nsteps_total = p.n1 + p.n2 + p.n3;
    const double* nz_base = A.noise ? A.noise + (size_t)(b % A.noise_batch) * nsteps_total * N : nullptr;
    const int64_t ns[3] = {p.n1, p.n2, p.n3};
    int64_t step = 0;
    int buf = 0;
    if (producer && i < N && nsteps_total > 0) nzs[i] = philox_normal_f64(0u, i, strm, k0, k1);
    __syncthreads();
    for (int ph = 0; ph < 3; ++ph) {
        const double tau_ip = p.tau_ip[ph];
        for (int64_t it = 0; it < ns[ph]; ++it, ++step) {
            if (live) Es[buf * N + i] = E;
            __syncthreads();
            if (producer) {
                // noise of the NEXT step into the other buffer (its readers passed the barrier above one step ago)
                if (i < N && step + 1 < nsteps_total) nzs[(buf ^ 1) * N + i] = philox_normal_f64((uint32_t)(step + 1), i, strm, k0, k1);
                buf ^= 1;
                continue;
            }
            if (ph == 2 && live && A.Y && (it % p.downsamp) == 0) {
                const int64_t r = it / p.downsamp;
                if (r < A.nrec) {
                    double* y = A.Y + (((size_t)b * A.nrec + r) * 3) * N + i;
                    y[0] = E; y[N] = I; y[2 * N] = a;
                }
            }
            if (live) {
                const double* e = Es + buf * N;
                double acc0 = 0.0, acc1 = 0.0;
                if (CM_SMEM) {
                    int j = 0;
                    for (; j + 1 < N; j += 2) {
                        acc0 = fma(CMt[j * N + i], e[j], acc0);
                        acc1 = fma(CMt[(j + 1) * N + i], e[j + 1], acc1);
                    }
                    if (j < N) acc0 = fma(CMt[j * N + i], e[j], acc0);
                } else {
                    const double* col = A.CM + i;                       // transposed copy: CMt[j*N + i]
                    int j = 0;
                    for (; j + 1 < N; j += 2) {
                        acc0 = fma(__ldg(col + (size_t)j * N), e[j], acc0);
                        acc1 = fma(__ldg(col + (size_t)(j + 1) * N), e[j + 1], acc1);
                    }
                    if (j < N) acc0 = fma(__ldg(col + (size_t)j * N), e[j], acc0);
                }
                const double coup = acc0 + acc1;
                double nz;
                if (nz_base) {
                    nz = nz_base[(size_t)step * N + i];
                } else if (NOISE_WARPS) {
                    nz = p.sqdtD * nzs[buf * N + i];
                } else {
                    nz = p.sqdtD * philox_normal_f64((uint32_t)step, i, strm, k0, k1);
                }
                const double dE = (-E + (1 - rE * E) * sigm(a_ee * E - a * I + G * coup + Pn + nz, sg, mu)) / tauE;
                const double dI = (-I + (1 - rI * I) * sigm(a_ei * E - a_ii * I, sigmaI, mu)) / tauI;
                const double da = (I * (E - rhoE)) / tau_ip;
                E += p.dtSim * dE;
                I += p.dtSim * dI;
                a += p.dtSim * da;
            }
            buf ^= 1;
        }
    }
    if (live && A.fin) {
        double* f = A.fin + (size_t)b * 3 * N + i;
        f[0] = E; f[N] = I; f[2 * N] = a;
    }
}

__global__ void transpose_f64_kernel(const double* in, int N, double* out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N * N) out[(size_t)(k % N) * N + k / N] = in[k];
}

__global__ void wc_derivative_f64_kernel(const nrem_wc_params p, const double* CM, const double* X, const double* G,
                                         const double* sg, const double* noise, double tau_ip, double* dX) {
    const int N = p.nnodes;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double* Ev = X;
    double acc = 0.0;
    for (int j = 0; j < N; ++j) acc = fma(CM[(size_t)i * N + j], Ev[j], acc);
    const double E = X[i], I = X[N + i], a = X[2 * N + i];
    const double nz = noise ? noise[i] : 0.0;
    dX[i] = (-E + (1 - p.rE * E) * sigm(p.a_ee * E - a * I + G[i] * acc + p.P + nz, sg[i], p.mu)) / p.tauE;
    dX[N + i] = (-I + (1 - p.rI * I) * sigm(p.a_ei * E - p.a_ii * I, p.sigmaI, p.mu)) / p.tauI;
    dX[2 * N + i] = (I * (E - p.rhoE)) / tau_ip;
}

}  // namespace nrem
