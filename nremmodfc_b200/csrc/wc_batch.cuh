// Batched float32 Wilson-Cowan integrator — the throughput path of the sweep.
//
// Replaces wilsonCowan()+run() (netwWilsonCowanPlastic.py:77-137) for a tile of 128
// simulations per CTA.  Layout: lane = simulation, warp = (32 simulations) x (24 nodes):
//   * node indices are warp-uniform, so SC entries and the NA/ACh map values are broadcast
//     operands and the whole batch is structure-of-arrays (simulation fastest) in HBM;
//   * E, I, a_ie of a thread's 24 nodes live in registers for the whole launch;
//   * noise is generated in-kernel (philox.cuh), one Philox call per 4 nodes;
//   * E is exchanged once per Euler step through shared memory ([node/4][sim] float4).
// Variant v0 (this kernel) does the SC.E contraction on the FP32 FMA pipe with SC staged in
// shared memory; variant v1 (wc_tc.cuh) does it with tcgen05.mma and TMEM accumulators.
#pragma once
#include "philox.cuh"

namespace nrem {

constexpr int kTile = NREM_TILE_SIMS;   // simulations per CTA
constexpr int kChunk = 24;              // nodes per thread
constexpr int kNPad = 96;               // padded node count (4 chunks)
constexpr int kBatchThreads = 512;

struct BatchConst {
    float a_ee, a_ei, a_ii, P, rhoE, rE, rI, mu, sq;
    float kE, kI;        // dtSim/tauE, dtSim/tauI
    float sigI2;         // -sigmaI * log2(e)
    float E0, I0, a0;
    uint32_t k0, k1;     // Philox key
    int N;
};

struct BatchArgs {
    BatchConst c;
    float* state;              // [4][N][Bs]: E, I, a_ie base, a_ie delta (a_ie = base + delta, see wc_tc.cuh)
    const float* SCp;          // [ld][ld] zero padded (ld = 96; 128 for the node-lane kernel with N > 96)
    const float* SCimg;        // NULL, or the tcgen05 B-operand image of SCp (wc_tc.cuh: stage_sc_image_kernel), ld = 96 only
    const float* mapG;         // [n_maps][ld]
    const float* mapS;         // [n_maps][ld]
    int ld;
    const double* node_par;    // node-lane kernel only: NULL or [NREM_NODE_PARAMS][N] per-node parameter table (device)
    double dtSim;              // node-lane kernel only (per-node dtSim/tau constants)
    const float* par;          // [4][Bs]: G0, dG, sigma0, dsigma
    const int32_t* tile_map;   // [Bs/128]
    const uint64_t* streams;   // [Bs]
    int64_t Bs;
    int tile0;                 // first tile of this launch (tile = tile0 + blockIdx.x)
    int homo;                  // 1: every map entry is exactly 1 (homogeneous sweep)
    uint32_t zero;             // always 0: an operand ptxas cannot constant-fold (scheduling ties in wc_tc.cuh)
    uint32_t step0;            // global Euler step index of the first step of this launch
    int nsteps;
    int init;                  // 1: start from (E0, I0, a_ie_0) instead of loading state
    float kA;                  // dtSim / tau_ip of this phase
    int rec;                   // 1: store E before every downsamp-th step
    int rec_phase;             // (phase-local index of the first step) % downsamp
    int downsamp;
    int64_t row0;              // Ebuf row that the first recorded sample of this launch goes to
    float* Ebuf;               // [rows][N][Bs]
};

// One Euler step of one node.  netwWilsonCowanPlastic.py:80-83 with the constant divisions
// folded (kE = dtSim/tauE ...) and S(x) = 1/(1+exp(-(x-mu)sigma)) = rcp(1 + ex2((x-mu)*(-sigma*log2e))).
__device__ __forceinline__ void wc_node_update(const BatchConst& c, float& E, float& I, float& a, float coup, float z,
                                               float Gi, float sg2, float kA) {
    float x = fmaf(c.a_ee, E, c.P);
    x = fmaf(-a, I, x);
    x = fmaf(Gi, coup, x);
    x = fmaf(c.sq, z, x);
    const float SE = rcpf(1.0f + ex2f((x - c.mu) * sg2));
    const float y = fmaf(c.a_ei, E, -c.a_ii * I);
    const float SI = rcpf(1.0f + ex2f((y - c.mu) * c.sigI2));
    const float dA = I * (E - c.rhoE);
    const float En = fmaf(c.kE, fmaf(fmaf(-c.rE, E, 1.0f), SE, -E), E);
    const float In = fmaf(c.kI, fmaf(fmaf(-c.rI, I, 1.0f), SI, -I), I);
    a = fmaf(kA, dA, a);
    E = En;
    I = In;
}

constexpr int kV0SmemBytes = kNPad * kNPad * 4 + 2 * (kNPad / 4) * kTile * 16 + 2 * kNPad * 4;

__global__ void __launch_bounds__(kBatchThreads, 1) wc_batch_v0_kernel(const BatchArgs A) {
    extern __shared__ __align__(128) unsigned char smraw[];
    float* SCs = reinterpret_cast<float*>(smraw);                               // [96][96]
    float4* Es = reinterpret_cast<float4*>(smraw + kNPad * kNPad * 4);          // [2][24][128]
    float* mG = reinterpret_cast<float*>(smraw + kNPad * kNPad * 4 + 2 * (kNPad / 4) * kTile * 16);
    float* mS = mG + kNPad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = warp >> 2;
    const int simt = ((warp & 3) << 5) | lane;
    const int tile = A.tile0 + (int)blockIdx.x;
    const int64_t sim = (int64_t)tile * kTile + simt;
    const BatchConst& c = A.c;
    const int N = c.N;

    for (int k = tid; k < kNPad * kNPad; k += kBatchThreads) SCs[k] = A.SCp[k];
    const int mid = A.tile_map[tile];
    if (tid < kNPad) { mG[tid] = A.mapG[mid * kNPad + tid]; mS[tid] = A.mapS[mid * kNPad + tid]; }

    float E[kChunk], I[kChunk], a[kChunk];
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
        const int node = chunk * kChunk + k;
        if (node < N) {
            if (A.init) { E[k] = c.E0; I[k] = c.I0; a[k] = c.a0; }
            else {
                E[k] = A.state[(0 * (int64_t)N + node) * A.Bs + sim];
                I[k] = A.state[(1 * (int64_t)N + node) * A.Bs + sim];
                a[k] = A.state[(2 * (int64_t)N + node) * A.Bs + sim] + A.state[(3 * (int64_t)N + node) * A.Bs + sim];
            }
        } else { E[k] = 0.f; I[k] = 0.f; a[k] = 0.f; }
    }
    const float G0 = A.par[sim], dG = A.par[A.Bs + sim];
    const float sg0 = __fmul_rn(-1.4426950408889634f, A.par[2 * A.Bs + sim]), dsg = __fmul_rn(-1.4426950408889634f, A.par[3 * A.Bs + sim]);
    const uint64_t strm = A.streams[sim];
    const uint32_t s_lo = (uint32_t)strm, s_hi = (uint32_t)(strm >> 32);
    const int njg = (N + 3) >> 2;
    int rc = A.rec_phase;
    int64_t row = A.row0;
    __syncthreads();

    for (int it = 0; it < A.nsteps; ++it) {
        float4* Eb = Es + (it & 1) * (kNPad / 4) * kTile;
#pragma unroll
        for (int g = 0; g < kChunk / 4; ++g)
            Eb[(chunk * (kChunk / 4) + g) * kTile + simt] = make_float4(E[4 * g], E[4 * g + 1], E[4 * g + 2], E[4 * g + 3]);
        __syncthreads();
        if (A.rec) {
            if (rc == 0) {
#pragma unroll
                for (int k = 0; k < kChunk; ++k) {
                    const int node = chunk * kChunk + k;
                    if (node < N) A.Ebuf[(row * N + node) * A.Bs + sim] = E[k];
                }
                ++row;
            }
            if (++rc == A.downsamp) rc = 0;
        }
        // coupling: cp[k] = sum_j SC[node_k][j] * E_sim[j]   (np.dot(CM, E), netwWilsonCowanPlastic.py:81)
        float cp[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k) cp[k] = 0.f;
        const float* scrow = SCs + chunk * kChunk * kNPad;
#pragma unroll 2
        for (int jg = 0; jg < njg; ++jg) {
            const float4 e = Eb[jg * kTile + simt];
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const float4 s = *reinterpret_cast<const float4*>(scrow + k * kNPad + 4 * jg);
                cp[k] = fmaf(s.x, e.x, cp[k]);
                cp[k] = fmaf(s.y, e.y, cp[k]);
                cp[k] = fmaf(s.z, e.z, cp[k]);
                cp[k] = fmaf(s.w, e.w, cp[k]);
            }
        }
        const uint32_t step = A.step0 + (uint32_t)it;
#pragma unroll
        for (int g = 0; g < kChunk / 4; ++g) {
            const int q = chunk * (kChunk / 4) + g;
            float z[4];
            normals4f(philox4x32(step, (uint32_t)q, s_lo, s_hi, c.k0, c.k1), z[0], z[1], z[2], z[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * g + j;
                const int node = 4 * q + j;
                const float Gi = fmaf(dG, mG[node], G0);
                const float sg2 = fmaf(dsg, mS[node], sg0);
                wc_node_update(c, E[k], I[k], a[k], cp[k], z[j], Gi, sg2, A.kA);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
        const int node = chunk * kChunk + k;
        if (node < N) {
            A.state[(0 * (int64_t)N + node) * A.Bs + sim] = E[k];
            A.state[(1 * (int64_t)N + node) * A.Bs + sim] = I[k];
            A.state[(2 * (int64_t)N + node) * A.Bs + sim] = a[k];
            A.state[(3 * (int64_t)N + node) * A.Bs + sim] = 0.f;       // this validation kernel integrates a_ie directly in float32
        }
    }
}

// final_state [3][N][Bs] = (E, I, a_base + delta) from the 4-component internal state
__global__ void combine_state_kernel(const float* st4, int64_t n, float* out3) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out3[k] = st4[k];
    out3[n + k] = st4[n + k];
    out3[2 * n + k] = st4[2 * n + k] + st4[3 * n + k];
}

// ---- host-side staging kernels (float64 API arrays -> padded float32 device layout) ----------
__global__ void stage_sc_kernel(const double* CM, int N, int ld, float* SCp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ld * ld) return;
    const int i = k / ld, j = k % ld;
    SCp[k] = (i < N && j < N) ? (float)CM[(size_t)i * N + j] : 0.f;
}
__global__ void stage_maps_kernel(const double* mapG, const double* mapS, int n_maps, int N, int ld, float* oG, float* oS) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_maps * ld) return;
    const int m = k / ld, i = k % ld;
    oG[k] = i < N ? (float)mapG[(size_t)m * N + i] : 0.f;
    oS[k] = i < N ? (float)mapS[(size_t)m * N + i] : 0.f;
}
// flag[0] (zeroed by the caller) becomes 1 when some real map entry differs from 1
__global__ void maps_not_all_ones_kernel(const float* mG, const float* mS, int n_maps, int N, int ld, int* flag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_maps * ld) return;
    if ((k % ld) < N && (mG[k] != 1.0f || mS[k] != 1.0f)) flag[0] = 1;
}
__global__ void stage_par_kernel(const double* G0, const double* dG, const double* s0, const double* ds,
                                 const uint64_t* streams, int B, int64_t Bs, float* par, uint64_t* st) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Bs) return;
    const int64_t src = k < B ? k : B - 1;              // padding simulations repeat the last real one
    par[k] = (float)G0[src];
    par[Bs + k] = (float)dG[src];
    par[2 * Bs + k] = (float)s0[src];
    par[3 * Bs + k] = (float)ds[src];
    st[k] = streams[src];
}

}  // namespace nrem
