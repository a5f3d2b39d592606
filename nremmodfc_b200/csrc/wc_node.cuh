// Node-lane variant of the batched integrator: the contraction is issued transposed,
//     D^T[128 nodes, S sims] = SC[128, KP] x E^T[KP, S]        (tcgen05.mma kind::tf32, M = 128, N = S, K = 8, 3xTF32 split)
// so that a TMEM lane is an output NODE and a column a simulation.  Thread = one node x SPT simulations
// (warp = node quadrant q = warp & 3, simulation group g = warp >> 2; tile = S = 4*SPT simulations per CTA).
//
// What that buys over the simulation-lane kernel (wc_tc.cuh, 128 simulations per CTA, 24 nodes per thread):
//   * small tiles: 16 or 32 simulations per CTA, so a SMALL batch (run_many_seeds.py: 200 simulations; one GPU's share of a
//     strong-scaled sweep) spreads over 4-8x more SMs and every Euler step handles 4-8x less work per thread — the time loop is
//     sequential, so per-step latency is the only lever for such batches;
//   * connectomes up to 128 nodes (M = 128 lanes; wc_tc.cuh is one 96-wide MMA tile);
//   * every node parameter as a per-node vector (netwWilsonCowanPlastic.py:21) for free: a node's parameters are per-THREAD
//     registers here, per-simulation quantities are what gets broadcast.
// The arithmetic per (simulation, node, step) is the same sequence of float32 operations as in wc_tc.cuh, on the same Philox
// stream (counter = step, node/4, replicate id), and the 3xTF32 passes accumulate in the same order — results are compared bit for
// bit with the 128-simulation kernel in tests/test_gpu_parity.py.
//
// Noise: one Philox call yields the normals of 4 consecutive nodes of one simulation, i.e. of 4 adjacent LANES here.  Lane r of a
// lane quad draws the calls of simulations r, r+4 (of the thread's SPT), runs Box-Muller on all four outputs, and the quad
// transposes through a per-warp shared-memory tile (one STS.128 per call, one LDS.32 per node update; rows swizzled by 2*sim so
// that both directions are conflict-free).
#pragma once
#include "wc_tc.cuh"

namespace nrem {

constexpr int kNodeThreads = 512;
constexpr uint32_t kNodeLboA = 128 * 16;            // next 4-column group of the A operand (128 rows x 16 B)

// Stage SC ([ld][ld] float32, zero padded) as the A operand(s): A[m][k] = SC[m][k], K-major canonical layout.
template <int KP>
__device__ __forceinline__ void node_stage_a(const float* SCp, int ld, float* Ah, float* Al, int tid) {
    for (int idx = tid; idx < 128 * KP; idx += kNodeThreads) {
        const int m = idx / KP, k = idx % KP;
        const float v = (m < ld && k < ld) ? SCp[m * ld + k] : 0.f;
        const float h = tf32_rn(v);
        const int o = (k >> 2) * (int)(kNodeLboA / 4) + m * 4 + (k & 3);
        Ah[o] = h;
        Al[o] = v - h;
    }
}

template <int SPT>
__device__ __forceinline__ void tmem_ld_spt(uint32_t taddr, uint32_t (&r)[8]) {
    if constexpr (SPT == 8) tmem_ld8(taddr, r); else tmem_ld4(taddr, r);
}

template <int SPT, int KQ>
constexpr int node_smem_bytes() {
    return 2 * (32 * KQ / 4) * (int)kNodeLboA + 2 * (32 * KQ / 4) * (4 * SPT * 16 + 16) + 16 * SPT * 32 * 4 + 64;
}

// SPT: simulations per thread (4 or 8) -> tile of 16 or 32 simulations.  KQ: live node quadrants (3: N <= 96, 4: N <= 128).
template <int SPT, int KQ>
__global__ void __launch_bounds__(kNodeThreads, 1) wc_node_kernel(const BatchArgs A) {
    constexpr int S = 4 * SPT;
    constexpr int KP = 32 * KQ;
    constexpr uint32_t LBO_B = S * 16 + 16;            // padded by one 16-byte slot: the 8 four-node groups of a warp land in 8 bank quads
    constexpr uint32_t A_BYTES = (KP / 4) * kNodeLboA, B_BYTES = (KP / 4) * LBO_B;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(S >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int ISSUER_WARP = KQ == 3 ? 3 : 15;      // N <= 96: a warp of the empty quadrant issues; else the last warp does both jobs
    constexpr int BAR_THREADS = (KQ == 3 ? 13 : 16) * 32;
    extern __shared__ __align__(128) unsigned char smraw[];
    float* Ah = reinterpret_cast<float*>(smraw);
    float* Al = reinterpret_cast<float*>(smraw + A_BYTES);
    float* Bh = reinterpret_cast<float*>(smraw + 2 * A_BYTES);
    float* Bl = reinterpret_cast<float*>(smraw + 2 * A_BYTES + B_BYTES);
    float* Zx = reinterpret_cast<float*>(smraw + 2 * A_BYTES + 2 * B_BYTES);          // [16 warps][SPT][32] normals in transit
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + 2 * A_BYTES + 2 * B_BYTES + 16 * SPT * 32 * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, g = warp >> 2;
    const int node = 32 * q + lane;
    const int tile = A.tile0 + (int)blockIdx.x;
    const int64_t sim0 = (int64_t)tile * S + g * SPT;          // first of this thread's simulations
    const BatchConst& c = A.c;
    const int N = c.N, ld = A.ld;
    const bool live_warp = q < KQ;
    const bool live = live_warp && node < N;

    node_stage_a<KP>(A.SCp, ld, Ah, Al, tid);
    for (int k = tid; k < (int)(2 * B_BYTES / 4); k += kNodeThreads) Bh[k] = 0.f;
    if (warp == 0) tmem_alloc(tmem_slot, 32);
    if (tid == 32) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (live_warp || warp == ISSUER_WARP) {
        const uint32_t tmem_mine = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * SPT);
        // ---- per-node constants: the scalars of the launch, or this node's entries of the per-node table -------------------
        float a_ee = c.a_ee, a_ei = c.a_ei, a_ii = c.a_ii, kE = c.kE, kI = c.kI, Pn = c.P, rhoE = c.rhoE, rE = c.rE, rI = c.rI, mu = c.mu,
              sigI2 = c.sigI2, a0 = c.a0;
        if (A.node_par && live) {
            const double* np = A.node_par + node;            // [NREM_NODE_PARAMS][N]: a_ee a_ei a_ii tauE tauI P rhoE rE rI mu sigmaI a_ie_0
            a_ee = (float)np[0]; a_ei = (float)np[(size_t)1 * N]; a_ii = (float)np[(size_t)2 * N];
            kE = (float)(A.dtSim / np[(size_t)3 * N]); kI = (float)(A.dtSim / np[(size_t)4 * N]);
            Pn = (float)np[(size_t)5 * N]; rhoE = (float)np[(size_t)6 * N]; rE = (float)np[(size_t)7 * N]; rI = (float)np[(size_t)8 * N];
            mu = (float)np[(size_t)9 * N]; sigI2 = (float)(-np[(size_t)10 * N] * 1.4426950408889634); a0 = (float)np[(size_t)11 * N];
        }
        const float Pmu = Pn - mu, nmu = -mu, nkr = -A.kA * rhoE;
        const float two_pi = 6.2831853071795865f, m2ln2 = -1.3862943611198906f;
        const int mid = A.tile_map[(tile * S) / kTile];
        const float mGn = live ? A.mapG[mid * ld + node] : 0.f, mSn = live ? A.mapS[mid * ld + node] : 0.f;

        // ---- state and per-(simulation, node) gains ---------------------------------------------------------------------------
        float E[SPT], I[SPT], ad[SPT], ab[SPT], Gi[SPT], sg2[SPT];
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            const int64_t sim = sim0 + j;
            if (live) {
                if (A.init) { E[j] = c.E0; I[j] = c.I0; ab[j] = a0; ad[j] = 0.f; }
                else {
                    E[j] = A.state[(0 * (int64_t)N + node) * A.Bs + sim];
                    I[j] = A.state[(1 * (int64_t)N + node) * A.Bs + sim];
                    ab[j] = A.state[(2 * (int64_t)N + node) * A.Bs + sim];
                    ad[j] = A.state[(3 * (int64_t)N + node) * A.Bs + sim];
                }
            } else { E[j] = 0.f; I[j] = 0.f; ab[j] = 0.f; ad[j] = 0.f; }
            const float G0 = A.par[sim], dG = A.par[A.Bs + sim];
            const float sg0 = __fmul_rn(-1.4426950408889634f, A.par[2 * A.Bs + sim]), dsg = __fmul_rn(-1.4426950408889634f, A.par[3 * A.Bs + sim]);
            Gi[j] = fmaf(dG, mGn, G0);                  // homogeneous maps: fmaf(dG, 1, G0) == G0 + dG, the scalar of wc_tc.cuh's HOMO kernel
            sg2[j] = fmaf(dsg, mSn, sg0);
        }
        // streams of the simulations whose Philox calls this lane makes: lane r of a quad serves simulations r, r + 4, ...
        const int r = lane & 3, lq = lane >> 2;
        uint32_t s_lo[SPT / 4], s_hi[SPT / 4];
#pragma unroll
        for (int cc = 0; cc < SPT / 4; ++cc) {
            const uint64_t strm = A.streams[sim0 + r + 4 * cc];
            s_lo[cc] = (uint32_t)strm; s_hi[cc] = (uint32_t)(strm >> 32);
        }
        const uint32_t quad = (uint32_t)(node >> 2);
        float* Zw = Zx + warp * (SPT * 32);
        // B operand slots of this thread: element (n = g*SPT + j, k = node)
        const int b_off = (node >> 2) * (int)(LBO_B / 4) + (node & 3) + ((g * SPT) >> 3) * 32 + ((g * SPT) & 7) * 4;
        const uint64_t ad_hi = umma_desc(smem_u32(Ah), kNodeLboA, kSBO), ad_lo = umma_desc(smem_u32(Al), kNodeLboA, kSBO);
        const uint64_t bd_hi = umma_desc(smem_u32(Bh), LBO_B, kSBO), bd_lo = umma_desc(smem_u32(Bl), LBO_B, kSBO);

        int rc = A.rec_phase;
        int64_t row = A.row0;
        for (int it = 0; it < A.nsteps; ++it) {
            // 1. publish E(t) as the B operand (zeros for padding nodes: their E stays 0)
            if (live_warp) {
#pragma unroll
                for (int j = 0; j < SPT; ++j) {
                    const float v = E[j], h = tf32_rn(v);
                    const int o = b_off + ((j >> 3) * 32) + (j & 7) * 4;
                    Bh[o] = h;
                    Bl[o] = v - h;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            if (warp == ISSUER_WARP) asm volatile("bar.sync 1, %0;" ::"r"(BAR_THREADS) : "memory");
            else asm volatile("bar.arrive 1, %0;" ::"r"(BAR_THREADS) : "memory");
            if (warp == ISSUER_WARP) {
                if (elect_one()) {
                    tc_fence_after();
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        // same pass order as wc_tc.cuh: Eh.Sh, El.Sh, Eh.Sl
                        const uint64_t a0d = (pass == 2) ? ad_lo : ad_hi;
                        const uint64_t b0d = (pass == 1) ? bd_lo : bd_hi;
#pragma unroll
                        for (int kk = 0; kk < KP / 8; ++kk) {
                            umma_tf32(tmem_d, a0d + (uint64_t)(kk * ((2 * kNodeLboA) >> 4)), b0d + (uint64_t)(kk * ((2 * LBO_B) >> 4)), IDESC, acc);
                            acc = 1;
                        }
                    }
                    umma_commit(bar);
                }
                __syncwarp();
            }
            if (!live_warp) continue;                  // the dedicated issuer warp (N <= 96) has nothing else to do
            // 2. record E(t) (state BEFORE the update, netwWilsonCowanPlastic.py:129-130)
            if (A.rec) {
                if (rc == 0) {
                    if (live) {
                        float* dst = A.Ebuf + (row * N + node) * A.Bs + sim0;
#pragma unroll
                        for (int j4 = 0; j4 < SPT / 4; ++j4)
                            *reinterpret_cast<float4*>(dst + 4 * j4) = make_float4(E[4 * j4], E[4 * j4 + 1], E[4 * j4 + 2], E[4 * j4 + 3]);
                    }
                    ++row;
                }
                if (++rc == A.downsamp) rc = 0;
            }
            // 3. everything that does not need the coupling, while the tensor core works
            const uint32_t step = A.step0 + (uint32_t)it;
            if ((step & (kRecombine - 1)) == 0 && step != 0) {       // rare: fold delta into a_base
#pragma unroll
                for (int j = 0; j < SPT; ++j) { ab[j] += ad[j]; ad[j] = 0.f; }
            }
#pragma unroll
            for (int cc = 0; cc < SPT / 4; ++cc) {
                const Philox4 ph = philox4x32(step, quad, s_lo[cc], s_hi[cc], c.k0, c.k1);
                const float l0 = lg2f(u23f(ph.x)), l1 = lg2f(u23f(ph.z));
                const float a0r = fmaf(__uint_as_float(0x3f800000u | (ph.y >> 9)), two_pi, -1.49999994f * two_pi);
                const float a1r = fmaf(__uint_as_float(0x3f800000u | (ph.w >> 9)), two_pi, -1.49999994f * two_pi);
                const float r0 = sqrtaf(m2ln2 * l0), r1 = sqrtaf(m2ln2 * l1);
                const int js = r + 4 * cc;                                 // simulation (of this thread's SPT) the call belongs to
                *reinterpret_cast<float4*>(Zw + js * 32 + ((lq + 2 * js) & 7) * 4) =
                    make_float4(r0 * cosaf(a0r), r0 * sinaf(a0r), r1 * cosaf(a1r), r1 * sinaf(a1r));
            }
            __syncwarp();
            float xp[SPT];
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const float z = Zw[j * 32 + ((lq + 2 * j) & 7) * 4 + r];
                xp[j] = fmaf(c.sq, z, Pmu);
                xp[j] = fmaf(-ab[j], I[j], fmaf(-ad[j], I[j], fmaf(a_ee, E[j], xp[j])));
                const float y = fmaf(-a_ii, I[j], fmaf(a_ei, E[j], nmu));
                const float SI = rcpf(1.0f + ex2f(y * sigI2));
                ad[j] = fmaf(I[j], fmaf(E[j], A.kA, nkr), ad[j]);
                I[j] = fmaf(kI, fmaf(fmaf(-rI, I[j], 1.0f), SI, -I[j]), I[j]);
            }
            __syncwarp();                                  // the normals have been read: the tile may be rewritten next step
            // 4. coupling -> E(t+1)
            mbar_wait(bar, (uint32_t)(it & 1));
            tc_fence_after();
            uint32_t cr[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_ld_spt<SPT>(tmem_mine, cr);
            tmem_ld_wait8(cr);
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const float x = fmaf(Gi[j], __uint_as_float(cr[j]), xp[j]);
                const float SE = rcpf(1.0f + ex2f(x * sg2[j]));
                E[j] = fmaf(kE, fmaf(fmaf(-rE, E[j], 1.0f), SE, -E[j]), E[j]);
            }
            if (!live) {
#pragma unroll
                for (int j = 0; j < SPT; ++j) E[j] = 0.f;  // padding nodes never feed the contraction
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const int64_t sim = sim0 + j;
                A.state[(0 * (int64_t)N + node) * A.Bs + sim] = E[j];
                A.state[(1 * (int64_t)N + node) * A.Bs + sim] = I[j];
                A.state[(2 * (int64_t)N + node) * A.Bs + sim] = ab[j];
                A.state[(3 * (int64_t)N + node) * A.Bs + sim] = ad[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 32);
}

template <int SPT, int KQ>
static int launch_wc_node_v(const BatchArgs& A, int64_t tiles, cudaStream_t st) {
    auto kern = wc_node_kernel<SPT, KQ>;
    NREM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, node_smem_bytes<SPT, KQ>()));
    kern<<<(unsigned)tiles, kNodeThreads, node_smem_bytes<SPT, KQ>(), st>>>(A);
    NREM_LAUNCHED();
    return NREM_OK;
}

// tile_sims: 16 or 32
static int launch_wc_node(int tile_sims, const BatchArgs& A, int64_t tiles, cudaStream_t st) {
    const bool wide = A.c.N > 96;
    if (tile_sims == 16) return wide ? launch_wc_node_v<4, 4>(A, tiles, st) : launch_wc_node_v<4, 3>(A, tiles, st);
    if (tile_sims == 32) return wide ? launch_wc_node_v<8, 4>(A, tiles, st) : launch_wc_node_v<8, 3>(A, tiles, st);
    return fail(NREM_ERR_ARG, "node-lane integrator: tile_sims must be 16 or 32%s%s");
}

}  // namespace nrem
