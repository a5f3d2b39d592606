// Large-connectome integrator (BASELINE configs[4]: N = 1000 nodes, 4096 instances): one launch per Euler step.
//
// Replaces wilsonCowan()+run() (netwWilsonCowanPlastic.py:77-137) when the nodes of a simulation no longer fit the
// registers of one CTA.  Per step the coupling of ALL simulations is one GEMM
//     D[Bs sims, N nodes] = E[Bs, N] x SC^T[N, N]            (2 Bs N^2 flop: 8.2 GFLOP at Bs = 4096, N = 1000)
// tiled 128 sims x 256 nodes per CTA (two such tiles per CTA pair: cta_group::2, M = 256): a producer thread streams K-slices
// of both operands from L2 into a shared-memory ring (4 to 14 stages, by mode) with cp.async.bulk (mbarrier complete_tx), one
// thread issues tcgen05.mma (kind::tf32 / kind::f16, N = 256; split precision, see the modes below) into a 128 x 256 FP32
// accumulator in TMEM, and the epilogue warps fuse the whole node update (Philox noise, both sigmoids, E/I/a_ie Euler step,
// recording) onto the accumulator, so the coupling never touches memory.  The default mode is 5 ("bf3").
//
// HBM/L2 layout: every operand lives in global memory as the exact shared-memory image the tensor core wants (no-swizzle
// K-major canonical layout, [k/4][row][4] floats), so a pipeline stage is ONE contiguous bulk copy per operand:
//   A image  [2 (hi, lo)][tile][KG][128 sims][4]   E(t) split into TF32-exact hi and the residual lo (E = hi + lo exactly);
//                                                   two such images ping-pong between steps (the epilogue of step t writes
//                                                   E(t+1) while other CTAs still read E(t))
//   B image  [2 (hi, lo)][slice][KG][256 nodes][4]  SC rows of the slice's output nodes, staged once
//   I, a_base, a_delta                               [tile][KG][128][4], updated in place (one owner thread per element)
// KG = ceil(N/16)*4 four-node groups; padding nodes are zero in every image and are never written.
//
// Four precisions of the contraction (template MODE):
//   1 "tc"   one TF32 pass (operands truncated to TF32 by the tensor core)
//   3 "tc3"  3xTF32: Eh.Sh + El.Sh + Eh.Sl with FP32 residuals (El = E - Eh), 6 TF32 MMAs per 16 input nodes
//   4 "tcb"  TF32 main pass on the raw FP32 operands (the tensor core ignores the low 13 mantissa bits: Eh = trunc(E)) plus
//            the two correction passes  lo(E).S  and  E.lo(S)  (lo(x) = x - trunc(x) ~ 2^-10 x) as kind::f16 BF16 MMAs at twice
//            the TF32 rate: BF16's 2^-9 rounding on a 2^-10 term leaves ~2^-19 ~ 2e-6 of the coupling, below the ~1e-5
//            accumulation bias of the tensor core itself at N = 1000.  2 TF32 + 2 BF16 MMAs per 16 input nodes (-33 % time).
//            Images: A = [F: float4 planes][L: bf16 lo(E)][H: bf16 E] with bf16 planes as [tile][KG/2][128 sims][8] (16 B rows),
//            B likewise [F][H: bf16 S][L: bf16 lo(S)].
//   5 "bf3"  3xBF16: x = h + l with h = bf16(x), l = bf16(x - h) (16 mantissa bits, |x - h - l| <= 2^-18 |x|);
//            D = Eh.Sh + El.Sh + Eh.Sl as three kind::f16 MMAs (K = 16) per 16 input nodes = 1.5 TF32-pass equivalents instead of
//            tcb's 2, and HALF of tcb's operand bytes per K step (4 B per element instead of 8: the FP32 plane no longer goes
//            through shared memory), which is what the ring-bound MMA phase responds to.  The dropped El.Sl term and the
//            residuals are ~4e-6 per product with random sign (they average in the N-term sum): below the tensor core's own
//            accumulation bias.  Same images as tcb: F = E in FP32 (state only, never staged), H, L.  Two footprint cuts
//            (the 124 MB of state + operands of configs[4] no longer fit the L2; ncu: half of the sectors missed): the F
//            plane is private to its owner thread, so it is updated IN PLACE (plane F of image 0; only H and L ping-pong),
//            and a_base is kept as bf16 (2 B; a_ie = a_base + delta with delta = the exact FP32 remainder, re-split every
//            kRecombine steps), which takes the working set to 93 MB.
#pragma once
#include <cuda_bf16.h>

#include "wc_tc.cuh"

namespace nrem {

constexpr int kBigNT = 256;                 // output nodes per CTA (= MMA N)
constexpr int kBigKS = 4;                   // four-node groups per pipeline stage (16 input nodes, two K = 8 MMAs)
constexpr int kBigStages = 4;
constexpr int kBigEpiWarps = 16;
constexpr int kBigThreads = (2 + kBigEpiWarps) * 32;
constexpr int kBigCols = kBigNT / (kBigEpiWarps / 4);     // accumulator columns per epilogue warp
// bf3 can take more input nodes per stage (its stages are half as big per node): NREM_BIG_KS_BF3 four-node groups
#ifndef NREM_BIG_KS_BF3
#define NREM_BIG_KS_BF3 4
#endif
template <int MODE> __host__ __device__ constexpr int big_ks() { return MODE == 5 ? NREM_BIG_KS_BF3 : kBigKS; }
constexpr uint32_t kBigAStage = kBigKS * kTile * 16;     // 8 KB per (hi | lo)
constexpr uint32_t kBigBStage = kBigKS * kBigNT * 16;    // 16 KB per (hi | lo)
constexpr uint32_t kBigLBO_A = kTile * 16;
constexpr uint32_t kBigLBO_B = kBigNT * 16;
constexpr uint32_t kBigTmemCols = 512;              // 0..255 coupling accumulator, 256..511 the coupling-free part of the sigmoid argument
constexpr uint32_t kBigIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBigNT >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

// PAIR (cta_group::2): two CTAs of a cluster (two 128-simulation tiles, same node slice) execute ONE M = 256 MMA per K slice; each
// CTA stages its own A tile and only HALF of the B tile (128 of the 256 output nodes), the tensor cores of the pair exchange the
// halves.  Per SM and 16 input nodes that is 32 KB instead of 48 KB through the bulk-copy ring and shared memory, which buys 
// 7-stage ring in the same shared memory.
#ifndef NREM_BIG_STAGES_PAIR
#define NREM_BIG_STAGES_PAIR 7
#endif
constexpr int kBigStagesPair = NREM_BIG_STAGES_PAIR;
template <int MODE, bool PAIR = false>
__host__ __device__ constexpr uint32_t big_stage_bytes() {
    return ((MODE == 1 || MODE == 5) ? 1u : 2u) * (uint32_t)(big_ks<MODE>() / kBigKS) * (kBigAStage + (PAIR ? kBigBStage / 2 : kBigBStage));
}
// ring depth: bf3 stages are half as big as tcb's, so twice as many fit the same shared memory
#ifndef NREM_BIG_STAGES_BF3_PAIR
#define NREM_BIG_STAGES_BF3_PAIR 14
#endif
template <int MODE, bool PAIR = false>
__host__ __device__ constexpr int big_stages() {
    return MODE == 5 ? (PAIR ? NREM_BIG_STAGES_BF3_PAIR : 9) * kBigKS / big_ks<5>() : (PAIR ? kBigStagesPair : kBigStages);
}
template <int MODE, bool PAIR = false>
constexpr int big_smem_bytes() { return (int)(big_stages<MODE, PAIR>() * big_stage_bytes<MODE, PAIR>()) + 512; }
constexpr uint32_t kBigIdescPair = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBigNT >> 3) << 17) | ((uint32_t)((2 * kTile) >> 4) << 24);
constexpr uint32_t kBigIdescBf16Pair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBigNT >> 3) << 17) | ((uint32_t)((2 * kTile) >> 4) << 24);
// kind::f16 instruction descriptor: D = F32, A = B = BF16 (format 1), K-major, N = 256, M = 128
constexpr uint32_t kBigIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBigNT >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- cta_group::2 (CTA pair) forms ----
__device__ __forceinline__ void umma_tf32_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.  RELAXED: the arriving thread publishes no
// data of its own — the stage was written into this CTA's shared memory by the bulk-copy engine (whose completion this thread has
// observed on the local mbarrier) and is read in place by the pair's tensor cores, like a cta_group::2 TMA that signals the leader's
// barrier directly.  (A .release.cluster arrive compiles to MEMBAR.ALL.GPU per stage and made the step 60 % slower.)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t bf16x2(float lo_elem, float hi_elem) {       // {lo_elem in bits 0..15, hi_elem in 16..31}, RN
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

// rows of the per-node table of the NODEPAR kernels ("Any of them can be redefined as a vector of length nnodes", netwWilsonCowanPlastic.py:21)
enum { kNpAee = 0, kNpAei, kNpAii, kNpKE, kNpKI, kNpPmu, kNpRhoE, kNpRE, kNpRI, kNpNmu, kNpSigI2, kNpA0, kBigNpar };
__device__ __forceinline__ float f4c(const float4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

struct BigArgs {
    BatchConst c;
    const float4* Acur;        // E(t) image
    float4* Anext;             // E(t+1) image
    const float4* Bimg;
    float4* I4;
    float4* ab4;               // a_ie base (bf3: bf16, addressed as uint2[])
    float4* Fst;               // bf3: the FP32 plane of E, updated in place by its owner thread (plane F of image 0)
    const float4* npar;        // NODEPAR kernels: per-node parameter table [kBigNpar][Kpad] floats (big_stage_npar_kernel), else NULL
    float4* ad4;               // a_ie delta (a_ie = base + delta, see wc_tc.cuh)
    const float* par;          // [4][Bs]: G0, dG, sigma0, dsigm
    const uint64_t* streams;   // [Bs]
    const float* mapG;         // [4*KG]
    const float* mapS;
    int64_t Bs;                // simulations incl. padding tiles (a multiple of 128; of 256 for the CTA-pair kernel)
    int64_t Bo;                // simulation stride of the OUTPUT arrays Ebuf / coup (B rounded up to 128, include/nremfc.h); Bo <= Bs
    int tiles, slices, KG;
    int homo;                  // maps are all ones
    uint32_t step;             // global Euler step index
    float kA;                  // dtSim / tau_ip of this phase
    int recombine;             // fold a_delta into a_base before this step (global steps that are multiples of kRecombine)
    int rec;                   // store E(t) into Ebuf row `row`
    int64_t row;
    float* Ebuf;               // [rows][N][Bs]
    float* coup;               // optional [N][Bs]: the coupling SC.E(t) of this step (tests)
    // persistent (cluster) mode: the launch runs `nsteps` Euler steps starting at global step `step`; the per-step quantities above
    // (kA, recombine, rec, row, coup, and which image is E(t)) are then derived in the kernel from the step index
    int nsteps;
    uint32_t n1, n12;          // phase boundaries: steps [0, n1) phase 1, [n1, n12) phase 2, the rest phase 3 (recording)
    float kA3[3];              // dtSim / tau_ip per phase
    int downsamp;
    float4* img[2];            // the two A images; E(t) of global step s is img[s & 1]
    unsigned long long* dbg;   // NULL, or [CTAs][8] %globaltimer stamps of this launch (NREM_BIG_DBG: phase timing of one step)
};

__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define NREM_BIG_STAMP(slot) do { if (A.dbg && lane == 0) A.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime_ns(); } while (0)

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct BigStep {               // per-step quantities of the persistent mode
    const float4* Acur;
    float4* Anext;
    uint32_t step;
    float kA;
    int recombine, rec;
    int64_t row;
    float* coup;
};
__device__ __forceinline__ BigStep big_step_of(const BigArgs& A, int it, bool persist) {
    BigStep S;
    if (!persist) {
        S.Acur = A.Acur; S.Anext = A.Anext; S.step = A.step; S.kA = A.kA; S.recombine = A.recombine; S.rec = A.rec; S.row = A.row; S.coup = A.coup;
        return S;
    }
    const uint32_t s = A.step + (uint32_t)it;
    S.step = s;
    S.Acur = A.img[s & 1]; S.Anext = A.img[(s + 1) & 1];
    const int ph = s < A.n1 ? 0 : (s < A.n12 ? 1 : 2);
    S.kA = A.kA3[ph];
    S.recombine = (s != 0 && (s & (kRecombine - 1)) == 0) ? 1 : 0;
    const uint32_t r = s - A.n12;
    S.rec = (A.Ebuf && ph == 2 && r % (uint32_t)A.downsamp == 0) ? 1 : 0;
    S.row = S.rec ? (int64_t)(r / (uint32_t)A.downsamp) : 0;
    S.coup = s == 0 ? A.coup : nullptr;
    return S;
}


// wait for two 8-column loads; the registers are in/out operands so that no use can be hoisted above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[8], uint32_t (&q)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7])
                 :: "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

struct BigQuad {             // state of one thread's four-node group
    float4 eh, el, i, b, d;
};

// FULL: N is a multiple of 8, so no node of a processed 8-node group is padding.  HOMO: no per-node maps (G, sigma per simulation)
// PERSIST: launched as thread-block clusters of `slices` CTAs (the node slices of one 128-simulation tile, which depend only on each
//          other); the cluster runs A.nsteps Euler steps with one barrier.cluster per step instead of one launch per step.
// PAIR   : launched as clusters (2, 1, 1) on a (tiles, slices) grid: tiles 2p and 2p+1 of a node slice form a CTA pair (cta_group::2).  Rank 0 issues the
//          M = 256 MMAs for both; every CTA streams its own A tile and its half of the B tile; the peer relays "my stage has
//          landed" to the leader's full barrier; tcgen05.commit multicasts the stage-free / accumulator-ready arrivals to both.
// NODEPAR: every node parameter from a per-node table (A.npar) instead of the scalars of the launch (bf3, one launch per step).
template <int MODE, bool FULL, bool HOMO, bool PERSIST, bool PAIR = false, bool NODEPAR = false>
__global__ void __launch_bounds__(kBigThreads, 1) wc_big_step_kernel(const BigArgs A) {
    extern __shared__ __align__(128) unsigned char smraw[];
    static_assert(!(PAIR && PERSIST), "the CTA-pair kernel is launched once per Euler step");
    static_assert(!NODEPAR || (MODE == 5 && !PERSIST && !HOMO), "per-node parameter tables: bf3, one launch per step, map kernel");
    constexpr bool SPLIT = MODE == 3;
    constexpr bool BF3 = MODE == 5;
    constexpr bool MIXED = MODE == 4 || BF3;        // state layout: E in one FP32 plane + two bf16 planes
    constexpr int NST = big_stages<MODE, PAIR>();
    constexpr int BROWS = PAIR ? kBigNT / 2 : kBigNT;              // B rows (output nodes) this CTA holds in shared memory
    constexpr int KS = big_ks<MODE>();                             // four-node groups per pipeline stage
    constexpr uint32_t ASTAGE = (uint32_t)KS * kTile * 16;          // bytes of one FP32 A stage (bf16 parts: half)
    constexpr uint32_t BSTAGE = (uint32_t)KS * BROWS * 16;         // bytes of one FP32 B stage
    constexpr uint32_t LBO_B = (uint32_t)BROWS * 16;
    constexpr uint32_t STAGE = big_stage_bytes<MODE, PAIR>();
    static_assert(STAGE == ((MODE == 1 || BF3) ? 1u : 2u) * (ASTAGE + BSTAGE), "stage size");
    constexpr uint32_t IDESC_TF32 = PAIR ? kBigIdescPair : kBigIdesc, IDESC_BF16 = PAIR ? kBigIdescBf16Pair : kBigIdescBf16;
    // stage layout   MODE 1: [A 8K][B 16K]      MODE 3: [Ah 8K][Al 8K][Bh 16K][Bl 16K]
    //                MODE 4: [Af 8K][Al bf16 4K][Ah bf16 4K][Bf 16K][Bh bf16 8K][Bl bf16 8K]        (PAIR: every B part is half as big)
    //                MODE 5: [Al bf16 4K][Ah bf16 4K][Bh bf16 8K][Bl bf16 8K]
    constexpr uint32_t OFF_B = ((MODE == 1 || BF3) ? 1u : 2u) * ASTAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw + NST * STAGE);
    uint64_t* empty = full + NST;
    uint64_t* accum = empty + NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;           // 0: leader (issues the MMAs of the pair)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // grid (slices, tiles); the CTA-pair kernel is launched as (tiles, slices) with clusters (2, 1, 1) = tiles 2p, 2p+1 of a slice
    const int slice = PAIR ? blockIdx.y : blockIdx.x, tile = PAIR ? blockIdx.x : blockIdx.y;
    const int KT = A.KG / KS;
    const BatchConst& c = A.c;
    const int N = c.N;

    const int nsteps = PERSIST ? A.nsteps : 1;
    if (warp == 2) NREM_BIG_STAMP(0);                    // CTA start
    // programmatic dependent launch: let the next step's grid start its prologue as soon as SMs free up ...
    if (!PERSIST) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        // leader of a pair: a stage is full when its own copies have landed AND the peer has reported its copies
        for (int s = 0; s < NST; ++s) { mbar_init(full + s, (PAIR && rank == 0) ? 2 : 1); mbar_init(empty + s, 1); }
        mbar_init(accum, 1);
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_2(tmem_slot, kBigTmemCols); else tmem_alloc(tmem_slot, kBigTmemCols); }
    tc_fence_before();
    if (PAIR) { cluster_arrive_relaxed(); cluster_wait(); }      // the peer's barriers (fence.mbarrier_init above) and TMEM are set up, too
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const size_t plane = (size_t)A.tiles * A.KG * kTile;                 // float4 per FP32 plane of an A image
    const size_t planeB = (size_t)A.slices * A.KG * kBigNT;              // float4 per FP32 plane of the B image
    // ... and wait here until the previous step's grid has completed and its E(t), I, a_ie are visible
    if (!PERSIST) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 2) NREM_BIG_STAMP(1);                    // previous step complete and visible

    if (warp == 0) {
      uint32_t ring = 0;                                  // stages issued so far (continues across the steps of a persistent launch)
      for (int it = 0; it < nsteps; ++it) {
        if (elect_one()) {      // (elect.sync under a warp-uniform branch: descriptors/addresses go to uniform registers once)
            // ---- producer: one contiguous bulk copy per operand and stage ----
            const float4* Acur = big_step_of(A, it, PERSIST).Acur;
            const char* a0 = reinterpret_cast<const char*>(Acur + (size_t)tile * A.KG * kTile);
            const char* a1 = reinterpret_cast<const char*>(Acur + plane + (size_t)tile * A.KG * kTile);              // MODE 3: lo plane
            const size_t blk = PAIR ? (size_t)slice * 2 + rank : (size_t)slice;          // B block: [slice] or [slice][half of the nodes]
            const char* b0 = reinterpret_cast<const char*>(A.Bimg + blk * A.KG * BROWS);
            const char* b1 = reinterpret_cast<const char*>(A.Bimg + planeB + blk * A.KG * BROWS);
            // MODE 4: bf16 planes, 16-byte rows of 8 nodes: [tile][KG/2][128] and [slice][KG/2][256] uint4
            const char* aL = reinterpret_cast<const char*>(Acur + plane) + (size_t)tile * (A.KG / 2) * kTile * 16;
            const char* aH = aL + plane * 8;
            const char* bH = reinterpret_cast<const char*>(A.Bimg + planeB) + blk * (A.KG / 2) * BROWS * 16;
            const char* bL = bH + planeB * 8;
            const uint32_t base = smem_u32(smraw);
            for (int kt = 0; kt < KT; ++kt, ++ring) {
                const int s = (int)(ring % NST);
                mbar_wait(empty + s, (uint32_t)(((ring / NST) & 1) ^ 1));
                mbar_expect_tx(full + s, STAGE);
                const uint32_t dst = base + (uint32_t)s * STAGE;
                if (BF3) {
                    bulk_g2s(dst, aL + (size_t)kt * (ASTAGE / 2), ASTAGE / 2, full + s);
                    bulk_g2s(dst + ASTAGE / 2, aH + (size_t)kt * (ASTAGE / 2), ASTAGE / 2, full + s);
                    bulk_g2s(dst + OFF_B, bH + (size_t)kt * (BSTAGE / 2), BSTAGE / 2, full + s);
                    bulk_g2s(dst + OFF_B + BSTAGE / 2, bL + (size_t)kt * (BSTAGE / 2), BSTAGE / 2, full + s);
                    continue;
                }
                bulk_g2s(dst, a0 + (size_t)kt * ASTAGE, ASTAGE, full + s);
                bulk_g2s(dst + OFF_B, b0 + (size_t)kt * BSTAGE, BSTAGE, full + s);
                if (SPLIT) {
                    bulk_g2s(dst + ASTAGE, a1 + (size_t)kt * ASTAGE, ASTAGE, full + s);
                    bulk_g2s(dst + OFF_B + BSTAGE, b1 + (size_t)kt * BSTAGE, BSTAGE, full + s);
                }
                if (MODE == 4) {
                    bulk_g2s(dst + ASTAGE, aL + (size_t)kt * (ASTAGE / 2), ASTAGE / 2, full + s);
                    bulk_g2s(dst + ASTAGE + ASTAGE / 2, aH + (size_t)kt * (ASTAGE / 2), ASTAGE / 2, full + s);
                    bulk_g2s(dst + OFF_B + BSTAGE, bH + (size_t)kt * (BSTAGE / 2), BSTAGE / 2, full + s);
                    bulk_g2s(dst + OFF_B + BSTAGE + BSTAGE / 2, bL + (size_t)kt * (BSTAGE / 2), BSTAGE / 2, full + s);
                }
            }
        }
        __syncwarp();
        if (PERSIST) { cluster_arrive_release(); cluster_wait_acquire(); }      // E(t+1) of all node slices is written and visible
      }
    } else if (warp == 1) {
      uint32_t ring = 0;
      for (int it = 0; it < nsteps; ++it) {
        if (PAIR && rank != 0) {
            // ---- peer of a CTA pair: no MMAs to issue; tell the leader when each of my stages has landed ----
            if (elect_one()) {
                for (int kt = 0; kt < KT; ++kt, ++ring) {
                    const int s = (int)(ring % NST);
                    mbar_wait(full + s, (uint32_t)((ring / NST) & 1));
                    mbar_arrive_remote(full + s, 0);
                }
            }
        } else if (elect_one()) {
            // ---- MMA issuer ----
            const uint32_t base = smem_u32(smraw);
            uint32_t acc = 0;
            tc_fence_after();                              // (persistent: the epilogue's TMEM reads of the previous step are done)
            for (int kt = 0; kt < KT; ++kt, ++ring) {
                const int s = (int)(ring % NST);
                mbar_wait(full + s, (uint32_t)((ring / NST) & 1));
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)s * STAGE;
                if (BF3) {
                    const uint64_t aL = umma_desc(sa, kBigLBO_A, kSBO), aH = umma_desc(sa + ASTAGE / 2, kBigLBO_A, kSBO);
                    const uint64_t bH = umma_desc(sa + OFF_B, LBO_B, kSBO), bL = umma_desc(sa + OFF_B + BSTAGE / 2, LBO_B, kSBO);
#pragma unroll
                    for (int k16 = 0; k16 < KS / 4; ++k16) {
                        const uint64_t da = (uint64_t)(k16 * ((2 * kBigLBO_A) >> 4)), db = (uint64_t)(k16 * ((2 * LBO_B) >> 4));
                        if (PAIR) { umma_bf16_2(tmem_d, aH + da, bH + db, IDESC_BF16, acc); umma_bf16_2(tmem_d, aL + da, bH + db, IDESC_BF16, 1); umma_bf16_2(tmem_d, aH + da, bL + db, IDESC_BF16, 1); }
                        else { umma_bf16(tmem_d, aH + da, bH + db, IDESC_BF16, acc); umma_bf16(tmem_d, aL + da, bH + db, IDESC_BF16, 1); umma_bf16(tmem_d, aH + da, bL + db, IDESC_BF16, 1); }
                        acc = 1;
                    }
                    if (PAIR) umma_commit_2(empty + s); else umma_commit(empty + s);
                    continue;
                }
                const uint64_t ad_hi = umma_desc(sa, kBigLBO_A, kSBO), ad_lo = umma_desc(sa + ASTAGE, kBigLBO_A, kSBO);
                const uint64_t bd_hi = umma_desc(sa + OFF_B, LBO_B, kSBO), bd_lo = umma_desc(sa + OFF_B + BSTAGE, LBO_B, kSBO);
#pragma unroll
                for (int k8 = 0; k8 < KS / 2; ++k8) {
#pragma unroll
                    for (int pass = 0; pass < (SPLIT ? 3 : 1); ++pass) {
                        const uint64_t a0 = (pass == 1) ? ad_lo : ad_hi;
                        const uint64_t b0 = (pass == 2) ? bd_lo : bd_hi;
                        const uint64_t ak = a0 + (uint64_t)(k8 * ((2 * kBigLBO_A) >> 4)), bk = b0 + (uint64_t)(k8 * ((2 * LBO_B) >> 4));
                        if (PAIR) umma_tf32_2(tmem_d, ak, bk, IDESC_TF32, acc); else umma_tf32(tmem_d, ak, bk, IDESC_TF32, acc);
                        acc = 1;
                    }
                }
                if (MODE == 4) {
                    // one K = 16 BF16 MMA per correction: the bf16 stage arrays are [2 eight-node groups][rows][16 B], i.e. the same
                    // core-matrix geometry (LBO = rows x 16 B, SBO = 128 B) as a K = 8 TF32 slice
                    const uint64_t aL = umma_desc(sa + ASTAGE, kBigLBO_A, kSBO), aH = umma_desc(sa + ASTAGE + ASTAGE / 2, kBigLBO_A, kSBO);
                    const uint64_t bH = umma_desc(sa + OFF_B + BSTAGE, LBO_B, kSBO), bL = umma_desc(sa + OFF_B + BSTAGE + BSTAGE / 2, LBO_B, kSBO);
#pragma unroll
                    for (int k16 = 0; k16 < KS / 4; ++k16) {
                        const uint64_t da = (uint64_t)(k16 * ((2 * kBigLBO_A) >> 4)), db = (uint64_t)(k16 * ((2 * LBO_B) >> 4));
                        if (PAIR) { umma_bf16_2(tmem_d, aL + da, bH + db, IDESC_BF16, 1); umma_bf16_2(tmem_d, aH + da, bL + db, IDESC_BF16, 1); }
                        else { umma_bf16(tmem_d, aL + da, bH + db, IDESC_BF16, 1); umma_bf16(tmem_d, aH + da, bL + db, IDESC_BF16, 1); }
                    }
                }
                // the stage may be refilled (in both CTAs of a pair) once these MMAs have read it
                if (PAIR) umma_commit_2(empty + s); else umma_commit(empty + s);
            }
            if (PAIR) umma_commit_2(accum); else umma_commit(accum);
            NREM_BIG_STAMP(2);                           // last MMA issued
        }
        __syncwarp();
        if (PERSIST) { cluster_arrive_release(); cluster_wait_acquire(); }
      }
    } else {
        // ---- epilogue: the node update of (128 sims) x (256 nodes), fused onto the TMEM accumulator ----
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const int cgp = (warp - 2) >> 2;                      // column group
        const int r = q * 32 + lane;
        const int64_t sim = (int64_t)tile * kTile + r;
        const uint32_t tmem_mine = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(cgp * kBigCols);
        const int node_base = slice * kBigNT + cgp * kBigCols;
        const size_t rowbase = (size_t)tile * A.KG * kTile + r;
        const float G0 = A.par[sim], dG = A.par[A.Bs + sim];
        const float sg0 = __fmul_rn(-1.4426950408889634f, A.par[2 * A.Bs + sim]), dsg = __fmul_rn(-1.4426950408889634f, A.par[3 * A.Bs + sim]);
        const uint64_t strm = A.streams[sim];
        const uint32_t s_lo = (uint32_t)strm, s_hi = (uint32_t)(strm >> 32);
        const float Pmu = c.P - c.mu, nmu = -c.mu;
        const float Gh = __fadd_rn(G0, dG), sgh = __fadd_rn(sg0, dsg);
        int ng = (N - node_base + 7) / 8;
        ng = ng < 0 ? 0 : (ng > kBigCols / 8 ? kBigCols / 8 : ng);

      for (int it = 0; it < nsteps; ++it) {
        const BigStep S = big_step_of(A, it, PERSIST);
        const float nkr = -S.kA * c.rhoE;
        // -- phase 1, while the tensor core works: everything that does not need the coupling (noise, I and a_ie updates,
        //    recording) and xp = a_ee E - a_ie I + P - mu + noise, parked in the spare TMEM columns 256..511.
        //    The state of quad qd + 1 is requested before quad qd is computed (two quads ahead measured 2 % slower in the
        //    MMA-bound modes: register pressure).
        {
            const int nq = 2 * ng;
            const size_t idx0 = rowbase + (size_t)(node_base >> 2) * kTile;
            auto load4 = [&](int qd, BigQuad& s) {
                const size_t idx = idx0 + (size_t)qd * kTile;
                s.eh = BF3 ? A.Fst[idx] : S.Acur[idx];
                if (!MIXED) s.el = S.Acur[plane + idx];
                s.i = A.I4[idx];
                if (BF3) {               // bf16 x 4 -> FP32
                    const uint2 u = reinterpret_cast<const uint2*>(A.ab4)[idx];
                    s.b = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
                } else s.b = A.ab4[idx];
                s.d = A.ad4[idx];
            };
            BigQuad cur, n1;
            if (nq > 0) load4(0, cur);
            for (int qd = 0; qd < nq; ++qd) {
                if (qd + 1 < nq) load4(qd + 1, n1);
                const int node0 = node_base + 4 * qd;
                const size_t idx = idx0 + (size_t)qd * kTile;
                uint32_t xp4[4];
                float z[4];
                normals4f(philox4x32(S.step, (uint32_t)(node0 >> 2), s_lo, s_hi, c.k0, c.k1), z[0], z[1], z[2], z[3]);
                float E[4] = {cur.eh.x, cur.eh.y, cur.eh.z, cur.eh.w};
                if (!MIXED) { E[0] += cur.el.x; E[1] += cur.el.y; E[2] += cur.el.z; E[3] += cur.el.w; }
                float I[4] = {cur.i.x, cur.i.y, cur.i.z, cur.i.w};
                float ab[4] = {cur.b.x, cur.b.y, cur.b.z, cur.b.w};
                float ad[4] = {cur.d.x, cur.d.y, cur.d.z, cur.d.w};
                // per-node table: one broadcast 16-byte load per parameter and quad (the node index is warp-uniform)
                float4 t_aee, t_aei, t_aii, t_kI, t_pmu, t_rho, t_rI, t_nmu, t_sI;
                if (NODEPAR) {
                    const float4* T = A.npar + (node0 >> 2);
                    const int KQ = A.KG;
                    t_aee = __ldg(T + kNpAee * KQ); t_aei = __ldg(T + kNpAei * KQ); t_aii = __ldg(T + kNpAii * KQ); t_kI = __ldg(T + kNpKI * KQ);
                    t_pmu = __ldg(T + kNpPmu * KQ); t_rho = __ldg(T + kNpRhoE * KQ); t_rI = __ldg(T + kNpRI * KQ); t_nmu = __ldg(T + kNpNmu * KQ);
                    t_sI = __ldg(T + kNpSigI2 * KQ);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int node = node0 + j;
                    const bool live = FULL || node < N;
                    if (S.recombine) {
                        ab[j] += ad[j];
                        if (BF3) { const float t = ab[j]; ab[j] = __uint_as_float(bf16x2(0.f, t) & 0xFFFF0000u); ad[j] = t - ab[j]; }      // bf16 base + exact remainder
                        else ad[j] = 0.f;
                    }
                    if (live && S.rec && sim < A.Bo) A.Ebuf[((size_t)S.row * N + node) * A.Bo + sim] = E[j];      // state BEFORE the update (WC:129-130)
                    const float p_aee = NODEPAR ? f4c(t_aee, j) : c.a_ee, p_aei = NODEPAR ? f4c(t_aei, j) : c.a_ei, p_aii = NODEPAR ? f4c(t_aii, j) : c.a_ii;
                    const float p_pmu = NODEPAR ? f4c(t_pmu, j) : Pmu, p_nmu = NODEPAR ? f4c(t_nmu, j) : nmu, p_sI = NODEPAR ? f4c(t_sI, j) : c.sigI2;
                    const float p_kI = NODEPAR ? f4c(t_kI, j) : c.kI, p_rI = NODEPAR ? f4c(t_rI, j) : c.rI;
                    const float p_nkr = NODEPAR ? -S.kA * f4c(t_rho, j) : nkr;
                    float xp = fmaf(c.sq, z[j], p_pmu);
                    xp = fmaf(-ab[j], I[j], fmaf(-ad[j], I[j], fmaf(p_aee, E[j], xp)));
                    xp4[j] = __float_as_uint(xp);
                    const float y = fmaf(-p_aii, I[j], fmaf(p_aei, E[j], p_nmu));
                    const float SI = rcpf(1.0f + ex2f(y * p_sI));
                    const float dn = fmaf(I[j], fmaf(E[j], S.kA, p_nkr), ad[j]);
                    const float In = fmaf(p_kI, fmaf(fmaf(-p_rI, I[j], 1.0f), SI, -I[j]), I[j]);
                    I[j] = live ? In : 0.f;
                    ad[j] = live ? dn : 0.f;
                }
                A.I4[idx] = make_float4(I[0], I[1], I[2], I[3]);
                A.ad4[idx] = make_float4(ad[0], ad[1], ad[2], ad[3]);
                if (S.recombine) {
                    if (BF3) reinterpret_cast<uint2*>(A.ab4)[idx] = make_uint2((__float_as_uint(ab[0]) >> 16) | (__float_as_uint(ab[1]) & 0xFFFF0000u),
                                                                               (__float_as_uint(ab[2]) >> 16) | (__float_as_uint(ab[3]) & 0xFFFF0000u));
                    else A.ab4[idx] = make_float4(ab[0], ab[1], ab[2], ab[3]);
                }
                tmem_st4(tmem_mine + kBigNT + 4 * qd, xp4);
                cur = n1;
            }
            tmem_st_wait();
        }
        if (warp == 2) NREM_BIG_STAMP(3);                // phase 1 done (warp 2)
        // -- phase 2, after the last MMA: x = xp + G coup -> E(t+1), stored as the next A image
        {
            constexpr int NE = MIXED ? 2 : 4;
            float4 ce[NE], ne[NE], n2[NE];       // MODE 1/3: eh[0], eh[1], el[0], el[1];  MODE 4: E[0], E[1]
            constexpr bool AHEAD2 = MIXED;       // E(t) of group g + 2 requested while group g is computed (8 registers in MODE 4): an L2
                                                 // round trip is longer than one group's arithmetic
            auto loadE = [&](int g, float4 (&e)[NE]) {
                const size_t idx = rowbase + (size_t)((node_base + 8 * g) >> 2) * kTile;
                if (BF3) { e[0] = A.Fst[idx]; e[1] = A.Fst[idx + kTile]; }
                else { e[0] = S.Acur[idx]; e[1] = S.Acur[idx + kTile]; }
                if (!MIXED) { e[NE - 2] = S.Acur[plane + idx]; e[NE - 1] = S.Acur[plane + idx + kTile]; }
            };
            // software pipeline: the TMEM loads (coupling + xp) and the E(t) loads of group g + 1 are in flight while group g is computed
            uint32_t cr[8], xp8[8], crn[8], xpn[8];
            if (ng > 0) { loadE(0, ce); tmem_ld8(tmem_mine + kBigNT, xp8); }          // xp does not depend on the MMAs
            if (AHEAD2 && ng > 1) loadE(1, ne);
            mbar_wait(accum, (uint32_t)(it & 1));
            tc_fence_after();
            if (warp == 2) NREM_BIG_STAMP(4);            // accumulator complete
            if (ng > 0) { tmem_ld8(tmem_mine, cr); tmem_ld_wait16(cr, xp8); }
            for (int g = 0; g < ng; ++g) {
                if (g + 1 < ng) {
                    if (!AHEAD2) loadE(g + 1, ne);
                    else if (g + 2 < ng) loadE(g + 2, n2);
                    tmem_ld8(tmem_mine + 8 * (g + 1), crn);
                    tmem_ld8(tmem_mine + kBigNT + 8 * (g + 1), xpn);
                }
                const int node0 = node_base + 8 * g;
                const size_t idx = rowbase + (size_t)(node0 >> 2) * kTile;
                if (S.coup && sim < A.Bo) {      // test hook (first step only)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (node0 + j < N) S.coup[(size_t)(node0 + j) * A.Bo + sim] = __uint_as_float(cr[j]);
                }
                float En[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float E[4] = {ce[h].x, ce[h].y, ce[h].z, ce[h].w};
                    if (!MIXED) { E[0] += ce[NE - 2 + h].x; E[1] += ce[NE - 2 + h].y; E[2] += ce[NE - 2 + h].z; E[3] += ce[NE - 2 + h].w; }
                    float4 t_kE, t_rE;
                    if (NODEPAR) {
                        const float4* T = A.npar + ((node0 + 4 * h) >> 2);
                        t_kE = __ldg(T + kNpKE * A.KG); t_rE = __ldg(T + kNpRE * A.KG);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int node = node0 + 4 * h + j;
                        const bool live = FULL || node < N;
                        const float coup = __uint_as_float(cr[4 * h + j]);
                        const float Gi = HOMO ? Gh : fmaf(dG, __ldg(A.mapG + node), G0);
                        const float sg2 = HOMO ? sgh : fmaf(dsg, __ldg(A.mapS + node), sg0);
                        const float x = fmaf(Gi, coup, __uint_as_float(xp8[4 * h + j]));
                        const float SE = rcpf(1.0f + ex2f(x * sg2));
                        const float p_kE = NODEPAR ? f4c(t_kE, j) : c.kE, p_rE = NODEPAR ? f4c(t_rE, j) : c.rE;
                        const float e1 = fmaf(p_kE, fmaf(fmaf(-p_rE, E[j], 1.0f), SE, -E[j]), E[j]);
                        En[4 * h + j] = live ? e1 : 0.f;
                    }
                }
                if (MIXED) {
                    float4* Fn = BF3 ? A.Fst : S.Anext;
                    Fn[idx] = make_float4(En[0], En[1], En[2], En[3]);
                    Fn[idx + kTile] = make_float4(En[4], En[5], En[6], En[7]);
                    float lo[8];
                    const uint4 h8 = make_uint4(bf16x2(En[0], En[1]), bf16x2(En[2], En[3]), bf16x2(En[4], En[5]), bf16x2(En[6], En[7]));
                    if (BF3) {           // residual of the bf16 rounding (exact in FP32)
                        const uint32_t hw[4] = {h8.x, h8.y, h8.z, h8.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            lo[2 * j] = En[2 * j] - __uint_as_float(hw[j] << 16);
                            lo[2 * j + 1] = En[2 * j + 1] - __uint_as_float(hw[j] & 0xFFFF0000u);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) lo[j] = En[j] - tf32_trunc(En[j]);
                    }
                    uint4* L = reinterpret_cast<uint4*>(S.Anext + plane);
                    uint4* H = L + plane / 2;
                    const size_t i8 = ((size_t)tile * (A.KG / 2) + (size_t)(node0 >> 3)) * kTile + r;
                    L[i8] = make_uint4(bf16x2(lo[0], lo[1]), bf16x2(lo[2], lo[3]), bf16x2(lo[4], lo[5]), bf16x2(lo[6], lo[7]));
                    H[i8] = h8;
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float hi[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) hi[j] = tf32_rn(En[4 * h + j]);
                        S.Anext[idx + h * kTile] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        S.Anext[plane + idx + h * kTile] = make_float4(En[4 * h] - hi[0], En[4 * h + 1] - hi[1], En[4 * h + 2] - hi[2], En[4 * h + 3] - hi[3]);
                    }
                }
                if (g + 1 < ng) tmem_ld_wait16(crn, xpn);
#pragma unroll
                for (int k = 0; k < NE; ++k) { ce[k] = ne[k]; if (AHEAD2) ne[k] = n2[k]; }
#pragma unroll
                for (int k = 0; k < 8; ++k) { cr[k] = crn[k]; xp8[k] = xpn[k]; }
            }
        }
            if (warp == 2) NREM_BIG_STAMP(5);            // phase 2 done (warp 2)
            if (warp == 17) NREM_BIG_STAMP(6);           // phase 2 done (last epilogue warp)
            if (PERSIST) {
            // E(t+1), I, a_ie of this CTA's node slice are written: make them visible to the other slices' bulk copies (async proxy) ...
            if (A.nsteps < 0) __threadfence();       // (never: barrier.cluster release/acquire already orders the stores at cluster scope)
            fence_proxy_async_all();
            tc_fence_before();                 // ... and order this step's TMEM reads before the next step's MMAs
            cluster_arrive_release();
            cluster_wait_acquire();
            tc_fence_after();
        }
      }
    }
    tc_fence_before();
    if (PAIR) { cluster_arrive_relaxed(); cluster_wait(); }      // neither CTA leaves while the pair's MMAs / arrivals may still touch it (no data ordered)
    else __syncthreads();
    if (warp == 1) { if (PAIR) tmem_dealloc_2(tmem_d, kBigTmemCols); else tmem_dealloc(tmem_d, kBigTmemCols); }
}

// ---- staging ---------------------------------------------------------------------------------------------------------
// SC (float64 [N][N], row = target node) -> B image hi/lo
// R = rows per block: 256 ([slice][KG][256][4]) or, for the CTA-pair kernel, 128 ([slice][half][KG][128][4])
__global__ void big_stage_b_kernel(const double* CM, int N, int KG, int slices, int mixed, int R, float* Bimg) {
    const size_t total = (size_t)slices * KG * kBigNT * 4;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int kk = (int)(idx & 3);
    const int row = (int)((idx >> 2) % R);
    const size_t rest = (idx >> 2) / R;
    const int kg = (int)(rest % KG);
    const size_t blk = rest / KG;
    const int per = kBigNT / R;
    const int n = (int)(blk / per) * kBigNT + (int)(blk % per) * R + row, k = kg * 4 + kk;
    const float v = (n < N && k < N) ? (float)CM[(size_t)n * N + k] : 0.f;
    if (!mixed) {
        const float h = tf32_rn(v);
        Bimg[idx] = h;
        Bimg[total + idx] = v - h;
    } else {
        Bimg[idx] = v;
        __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(Bimg + total);       // [slice][KG/2][256][8]
        __nv_bfloat16* L = H + total;
        const size_t o = ((blk * (KG / 2) + (kg >> 1)) * R + row) * 8 + (size_t)((kg & 1) * 4 + kk);
        const __nv_bfloat16 hb = __float2bfloat16_rn(v);
        H[o] = hb;
        L[o] = __float2bfloat16_rn(mixed == 2 ? v - __bfloat162float(hb) : v - tf32_trunc(v));
    }
}

__global__ void big_stage_maps_kernel(const double* mapG, const double* mapS, int N, int Kpad, float* mG, float* mS) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    mG[k] = (k < N && mapG) ? (float)mapG[k] : 1.f;
    mS[k] = (k < N && mapS) ? (float)mapS[k] : 1.f;
}

// node_params [NREM_NODE_PARAMS][N] float64 (a_ee a_ei a_ii tauE tauI P rhoE rE rI mu sigmaI a_ie_0) -> the float32 table of the NODEPAR
// kernels, [kBigNpar][Kpad] with the derived quantities the step code uses (the same float32 expressions as make_const for scalars)
__global__ void big_stage_npar_kernel(const double* np, int N, int Kpad, double dtSim, float* out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    const bool live = k < N;
    auto at = [&](int row) { return live ? np[(size_t)row * N + k] : 1.0; };
    const float P = (float)at(5), mu = (float)at(9);
    out[kNpAee * Kpad + k] = (float)at(0); out[kNpAei * Kpad + k] = (float)at(1); out[kNpAii * Kpad + k] = (float)at(2);
    out[kNpKE * Kpad + k] = (float)(dtSim / at(3)); out[kNpKI * Kpad + k] = (float)(dtSim / at(4));
    out[kNpPmu * Kpad + k] = P - mu; out[kNpRhoE * Kpad + k] = (float)at(6); out[kNpRE * Kpad + k] = (float)at(7);
    out[kNpRI * Kpad + k] = (float)at(8); out[kNpNmu * Kpad + k] = -mu; out[kNpSigI2 * Kpad + k] = (float)(-at(10) * 1.4426950408889634);
    out[kNpA0 * Kpad + k] = (float)at(11);
}

// initial condition (netwWilsonCowanPlastic.py:90-99) into the images; every array was zeroed before
__global__ void big_init_kernel(BatchConst c, int64_t nf4, int KG, int mixed, float4* A0, size_t plane, float4* I4, float4* ab4, float4* ad4,
                                const float* npar /* NULL or the per-node table: a_ie_0 per node */) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index [tile][kg][r]
    if (idx >= nf4) return;
    const int kg = (int)((idx / kTile) % KG);
    const int64_t tile = idx / ((int64_t)KG * kTile);
    const int r = (int)(idx % kTile);
    float e[4], l[4], i[4], a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool live = kg * 4 + j < c.N;
        e[j] = live ? (mixed ? c.E0 : tf32_rn(c.E0)) : 0.f;
        l[j] = live ? (mixed == 2 ? c.E0 - __bfloat162float(__float2bfloat16_rn(c.E0)) : mixed ? c.E0 - tf32_trunc(c.E0) : c.E0 - e[j]) : 0.f;
        i[j] = live ? c.I0 : 0.f;
        a[j] = live ? (npar ? npar[(size_t)kNpA0 * KG * 4 + kg * 4 + j] : c.a0) : 0.f;
    }
    A0[idx] = make_float4(e[0], e[1], e[2], e[3]);
    if (!mixed) {
        A0[plane + idx] = make_float4(l[0], l[1], l[2], l[3]);
    } else {
        __nv_bfloat16* L = reinterpret_cast<__nv_bfloat16*>(A0 + plane);        // [tile][KG/2][128][8]
        __nv_bfloat16* H = L + plane * 4;
        const size_t o = (((size_t)tile * (KG / 2) + (kg >> 1)) * kTile + r) * 8 + (size_t)((kg & 1) * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) { L[o + j] = __float2bfloat16_rn(l[j]); H[o + j] = __float2bfloat16_rn(e[j]); }
    }
    I4[idx] = make_float4(i[0], i[1], i[2], i[3]);
    if (mixed == 2) {            // bf16 base (the remainder a0 - bf16(a0) goes to the delta array: see big_init_delta below)
        reinterpret_cast<uint2*>(ab4)[idx] = make_uint2((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a[1])) << 16),
                                                        (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a[2])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a[3])) << 16));
        float4 d;
        d.x = a[0] - __bfloat162float(__float2bfloat16_rn(a[0])); d.y = a[1] - __bfloat162float(__float2bfloat16_rn(a[1]));
        d.z = a[2] - __bfloat162float(__float2bfloat16_rn(a[2])); d.w = a[3] - __bfloat162float(__float2bfloat16_rn(a[3]));
        ad4[idx] = d;
    } else ab4[idx] = make_float4(a[0], a[1], a[2], a[3]);
}

// images -> final state [3][N][Bs] (E, I, a_ie), simulation fastest
// (Bo = simulation stride of `fin`; the images may hold one more padding tile)
__global__ void big_export_kernel(int N, int KG, int64_t Bo, int mixed, const float* Aimg, size_t plane_f, const float* I, const float* ab,
                                  const float* ad, float* fin) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)N * Bo) return;
    const int64_t sim = k % Bo;
    const int node = (int)(k / Bo);
    const size_t src = (((size_t)(sim / kTile) * KG + (node >> 2)) * kTile + (size_t)(sim % kTile)) * 4 + (node & 3);
    fin[k] = mixed ? Aimg[src] : Aimg[src] + Aimg[plane_f + src];
    fin[(int64_t)N * Bo + k] = I[src];
    const float base = mixed == 2 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ab)[src]) : ab[src];
    fin[2 * (int64_t)N * Bo + k] = base + ad[src];
}

}  // namespace nrem
