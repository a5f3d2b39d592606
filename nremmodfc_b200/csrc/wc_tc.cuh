// tcgen05 variant of the batched integrator: the SC.E contraction of a 128-simulation tile is one
// GEMM per Euler step,  D[128 sims, 96 nodes] = E[128, 96] x SC^T[96, 96],  issued by one thread as
// tcgen05.mma (kind::tf32, M=128, N=96, K=8 per instruction) with both operands in shared memory
// and the FP32 accumulator in tensor memory (TMEM).  TMEM lane = simulation, column = node, which
// is exactly the lane=simulation register layout of wc_batch.cuh, so tcgen05.ld hands every
// thread the coupling of its own (simulation, 24 nodes).
//
//   kernel 2 ("tc")  : operands rounded to TF32 (RN), 12 MMAs per step
//   kernel 3 ("tc3") : 3xTF32 split  E = Eh + El, SC = Sh + Sl;  D = Eh.Sh + El.Sh + Eh.Sl
//                      (36 MMAs per step, ~2^-21 relative error: FP32-grade coupling)
//
// Per step the MMA runs asynchronously while the threads draw the noise and advance I and a_ie
// (none of which needs the coupling); only the E update waits for the accumulator.
//
// Shared-memory operand layout (no swizzle, K-major "interleaved" canonical form): 8x(16 B) core
// matrices, address(m,k) = (m/8)*SBO + (m%8)*16 + (k/4)*LBO + (k%4)*4 with SBO = 128 B, i.e. the
// tile is [k/4][row] float4 — threads of a warp (consecutive simulations) write consecutive
// float4, conflict-free.
#pragma once
#include <type_traits>

#include "wc_batch.cuh"

namespace nrem {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();      // a lost MMA completion must fail loudly, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map); completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 8 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// wait for the loads; the registers are in/out operands so that no use can be hoisted above the wait
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :: "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// wait for three 8-column loads issued back to back (coupling, G_i, sigma_i)
__device__ __forceinline__ void tmem_ld_wait24(uint32_t (&r)[8], uint32_t (&q)[8], uint32_t (&p)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]),
                   "+r"(p[0]), "+r"(p[1]), "+r"(p[2]), "+r"(p[3]), "+r"(p[4]), "+r"(p[5]), "+r"(p[6]), "+r"(p[7])
                 :: "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell).  Offsets in bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}

constexpr int kTcM = kTile;         // 128 simulations
constexpr int kTcN = kNPad;         // 96 output nodes
constexpr int kTcK = kNPad;         // 96 input nodes
constexpr uint32_t kSBO = 128;                      // next 8-row group
constexpr uint32_t kLBO_A = kTcM * 16;              // next 4-column group of A: 2048 B
constexpr uint32_t kLBO_B = kTcN * 16;              // next 4-column group of B: 1536 B
constexpr uint32_t kABytes = (kTcK / 4) * kLBO_A;   // 49152
constexpr uint32_t kBBytes = (kTcK / 4) * kLBO_B;   // 36864
constexpr uint32_t kTmemCols = 128;                 // columns 0..95: coupling accumulator D
constexpr uint32_t kTmemColsMaps = 512;             // heterogeneous maps: + columns 128..223 G_i, 256..351 sigma_i (per simulation, node)
constexpr uint32_t kTmemG = 128, kTmemS = 256;
constexpr uint32_t kRecombine = 4096;               // a_base <- a_base + delta at global steps that are multiples of this
// instruction descriptor: D=F32 (bits 4-5 = 1), A=B=TF32 (bits 7-9 / 10-12 = 2), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// The B operand image of the integrator: SC (padded [96][96] float32) split into TF32 hi / residual lo in the exact shared-memory
// layout of the MMA, [hi: kBBytes][lo: kBBytes].  Built ONCE per sweep (stage_inputs); every launch then brings it into shared
// memory with one bulk copy of the TMA engine.
__global__ void stage_sc_image_kernel(const float* SCp, float* img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kTcN * kTcK) return;
    const int n = idx / kTcK, k = idx % kTcK;
    const float v = SCp[idx];
    const float h = tf32_rn(v);
    const int o = (k >> 2) * (int)(kLBO_B / 4) + n * 4 + (k & 3);
    img[o] = h;
    img[kBBytes / 4 + o] = v - h;
}

// Stage SC (padded [96][96] float32 in global) as the B operand(s) with generic loads (self-test, callers without an image).
template <int NPASS>
__device__ __forceinline__ void stage_b(const float* SCp, float* Bh, float* Bl, int tid, int nthreads) {
    for (int idx = tid; idx < kTcN * kTcK; idx += nthreads) {
        const int n = idx / kTcK, k = idx % kTcK;
        const float v = SCp[idx];
        const float h = tf32_rn(v);
        const int o = (k >> 2) * (kLBO_B / 4) + n * 4 + (k & 3);
        Bh[o] = h;
        if (NPASS == 3) Bl[o] = v - h;
    }
}

// One elected thread: issue the MMAs of one Euler step and commit them to `bar`.  adesc/bdesc are the
// descriptors of the first K-slice; the next slice is 2 four-column groups further (start-address field += bytes/16).
template <int NPASS>
__device__ __forceinline__ void issue_coupling(uint64_t ad_hi, uint64_t ad_lo, uint64_t bd_hi, uint64_t bd_lo, uint32_t tmem_d,
                                               uint32_t idesc, uint64_t* bar) {
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
        const uint64_t a0 = (pass == 1) ? ad_lo : ad_hi;
        const uint64_t b0 = (pass == 2) ? bd_lo : bd_hi;
#pragma unroll
        for (int kk = 0; kk < kTcK / 8; ++kk) {
            umma_tf32(tmem_d, a0 + (uint64_t)(kk * ((2 * kLBO_A) >> 4)), b0 + (uint64_t)(kk * ((2 * kLBO_B) >> 4)), idesc, acc);
            acc = 1;
        }
    }
    umma_commit(bar);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

template <int NPASS>
constexpr int tc_smem_bytes() { return (NPASS == 3 ? 2 : 1) * (int)(kABytes + kBBytes) + 2 * kNPad * 4 + 32 + (int)kABytes; }

// CH = nodes per thread (24, 16 or 12): the CTA has (96/CH) * 4 warps.
// Measured alternatives that LOST on B200 (profiles/r01_kernel_variants.md): an elected issuer warp with a
// bar.arrive/bar.sync split (+6 %), skipping the padding quads with warp-uniform branches (+37 %: the branches
// stop the compiler from interleaving quads), 16 or 12 nodes per thread (+2 %).
// HOMO : homogeneous sweep (every map entry is 1, so G_i and sigma_i are per-simulation scalars: -2 FMA, -2 LDS per
//        node-step).  Constants (-mu, -kA*rhoE) are folded into the FMAs.
// LIGHT: N == 90 only.  The last node chunk (nodes 72..95) holds 18 real nodes and 6 padding nodes; its warps run a
//        second copy of the step loop compiled for 18 nodes, and the MMA issuer is the first warp of that chunk (one lane
//        chosen by elect.sync), so the issue of the 36 tcgen05.mma per step is taken from the padding slack instead of
//        making the other 15 warps wait at the CTA barrier (ncu: 11.7 % of all samples were that wait); the per-step CTA
//        barrier is split into bar.arrive (workers) / bar.sync (issuer warp).  The 24-node copy carries no issue code.
// a_ie : the plasticity increment dtSim/tau_ip * I (E - rhoE) is ~5e-7 per step while a_ie is 2.5..10, i.e. about ONE
//        float32 ulp: accumulated directly in float32 the homeostatic loop loses the small corrections (measured: the
//        homogeneous high-G cells of the full sweep drift off the reference's table).  So a_ie = a_base + delta: a_base
//        sits in shared memory ([node/4][sim] float4, one conflict-free LDS.128 per quad and step; keeping it in spare
//        TMEM columns instead costs 5 % more because tcgen05.wait::ld cannot be hoisted), only the small delta is
//        integrated in registers, and the two are recombined at global steps that are multiples of kRecombine.
// PIPE : tie the Philox rounds of quad g+1 behind a MUFU result of quad g (1: the first lg2, 2: the first normal) so
//        that ptxas cannot hoist all integer work in front of all MUFU work (-2 %).
template <int NPASS, int CH, bool HOMO, bool LIGHT, int PIPE>
__global__ void __launch_bounds__((kNPad / CH) * kTile, 1) wc_batch_tc_kernel(const BatchArgs A) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr bool SPLIT = NPASS == 3;
    constexpr int NT = (kNPad / CH) * kTile;
    constexpr int NCHUNK = kNPad / CH;
    float* Ah = reinterpret_cast<float*>(smraw);
    float* Al = reinterpret_cast<float*>(smraw + kABytes);
    float* Bh = reinterpret_cast<float*>(smraw + (SPLIT ? 2 : 1) * kABytes);
    float* Bl = reinterpret_cast<float*>(smraw + (SPLIT ? 2 : 1) * kABytes + kBBytes);
    unsigned char* tail = smraw + (SPLIT ? 2 : 1) * (kABytes + kBBytes);
    float* mG = reinterpret_cast<float*>(tail);
    float* mS = mG + kNPad;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tail + 2 * kNPad * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 2 * kNPad * 4 + 16);
    float4* Ab4 = reinterpret_cast<float4*>(tail + 2 * kNPad * 4 + 32);      // a_base [node/4][sim]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = warp >> 2;
    const int simt = ((warp & 3) << 5) | lane;
    const int tile = A.tile0 + (int)blockIdx.x;
    const int64_t sim = (int64_t)tile * kTile + simt;
    const BatchConst& c = A.c;
    const int N = c.N;
    const bool issuer = tid == 4 * (NCHUNK - 1) * 32;      // first thread of the last chunk

    uint64_t* barB = bar + 1;                                                // SC image landed (bulk copy)
    if (!A.SCimg) stage_b<NPASS>(A.SCp, Bh, Bl, tid, NT);
    for (int k = tid; k < (int)(kABytes / 4) * (SPLIT ? 2 : 1); k += NT) Ah[k] = 0.f;     // padding columns of A stay 0
    const int mid = A.tile_map[tile];
    if (tid < kNPad) { mG[tid] = A.mapG[mid * kNPad + tid]; mS[tid] = A.mapS[mid * kNPad + tid]; }
    constexpr bool MAPS_TMEM = !HOMO && CH == 24;       // per-node G_i, sigma_i parked in spare TMEM columns (see the E update)
    if (warp == 0) tmem_alloc(tmem_slot, MAPS_TMEM ? kTmemColsMaps : kTmemCols);
    if (tid == 32) {
        mbar_init(bar, 1); mbar_init(barB, 1); fence_barrier_init();
        if (A.SCimg) {                    // SC staged once per sweep as the operand image: ONE TMA bulk copy per launch (hi [+ lo], contiguous)
            constexpr uint32_t bytes = (SPLIT ? 2u : 1u) * kBBytes;
            mbar_expect_tx(barB, bytes);
            bulk_g2s(smem_u32(Bh), A.SCimg, bytes, barB);
        }
    }
    fence_proxy_async();                  // A padding / B tile (generic-proxy stores) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (A.SCimg) mbar_wait(barB, 0);
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t tmem_mine = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chunk * CH);

    float E[CH], I[CH], a[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int node = chunk * CH + k;
        if (node < N) {
            if (A.init) { E[k] = c.E0; I[k] = c.I0; a[k] = c.a0; }
            else {
                E[k] = A.state[(0 * (int64_t)N + node) * A.Bs + sim];
                I[k] = A.state[(1 * (int64_t)N + node) * A.Bs + sim];
                a[k] = A.state[(2 * (int64_t)N + node) * A.Bs + sim];
            }
        } else { E[k] = 0.f; I[k] = 0.f; a[k] = 0.f; }
    }
    // a_ie = a_base (shared memory) + delta (registers, a[] from here on); both are part of the stored state (components 2, 3), so
    // that launch boundaries never round: a_base <- a_base + delta happens only at global steps that are multiples of
    // kRecombine, which makes the result independent of how the run is cut into launches
#pragma unroll
    for (int g = 0; g < CH / 4; ++g) {
        Ab4[(chunk * (CH / 4) + g) * kTile + simt] = make_float4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * g + j, node = chunk * CH + k;
            a[k] = (node < N && !A.init) ? A.state[(3 * (int64_t)N + node) * A.Bs + sim] : 0.f;
        }
    }
    const float G0 = A.par[sim], dG = A.par[A.Bs + sim];
    // (explicit roundings: with -fmad the compiler would otherwise contract sg0 + dsg into an FMA of the unrounded product, and the
    // homogeneous kernel would differ in the last bit from fmaf(dsg, 1, sg0) of the map kernels / wc_node.cuh)
    const float sg0 = __fmul_rn(-1.4426950408889634f, A.par[2 * A.Bs + sim]), dsg = __fmul_rn(-1.4426950408889634f, A.par[3 * A.Bs + sim]);
    const uint64_t strm = A.streams[sim];
    const uint32_t s_lo = (uint32_t)strm, s_hi = (uint32_t)(strm >> 32);
    if (MAPS_TMEM) {
        // G_i = G0 + dG mapG_i and sigma_i (in ex2 units) are constant over the launch: one tcgen05.st per 8 nodes here replaces
        // 2 LDS + 2 FFMA per node and Euler step; they come back with the coupling in the same tcgen05.wait::ld
#pragma unroll
        for (int h = 0; h < CH / 8; ++h) {
            uint32_t g8[8], s8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int node = chunk * CH + 8 * h + j;
                g8[j] = __float_as_uint(fmaf(dG, mG[node], G0));
                s8[j] = __float_as_uint(fmaf(dsg, mS[node], sg0));
            }
            tmem_st8(tmem_mine + kTmemG + 8 * h, g8);
            tmem_st8(tmem_mine + kTmemS + 8 * h, s8);
        }
        tmem_st_wait();
    }
    const uint64_t ad_hi = umma_desc(smem_u32(Ah), kLBO_A, kSBO), ad_lo = umma_desc(smem_u32(Al), kLBO_A, kSBO);
    const uint64_t bd_hi = umma_desc(smem_u32(Bh), kLBO_B, kSBO), bd_lo = umma_desc(smem_u32(Bl), kLBO_B, kSBO);
    float4* Ah4 = reinterpret_cast<float4*>(Ah);
    float4* Al4 = reinterpret_cast<float4*>(Al);
    const float Pmu = c.P - c.mu;                        // constant part of the E sigmoid argument
    const float nmu = -c.mu, nkr = -A.kA * c.rhoE;
    const float Gh = __fadd_rn(G0, dG), sgh = __fadd_rn(sg0, dsg);     // HOMO: map == 1 everywhere (== fmaf(dG, 1, G0), fmaf(dsg, 1, sg0))
    const float two_pi = 6.2831853071795865f, m2ln2 = -1.3862943611198906f;

    // The step loop for a chunk with KN live nodes (KN = CH, or 18 for the last chunk of N = 90).
    auto run = [&](auto KNc) {
        constexpr int KN = decltype(KNc)::value;
        constexpr int NQ = (KN + 3) / 4;
        int rc = A.rec_phase;
        int64_t row = A.row0;
        float xp[CH];
        for (int it = 0; it < A.nsteps; ++it) {
            // 1. publish E(t) as the A operand
#pragma unroll
            for (int g = 0; g < NQ; ++g) {
                const float4 v = make_float4(E[4 * g], E[4 * g + 1], E[4 * g + 2], E[4 * g + 3]);
                const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                Ah4[(chunk * (CH / 4) + g) * kTile + simt] = h;
                if (SPLIT) Al4[(chunk * (CH / 4) + g) * kTile + simt] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
            fence_proxy_async();
            tc_fence_before();             // this thread's tcgen05.ld of the previous step precede the barrier
            // Every warp has published its slice and drained its TMEM loads.  With LIGHT only the issuer's warp WAITS for
            // that (it has the 18-node chunk, so it is early anyway); the others just arrive and go on with the
            // coupling-free work — they meet the result at the MMA-completion mbarrier (-7 % per step).
            if (LIGHT) {
                if (warp == 4 * (NCHUNK - 1)) asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
                else asm volatile("bar.arrive 1, %0;" ::"r"(NT) : "memory");
            } else {
                __syncthreads();
            }
            // (with LIGHT the issuer runs the 18-node copy of this loop: the 24-node copy carries no MMA code, -5.6 KB of hot
            // instruction footprint)
            if constexpr (!LIGHT || KN != CH) {
                if (LIGHT) {
                    // warp-uniform branch + elect.sync: ptxas then knows a single lane issues and moves the descriptors to uniform
                    // registers once, instead of wrapping every tcgen05.mma in a lane-by-lane (BRA.U.ANY) loop
                    if (warp == 4 * (NCHUNK - 1)) {
                        if (elect_one()) {
                            tc_fence_after();
                            issue_coupling<NPASS>(ad_hi, ad_lo, bd_hi, bd_lo, tmem_d, kIdescTf32, bar);
                        }
                        __syncwarp();
                    }
                } else if (issuer) {
                    tc_fence_after();
                    issue_coupling<NPASS>(ad_hi, ad_lo, bd_hi, bd_lo, tmem_d, kIdescTf32, bar);
                }
            }
            // 2. record E(t) (state BEFORE the update, netwWilsonCowanPlastic.py:129-130)
            if (A.rec) {
                if (rc == 0) {
                    // one base pointer per row, then a constant stride per node (LIGHT: N == 90, every node of the loop is live)
                    float* dst = A.Ebuf + (row * N + chunk * CH) * A.Bs + sim;
#pragma unroll
                    for (int k = 0; k < KN; ++k) {
                        if (LIGHT || chunk * CH + k < N) dst[(int64_t)k * A.Bs] = E[k];
                    }
                    ++row;
                }
                if (++rc == A.downsamp) rc = 0;
            }
            // 3. everything that does not need the coupling, while the tensor core works
            const uint32_t step = A.step0 + (uint32_t)it;
            if ((step & (kRecombine - 1)) == 0 && step != 0) {       // rare: fold delta into a_base
#pragma unroll
                for (int g = 0; g < CH / 4; ++g) {
                    float4 v = Ab4[(chunk * (CH / 4) + g) * kTile + simt];
                    v.x += a[4 * g]; v.y += a[4 * g + 1]; v.z += a[4 * g + 2]; v.w += a[4 * g + 3];
                    a[4 * g] = a[4 * g + 1] = a[4 * g + 2] = a[4 * g + 3] = 0.f;
                    Ab4[(chunk * (CH / 4) + g) * kTile + simt] = v;
                }
            }
            float4 ab4 = make_float4(0.f, 0.f, 0.f, 0.f);      // a_base of the current quad
            auto node_pre = [&](int k) {
                const float abk = (k & 3) == 0 ? ab4.x : (k & 3) == 1 ? ab4.y : (k & 3) == 2 ? ab4.z : ab4.w;
                xp[k] = fmaf(-abk, I[k], fmaf(-a[k], I[k], fmaf(c.a_ee, E[k], xp[k])));
                const float y = fmaf(-c.a_ii, I[k], fmaf(c.a_ei, E[k], nmu));
                const float SI = rcpf(1.0f + ex2f(y * c.sigI2));
                a[k] = fmaf(I[k], fmaf(E[k], A.kA, nkr), a[k]);
                I[k] = fmaf(c.kI, fmaf(fmaf(-c.rI, I[k], 1.0f), SI, -I[k]), I[k]);
            };
            Philox4 ph = philox4x32(step, (uint32_t)(chunk * (CH / 4)), s_lo, s_hi, c.k0, c.k1);
#pragma unroll
            for (int g = 0; g < NQ; ++g) {
                const float l0 = lg2f(u23f(ph.x)), l1 = lg2f(u23f(ph.z));
                const float a0 = fmaf(__uint_as_float(0x3f800000u | (ph.y >> 9)), two_pi, -1.49999994f * two_pi);
                const float a1 = fmaf(__uint_as_float(0x3f800000u | (ph.w >> 9)), two_pi, -1.49999994f * two_pi);
                const float r0 = sqrtaf(m2ln2 * l0), r1 = sqrtaf(m2ln2 * l1);
                const float z0 = r0 * cosaf(a0);
                if (g + 1 < NQ) {
                    // PIPE != 0: A.zero is 0 at run time but opaque to ptxas, so the next quad's counter truly depends on a
                    // MUFU result of this quad (one LOP3 per quad) and its Philox rounds cannot be hoisted in front of it
                    const uint32_t ctr = PIPE == 0 ? step : (step ^ (__float_as_uint(PIPE == 1 ? l0 : z0) & A.zero));
                    ph = philox4x32(ctr, (uint32_t)(chunk * (CH / 4) + g + 1), s_lo, s_hi, c.k0, c.k1);
                }
                xp[4 * g + 0] = fmaf(c.sq, z0, Pmu);
                xp[4 * g + 1] = fmaf(c.sq, r0 * sinaf(a0), Pmu);
                xp[4 * g + 2] = fmaf(c.sq, r1 * cosaf(a1), Pmu);
                xp[4 * g + 3] = fmaf(c.sq, r1 * sinaf(a1), Pmu);
                ab4 = Ab4[(chunk * (CH / 4) + g) * kTile + simt];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * g + j < KN) node_pre(4 * g + j);
            }
            // 4. coupling -> E(t+1)
            mbar_wait(bar, (uint32_t)(it & 1));
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < (KN + 7) / 8; ++h) {
                uint32_t cr[8], g8[8], s8[8];
                if (8 * h + 8 <= CH) tmem_ld8(tmem_mine + 8 * h, cr); else tmem_ld4(tmem_mine + 8 * h, cr);
                if (MAPS_TMEM) {
                    tmem_ld8(tmem_mine + kTmemG + 8 * h, g8);
                    tmem_ld8(tmem_mine + kTmemS + 8 * h, s8);
                    tmem_ld_wait24(cr, g8, s8);
                } else {
                    tmem_ld_wait8(cr);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * h + j;
                    if (k < KN) {
                        const int node = chunk * CH + k;
                        const float Gi = HOMO ? Gh : MAPS_TMEM ? __uint_as_float(g8[j]) : fmaf(dG, mG[node], G0);
                        const float sg2 = HOMO ? sgh : MAPS_TMEM ? __uint_as_float(s8[j]) : fmaf(dsg, mS[node], sg0);
                        const float x = fmaf(Gi, __uint_as_float(cr[j]), xp[k]);
                        const float SE = rcpf(1.0f + ex2f(x * sg2));
                        E[k] = fmaf(c.kE, fmaf(fmaf(-c.rE, E[k], 1.0f), SE, -E[k]), E[k]);
                    }
                }
            }
        }
    };
    if (LIGHT && chunk == NCHUNK - 1) run(std::integral_constant<int, 90 - (NCHUNK - 1) * CH>{});
    else run(std::integral_constant<int, CH>{});

    float abase[CH];
#pragma unroll
    for (int g = 0; g < CH / 4; ++g) {
        const float4 v = Ab4[(chunk * (CH / 4) + g) * kTile + simt];
        abase[4 * g] = v.x; abase[4 * g + 1] = v.y; abase[4 * g + 2] = v.z; abase[4 * g + 3] = v.w;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, MAPS_TMEM ? kTmemColsMaps : kTmemCols);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int node = chunk * CH + k;
        if (node < N) {
            A.state[(0 * (int64_t)N + node) * A.Bs + sim] = E[k];
            A.state[(1 * (int64_t)N + node) * A.Bs + sim] = I[k];
            A.state[(2 * (int64_t)N + node) * A.Bs + sim] = abase[k];
            A.state[(3 * (int64_t)N + node) * A.Bs + sim] = a[k];
        }
    }
}

template <int NPASS, bool HOMO, bool LIGHT, int PIPE>
static int launch_wc_tc_v(const BatchArgs& A, int64_t tiles, cudaStream_t st) {
    auto kern = wc_batch_tc_kernel<NPASS, 24, HOMO, LIGHT, PIPE>;
    NREM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<NPASS>()));
    kern<<<(unsigned)tiles, (kNPad / 24) * kTile, tc_smem_bytes<NPASS>(), st>>>(A);
    NREM_LAUNCHED();
    return NREM_OK;
}

// experiment switches (defaults are the measured winners): NREM_TC_PIPE = 0|1|2, NREM_TC_LIGHT = 0|1.  The PIPE ties paid -2 % on the
// early kernels; on the final one (elect-based issue, lean recording) the untied schedule is 2.3 % faster, so the default is 0.
static int g_tc_pipe = []() { const char* e = getenv("NREM_TC_PIPE"); const int v = e ? atoi(e) : 0; return (v >= 0 && v <= 2) ? v : 0; }();
static int g_tc_light = []() { const char* e = getenv("NREM_TC_LIGHT"); return e ? (atoi(e) != 0) : 1; }();

template <int NPASS, bool HOMO>
static int launch_wc_tc_h(bool light, const BatchArgs& A, int64_t tiles, cudaStream_t st) {
    if (NPASS == 3 && g_tc_pipe == 1) return light ? launch_wc_tc_v<NPASS, HOMO, true, 1>(A, tiles, st) : launch_wc_tc_v<NPASS, HOMO, false, 1>(A, tiles, st);
    if (NPASS == 3 && g_tc_pipe == 2) return light ? launch_wc_tc_v<NPASS, HOMO, true, 2>(A, tiles, st) : launch_wc_tc_v<NPASS, HOMO, false, 2>(A, tiles, st);
    return light ? launch_wc_tc_v<NPASS, HOMO, true, 0>(A, tiles, st) : launch_wc_tc_v<NPASS, HOMO, false, 0>(A, tiles, st);
}

// kernel: 2 = TF32, 3 = 3xTF32; homo: every map entry is exactly 1
static int launch_wc_tc(int kernel, bool homo, const BatchArgs& A, int64_t tiles, cudaStream_t st) {
    const bool light = g_tc_light && A.c.N == 90;
    if (kernel == 3) return homo ? launch_wc_tc_h<3, true>(light, A, tiles, st) : launch_wc_tc_h<3, false>(light, A, tiles, st);
    return homo ? launch_wc_tc_h<1, true>(light, A, tiles, st) : launch_wc_tc_h<1, false>(light, A, tiles, st);
}

// ---- self-test of the contraction alone --------------------------------------------------------
// out[128][96] = E[128][96] x SC[96][96]^T through exactly the staging / descriptor / MMA / TMEM-load
// code of the integrator.  The descriptor fields are arguments so that the encoding can be probed.
template <int NPASS>
__global__ void __launch_bounds__(kBatchThreads, 1) tc_selftest_kernel(const float* Ein, const float* SCp, float* out, uint32_t lboA,
                                                                      uint32_t sboA, uint32_t lboB, uint32_t sboB, uint32_t idesc) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr bool SPLIT = NPASS == 3;
    float* Ah = reinterpret_cast<float*>(smraw);
    float* Al = reinterpret_cast<float*>(smraw + kABytes);
    float* Bh = reinterpret_cast<float*>(smraw + (SPLIT ? 2 : 1) * kABytes);
    float* Bl = reinterpret_cast<float*>(smraw + (SPLIT ? 2 : 1) * kABytes + kBBytes);
    unsigned char* tail = smraw + (SPLIT ? 2 : 1) * (kABytes + kBBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tail + 2 * kNPad * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 2 * kNPad * 4 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = warp >> 2;
    const int simt = ((warp & 3) << 5) | lane;
    stage_b<NPASS>(SCp, Bh, Bl, tid, kBatchThreads);
    for (int idx = tid; idx < kTcM * kTcK; idx += kBatchThreads) {
        const int m = idx / kTcK, k = idx % kTcK;
        const float v = Ein[idx], h = tf32_rn(v);
        const int o = (k >> 2) * (kLBO_A / 4) + m * 4 + (k & 3);
        Ah[o] = h;
        if (SPLIT) Al[o] = v - h;
    }
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    if (tid == 32) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    if (warp == 0 && elect_one())
        issue_coupling<NPASS>(umma_desc(smem_u32(Ah), lboA, sboA), umma_desc(smem_u32(Al), lboA, sboA), umma_desc(smem_u32(Bh), lboB, sboB),
                              umma_desc(smem_u32(Bl), lboB, sboB), tmem_d, idesc, bar);
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t tmem_mine = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chunk * kChunk);
#pragma unroll
    for (int h = 0; h < kChunk / 8; ++h) {
        uint32_t cr[8];
        tmem_ld8(tmem_mine + 8 * h, cr);
        tmem_ld_wait8(cr);
#pragma unroll
        for (int j = 0; j < 8; ++j) out[simt * kTcN + chunk * kChunk + 8 * h + j] = __uint_as_float(cr[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, kTmemCols);
}

}  // namespace nrem
