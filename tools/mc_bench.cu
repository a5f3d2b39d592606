// Micro-benchmark behind profiles/r02_big_connectome.md: does a multicast bulk copy relieve the L2 -> SM limit that bounds the
// operand ring of wc_big.cuh?  128 CTAs in clusters of 4 stream K stages; per stage a CTA needs 16 KB "A" (the same bytes for the 4
// CTAs of a cluster) and 16 KB "B" (the same bytes for CTAs of equal rank in different clusters).
//   mode 0: every CTA loads its 32 KB itself (unicast, what the CTA-pair kernel does)
//   mode 1: every CTA loads a quarter of A and multicasts it to the 4 CTAs of the cluster (+ its own B): 20 KB requested, 32 KB delivered
//   mode 2: B only (16 KB per CTA-stage, unicast) -- the floor if A were free
// Consumers release a stage as soon as it is full (no MMA): a pure delivery test.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mc_bench tools/mc_bench.cu && tools/mc_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NST = 6, STAGE_A = 16384, STAGE_B = 16384, CL = 4;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_mc(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint64_t* bar, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(s32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(64, 1) mc_kernel(const char* A, const char* B, int K, int iters, int mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NST * (STAGE_A + STAGE_B));
    uint64_t* empty = full + NST;
    const uint32_t rank = ctarank();
    const int cluster = blockIdx.x / CL;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, mode == 1 ? CL : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const char* a = A + (size_t)cluster * K * STAGE_A;
    const char* b = B + (size_t)rank * K * STAGE_B;
    const uint32_t base = s32(sm);
    if (threadIdx.x == 0) {                       // producer
        uint32_t ring = 0;
        for (int it = 0; it < iters; ++it)
            for (int kt = 0; kt < K; ++kt, ++ring) {
                const int s = ring % NST;
                mbar_wait(empty + s, ((ring / NST) & 1) ^ 1);
                const uint32_t dst = base + s * (STAGE_A + STAGE_B);
                if (mode == 0) {
                    expect_tx(full + s, STAGE_A + STAGE_B);
                    bulk(dst, a + (size_t)kt * STAGE_A, STAGE_A, full + s);
                    bulk(dst + STAGE_A, b + (size_t)kt * STAGE_B, STAGE_B, full + s);
                } else if (mode == 1) {
                    expect_tx(full + s, STAGE_A + STAGE_B);
                    bulk_mc(dst + rank * (STAGE_A / CL), a + (size_t)kt * STAGE_A + rank * (STAGE_A / CL), STAGE_A / CL, full + s, (uint16_t)((1 << CL) - 1));
                    bulk(dst + STAGE_A, b + (size_t)kt * STAGE_B, STAGE_B, full + s);
                } else {
                    expect_tx(full + s, STAGE_B);
                    bulk(dst + STAGE_A, b + (size_t)kt * STAGE_B, STAGE_B, full + s);
                }
            }
    } else if (threadIdx.x == 32) {               // consumer: release at once
        uint32_t ring = 0;
        for (int it = 0; it < iters; ++it)
            for (int kt = 0; kt < K; ++kt, ++ring) {
                const int s = ring % NST;
                mbar_wait(full + s, (ring / NST) & 1);
                if (mode == 1) { for (uint32_t c = 0; c < CL; ++c) arrive_remote(empty + s, c); }
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + s)) : "memory");
            }
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
    const int K = 63, iters = 200, ctas = 128;
    char *A, *B;
    cudaMalloc(&A, (size_t)(ctas / CL) * K * STAGE_A);
    cudaMalloc(&B, (size_t)CL * K * STAGE_B);
    cudaMemset(A, 1, (size_t)(ctas / CL) * K * STAGE_A);
    cudaMemset(B, 1, (size_t)CL * K * STAGE_B);
    const int smem = NST * (STAGE_A + STAGE_B) + 256;
    cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"unicast A + B (32 KB requested per CTA-stage)", "A multicast to 4 + B (20 KB requested, 32 KB delivered)", "B only (16 KB)"};
    for (int mode = 0; mode < 3; ++mode) {
        mc_kernel<<<ctas, 64, smem>>>(A, B, K, 2, mode);
        cudaEventRecord(e0);
        mc_kernel<<<ctas, 64, smem>>>(A, B, K, iters, mode);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us_stage = ms * 1e3 / ((double)K * iters);
        const double delivered = (mode == 2 ? STAGE_B : STAGE_A + STAGE_B) * (double)ctas / (us_stage * 1e-6) / 1e12;
        printf("mode %d %-60s %s  %.3f us per stage  %.2f TB/s delivered\n", mode, names[mode], cudaGetErrorString(e), us_stage, delivered);
    }
    return 0;
}
