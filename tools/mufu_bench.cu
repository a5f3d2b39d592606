// Micro-benchmark: per-SM throughput of the MUFU (XU pipe) operations used by the integrator.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench tools/mufu_bench.cu && /tmp/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>

#define OPS(X) X(ex2, "ex2.approx.ftz.f32") X(rcp, "rcp.approx.ftz.f32") X(lg2, "lg2.approx.ftz.f32") X(sqrt, "sqrt.approx.ftz.f32") \
               X(rsqrt, "rsqrt.approx.ftz.f32") X(sin, "sin.approx.ftz.f32") X(cos, "cos.approx.ftz.f32") X(tanh, "tanh.approx.f32")

#define DEF(name, ins)                                                                    \
    __global__ void k_##name(float* out, int iters) {                                     \
        float x[8];                                                                       \
        for (int i = 0; i < 8; ++i) x[i] = 0.3f + 0.01f * (threadIdx.x + i);              \
        for (int it = 0; it < iters; ++it) {                                              \
            _Pragma("unroll") for (int u = 0; u < 4; ++u)                                 \
                _Pragma("unroll") for (int i = 0; i < 8; ++i)                             \
                    asm volatile(ins " %0, %0;" : "+f"(x[i]));                            \
        }                                                                                 \
        float s = 0;                                                                      \
        for (int i = 0; i < 8; ++i) s += x[i];                                            \
        if (s == 123.f) out[0] = s;                                                       \
    }
OPS(DEF)

// mix: 1 MUFU per M FFMA
template <int M>
__global__ void k_mix(float* out, int iters) {
    float x[8], y[8];
    for (int i = 0; i < 8; ++i) { x[i] = 0.3f + 0.01f * (threadIdx.x + i); y[i] = x[i] + 1.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
                for (int m = 0; m < M; ++m) y[i] = fmaf(y[i], 0.999f, 0.001f);
            }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    if (s == 123.f) out[0] = s;
}

template <typename F>
static double run(F f, int sms, int threads, int iters, double ops_per_thread_iter) {
    float* d;
    cudaMalloc(&d, 64);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f<<<sms, threads>>>(d, 10);
    cudaEventRecord(a);
    f<<<sms, threads>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(d);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    return ops_per_thread_iter * iters * threads / cycles;     // lane-ops per clock per SM (at the nominal max clock)
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    for (int threads : {512, 1024}) {
        printf("threads/SM = %d (one CTA per SM)\n", threads);
#define RUN(name, ins) printf("  %-6s %6.2f lane-ops/clk/SM\n", #name, run(k_##name, sms, threads, iters, 32.0));
        OPS(RUN)
        printf("  ex2 + 2 ffma: %6.2f MUFU/clk/SM\n", run(k_mix<2>, sms, threads, iters, 32.0));
        printf("  ex2 + 8 ffma: %6.2f MUFU/clk/SM\n", run(k_mix<8>, sms, threads, iters, 32.0));
        printf("  ex2 + 16 ffma: %6.2f MUFU/clk/SM\n", run(k_mix<16>, sms, threads, iters, 32.0));
    }
    return 0;
}
