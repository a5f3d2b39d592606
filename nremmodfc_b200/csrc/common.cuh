// Shared helpers of the nremfc CUDA library: error reporting, launch accounting, math intrinsics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nremfc.h"

namespace nrem {

extern thread_local char g_err[512];
extern thread_local int64_t g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define NREM_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            cudaGetLastError();      /* reset the per-thread error so that the next call does not report this one again */ \
            return nrem::fail(NREM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));      \
        }                                                                                   \
    } while (0)

#define NREM_REQUIRE(cond, msg)                                          \
    do {                                                                 \
        if (!(cond)) return nrem::fail(NREM_ERR_ARG, "%s (%s)", msg, #cond); \
    } while (0)

#define NREM_LAUNCHED()                                                  \
    do {                                                                 \
        ++nrem::g_launches;                                              \
        NREM_CUDA(cudaGetLastError());                                   \
    } while (0)

// ---- single-instruction MUFU wrappers (ftz: every operand here is a normal number) ----
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrtaf(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sinaf(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cosaf(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

}  // namespace nrem
