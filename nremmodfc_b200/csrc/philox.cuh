// Counter-based noise stream "nrem-philox-v2" (definition: oracle/philox.py): Philox4x32 with 7 rounds + Box-Muller.
// (v1 used 10 rounds.  7 is the smallest round count of Philox4x32 that passes BigCrush — Salmon et al., SC'11, table 2 —
// and the 3 rounds less are 6 % of the integrator's time: profiles/r02_kernel_variants.md.)
// Replaces np.random.normal(0, sqdtD, size=N) of netwWilsonCowanPlastic.py:80, whose numba
// MT19937 stream the reference never seeds.
#pragma once
#include "common.cuh"

namespace nrem {

struct Philox4 { uint32_t x, y, z, w; };

#ifndef NREM_PHILOX_ROUNDS
#define NREM_PHILOX_ROUNDS 7           // the stream definition (oracle/philox.py ROUNDS); other values are timing experiments only
#endif

__device__ __forceinline__ Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < NREM_PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// u = ((x >> 9) + 0.5) * 2^-23, exact in fp32: build 1.m and subtract (1 - 2^-24).
__device__ __forceinline__ float u23f(uint32_t x) { return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994f; }
__device__ __forceinline__ double u23d(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }

// Four standard normals of one counter (fp32, MUFU lg2/sqrt/sin/cos).
__device__ __forceinline__ void normals4f(const Philox4 r, float& z0, float& z1, float& z2, float& z3) {
    const float two_pi = 6.2831853071795865f;
    const float m2ln2 = -1.3862943611198906f;          // -2 ln 2
    const float r0 = sqrtaf(m2ln2 * lg2f(u23f(r.x)));
    const float r1 = sqrtaf(m2ln2 * lg2f(u23f(r.z)));
    // theta = 2 pi (u - 0.5) with u = m - (1 - 2^-24), m = 1.mantissa: one FMA
    const float a0 = fmaf(__uint_as_float(0x3f800000u | (r.y >> 9)), two_pi, -1.49999994f * two_pi);
    const float a1 = fmaf(__uint_as_float(0x3f800000u | (r.w >> 9)), two_pi, -1.49999994f * two_pi);
    z0 = r0 * cosaf(a0); z1 = r0 * sinaf(a0);
    z2 = r1 * cosaf(a1); z3 = r1 * sinaf(a1);
}

// Same stream in float64 (parity kernels).
__device__ __forceinline__ void normals4d(const Philox4 r, double z[4]) {
    const double two_pi = 6.283185307179586476925286766559;
    const double r0 = sqrt(-2.0 * log(u23d(r.x))), a0 = two_pi * (u23d(r.y) - 0.5);
    const double r1 = sqrt(-2.0 * log(u23d(r.z))), a1 = two_pi * (u23d(r.w) - 0.5);
    double s0, c0, s1, c1;
    sincos(a0, &s0, &c0);
    sincos(a1, &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

}  // namespace nrem
