"""Philox4x32-7 + Box-Muller noise stream "nrem-philox-v2" (numpy reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws its noise from numba's MT19937 stream, which it never
seeds (SURVEY.md "established by running" item 3): a "seed" there is only a
replicate label.  The B200 path therefore defines its own counter-based stream
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11 —
Philox4x32 with ROUNDS = 7 rounds, the smallest count that passes BigCrush in
that paper; stream v1 of round 1 used the library default of 10) and this file
restates it on the CPU so the CUDA generator can be checked bit-for-bit
(integers, incl. Random123's known-answer vectors for 7 and 10 rounds) and to
1 ulp-ish (normals).

Stream definition
-----------------
key      = (seed & 0xffffffff, seed >> 32)                 user seed, 64 bit
counter  = (step, quad, stream & 0xffffffff, stream >> 32)
           step   : global Euler step index 0 .. n1+n2+n3-1
           quad   : node // 4
           stream : per-simulation 64-bit replicate id
outputs  = x0..x3 (uint32)
u(x)     = ((x >> 9) + 0.5) * 2**-23                       in (0,1), exact in fp32
theta(x) = 2 pi (u(x) - 0.5)                             in (-pi, pi): the accurate range of MUFU.SIN/COS
normals  : r0 = sqrt(-2 ln u(x0)); z[4q+0] = r0 cos(theta(x1)); z[4q+1] = r0 sin(theta(x1))
           r1 = sqrt(-2 ln u(x2)); z[4q+2] = r1 cos(theta(x3)); z[4q+3] = r1 sin(theta(x3))
noise    = sqdtD * z     (reference: np.random.normal(0, sqdtD, size=N),
                          netwWilsonCowanPlastic.py:80)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
ROUNDS = 7            # == NREM_PHILOX_ROUNDS of nremmodfc_b200/csrc/philox.cuh


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=ROUNDS):
    """Vectorised Philox4x32-<rounds>.  All arguments broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def uniform23(x):
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23


def normals(seed, stream, step, nnodes):
    """Standard normals z[..., nnodes] (float64) for the given step(s)/stream(s).

    ``stream`` and ``step`` broadcast against each other; the node axis is appended.
    """
    seed = int(seed)
    stream = np.asarray(stream, dtype=np.uint64)
    step = np.asarray(step, dtype=np.uint64)
    stream, step = np.broadcast_arrays(stream, step)
    nq = (nnodes + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32(step[..., None], q, (stream & MASK)[..., None], (stream >> np.uint64(32))[..., None],
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r0 = np.sqrt(-2.0 * np.log(uniform23(x0)))
    r1 = np.sqrt(-2.0 * np.log(uniform23(x2)))
    a0 = 2.0 * np.pi * (uniform23(x1) - 0.5)
    a1 = 2.0 * np.pi * (uniform23(x3) - 0.5)
    z = np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1)
    return z.reshape(z.shape[:-2] + (nq * 4,))[..., :nnodes]
