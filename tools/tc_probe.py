"""Diagnostic: checks the tcgen05 contraction (nrem_selftest_tc_coupling) against numpy for a few
descriptor encodings, each in its own process so that a trapping variant cannot poison the others.

    python tools/tc_probe.py            # all variants
    python tools/tc_probe.py <variant>  # one variant (worker)
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "default_1x": dict(passes=1),
    "default_3x": dict(passes=3),
    "swapped_1x": dict(passes=1, lboA=128, sboA=2048, lboB=128, sboB=1536),
}


def tf32(x):
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def worker(name):
    from nremmodfc_b200 import ops
    rng = np.random.default_rng(0)
    E = rng.random((128, 96), dtype=np.float32)
    SC = (rng.random((96, 96), dtype=np.float32) * (rng.random((96, 96)) < 0.4)).astype(np.float32)
    kw = VARIANTS[name]
    out = ops.selftest_tc_coupling(E, SC, **kw)
    exact = E.astype(np.float64) @ SC.astype(np.float64).T
    ref1 = tf32(E).astype(np.float64) @ tf32(SC).astype(np.float64).T
    err_exact = np.abs(out - exact).max() / np.abs(exact).max()
    err_tf32 = np.abs(out - ref1).max() / np.abs(exact).max()
    print(f"{name}: rel err vs float64 {err_exact:.3e}, vs TF32-rounded operands {err_tf32:.3e}, out[0,:4]={out[0,:4]}, exact[0,:4]={exact[0,:4]}")
    if err_exact > 1e-2:
        # which structured permutation is it?  try matching single-hot probes
        Ep = np.zeros((128, 96), np.float32)
        Sp = np.zeros((96, 96), np.float32)
        Ep[5, 7] = 1.0
        Sp[11, 7] = 1.0
        o = ops.selftest_tc_coupling(Ep, Sp, **kw)
        print("  one-hot E[5,7] x SC[11,7]: nonzeros at", np.argwhere(o != 0)[:8].tolist(), "(expect [[5, 11]])")
        for (m, k) in [(5, 0), (5, 4), (13, 0), (0, 1)]:
            Ep[:] = 0
            Sp[:] = 0
            Ep[m, k] = 1.0
            Sp[:, :] = np.arange(96 * 96, dtype=np.float32).reshape(96, 96) + 1
            o = ops.selftest_tc_coupling(Ep, Sp, **kw)
            nz = np.argwhere(o != 0)
            print(f"  E one-hot ({m},{k}): rows {sorted(set(nz[:,0].tolist()))[:6]} first vals {o[nz[0][0], :3] if len(nz) else None}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker(sys.argv[1])
    else:
        for v in VARIANTS:
            try:
                r = subprocess.run([sys.executable, __file__, v], capture_output=True, text=True, timeout=120)
                print(r.stdout.strip() or f"{v}: no output")
                if r.returncode != 0:
                    print(f"{v}: exit {r.returncode}: {r.stderr.strip()[-400:]}")
            except subprocess.TimeoutExpired:
                print(f"{v}: TIMEOUT")
