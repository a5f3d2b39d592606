"""Diagnostic: float32 integrator (each kernel) vs the float64 oracle on the same Philox stream, error vs time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nremmodfc_b200 import ops  # noqa: E402
from oracle import cwrap, wc_oracle  # noqa: E402

d = np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"))
n1, n2, n3 = 200, 800, 10000
p = ops.make_params(90, n1, n2, n3, P=0.4, rhoE=0.18, seed=9)
po = wc_oracle.params(P=0.4, rhoE=0.18)
B = 4
streams = np.arange(B, dtype=np.uint64) * 7 + 1
Eo = np.stack([cwrap.wc_run(d["SC"], 0.16, 7.68, n1, n2, n3, seed=9, stream=int(s), p=po, want="E") for s in streams], axis=2)
for kern in ("fma", "tc", "tc3"):
    Eg, _ = ops.integrate_f32(p, d["SC"], np.full(B, 0.16), np.zeros(B), np.full(B, 7.68), np.zeros(B), streams=streams, kernel=kern)
    rel = np.max(np.abs(Eg - Eo) / np.abs(Eo), axis=(1, 2))
    print(kern, "max rel err at rows 0,10,50,100,200,499:", [f"{rel[r]:.2e}" for r in (0, 10, 50, 100, 200, 499)])
