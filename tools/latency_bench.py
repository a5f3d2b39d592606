"""Per-step latency of the integrator kernels on SMALL batches (the time loop is sequential: for a batch that cannot fill the SMs
the only lever is the time one Euler step of a tile takes).

    python tools/latency_bench.py [euler_steps]

Batches: 200 simulations (run_many_seeds.py: 50 seeds x 4 states), 2500 (one GPU's share of a 20 000-simulation sweep on 8 GPUs),
7500 (its share of the paper's 60 000-simulation job), 18944 (one full wave of 128-simulation tiles); kernels: tc3 (128
simulations per CTA), node32, node16 (wc_node.cuh).  Prints microseconds per Euler step and the projected time of a full-length
(10.01 M step) run.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nremmodfc_b200 import ops  # noqa: E402

d = np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
rng = np.random.default_rng(0)
print(f"| simulations | kernel | tiles | us / Euler step | full-length run (10.01 M steps) | sims/s |")
print("|---:|---|---:|---:|---:|---:|")
for B in (200, 2500, 7500, 18944):
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    for kern, tile in (("tc3", 128), ("node32", 32), ("node16", 16)):
        p = ops.make_params(90, 0, steps, 0, P=0.4, rhoE=0.18, seed=1)
        pw = ops.make_params(90, 0, 2000, 0, P=0.4, rhoE=0.18, seed=1)
        ops.integrate_f32(pw, d["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, kernel=kern, record=False)
        torch.cuda.synchronize()
        _, fin = ops.integrate_f32(p, d["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, kernel=kern, record=False)
        ms = ops.last_integrate_ms()
        us = ms * 1e3 / steps
        full = us * 10.01
        print(f"| {B} | {kern} | {(B + tile - 1) // tile} | {us:.3f} | {full:.1f} s | {B / full:.1f} | finite={np.isfinite(fin).all()}")
