"""Quick integrator-only throughput probe (not the bench): full wave of 148 tiles, a few thousand steps."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nremmodfc_b200 import ops  # noqa: E402

d = np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"))
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 148
B = tiles * 128
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.default_rng(0)
dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
hetero = os.environ.get("NREM_QB_HETERO", "0") != "0"       # heterogeneous NA/ACh maps (the map / shuffled modalities)
mG = d["map_ACh"] / d["map_ACh"].mean() if hetero else None
mS = d["map_NA"] / d["map_NA"].mean() if hetero else None
for kern in sys.argv[3:] or ["fma", "tc", "tc3"]:
    for rec in (False, True):
        p = ops.make_params(90, 0, 0 if rec else steps, steps if rec else 0, P=0.4, rhoE=0.18, seed=1)
        try:
            ops.integrate_f32(p, d["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, kernel=kern, record=False)  # warm
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, fin = ops.integrate_f32(p, d["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, kernel=kern, record=False)
            torch.cuda.synchronize()
            dt = ops.last_integrate_ms() * 1e-3
        except Exception as e:  # noqa: BLE001
            print(kern, "FAILED", e)
            break
        sim_steps = B * steps
        print(f"{kern:4s} rec={int(rec)} tiles={tiles} steps={steps}: {dt*1e3:8.1f} ms  {sim_steps/dt/1e9:6.3f} G sim-steps/s  "
              f"=> {sim_steps/dt/1.001e7:7.1f} sims/s   finite={np.isfinite(fin).all()} meanE={fin[0].mean():.4f}")
