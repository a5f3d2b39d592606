"""Fused sweep beyond 128 nodes (BASELINE configs[4] shape: random 1000-node SC): integrate with the large-connectome kernel and run the
BOLD -> filter -> FC -> GoF -> Kuramoto tail of the plan; prints the time of the integration and of the tail.

    python tools/big_sweep.py [N] [B] [seconds of recording] [kernel]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nremmodfc_b200 import ops, sweep  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rec_s = float(sys.argv[3]) if len(sys.argv) > 3 else 40.0
kernel = sys.argv[4] if len(sys.argv) > 4 else "auto"
rng = np.random.default_rng(5)
SC = rng.uniform(size=(N, N))
np.fill_diagonal(SC, 0.0)
SC *= 2.5 / SC.sum(axis=1).mean()
emp = np.stack([np.corrcoef(rng.normal(size=(N, 300)) + rng.normal(size=(1, 300))) for _ in range(4)])
dt = 1e-4
p = ops.make_params(N, int(1 / dt), int(2 / dt), int(rec_s / dt), P=0.4, rhoE=0.18, seed=1)
dG, dS = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
plan = sweep.SweepPlan(p, B, n_maps=1, K=4, kernel=kernel, bold_f32=True)
t0 = time.perf_counter()
plan.begin(SC, np.full(B, 0.16), dG, np.full(B, 7.68), dS, np.arange(B, dtype=np.uint64) + 1)
plan.advance()
torch.cuda.synchronize()
t1 = time.perf_counter()
out = plan.finish(emp)
torch.cuda.synchronize()
t2 = time.perf_counter()
steps = p.n1 + p.n2 + p.n3
print(json.dumps({"N": N, "B": B, "kernel": plan.kernel_name, "euler_steps": steps, "J": plan.J, "plan_GB": plan.device_bytes / 1e9,
                  "integrate_s": t1 - t0, "us_per_step": (t1 - t0) / steps * 1e6, "tail_s": t2 - t1,
                  "gof_finite": bool(np.isfinite(out["gof"]).all()), "mean_corr_vs_target0": float(out["gof"][:, 0, 0].mean()),
                  "mean_fc": float(out["mean"].mean()), "sync": float(out["sync"].mean()), "meta": float(out["meta"].mean())}))
plan.close()
