"""BASELINE configs[4]: scaled synthetic connectome (1000-node random SC, 4096 instances) on the per-step tcgen05 kernel.

    python tools/big_bench.py [N] [B] [steps] [kernel]

Prints Euler steps/s of the whole batch, the algorithmic coupling rate 2*B*N^2 flop/step (x3 executed for 3xTF32) and
node-updates/s, from CUDA events around the step launches (ops.last_integrate_ms())."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nremmodfc_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
kernel = sys.argv[4] if len(sys.argv) > 4 else "tc3"
rng = np.random.default_rng(5)
SC = rng.uniform(size=(N, N))
np.fill_diagonal(SC, 0.0)
SC *= 2.5 / SC.sum(axis=1).mean()
dGv, dSv = np.linspace(-0.1, 0.3, 20, endpoint=False), np.linspace(-0.2, 0.2, 20, endpoint=False)
dG, dS = dGv[rng.integers(0, 20, B)], dSv[rng.integers(0, 20, B)]
n1 = steps // 10
p = ops.make_params(N, n1, steps - n1, 0, P=0.4, rhoE=0.18, seed=1)
for rep in range(2):                     # first call warms up (module load, allocation)
    _, fin = ops.big_integrate_f32(p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), dS, kernel=kernel, record=False)
    ms = ops.last_integrate_ms()
us = ms * 1e3 / steps
print(json.dumps({"N": N, "B": B, "steps": steps, "kernel": kernel, "us_per_step": us, "steps_per_s": 1e6 / us,
                  "coupling_TFLOPs_algorithmic": 2.0 * B * N * N / us / 1e6, "executed_tf32_TFLOPs": (3 if kernel != "tc" else 1) * 2.0 * B * N * N / us / 1e6,
                  "node_updates_per_s": B * N / us * 1e6, "finite": bool(np.isfinite(fin).all()),
                  "mean_E": float(fin[0].mean()), "mean_a": float(fin[2].mean())}))
