"""Developer tool: the large-connectome integrator under two builds of the library (NREM_LIB_PATH), same inputs: are the outputs
bit-identical, and how long does a step take?

    python tools/lib_ab.py <libA.so> <libB.so> [N] [B] [steps] [kernel]
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from nremmodfc_b200 import ops
N, B, steps, kernel, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
rng = np.random.default_rng(5)
SC = rng.uniform(size=(N, N)); np.fill_diagonal(SC, 0.0); SC *= 2.5 / SC.sum(axis=1).mean()
dG, dS = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
mG, mS = rng.uniform(0.5, 1.5, N), rng.uniform(0.8, 1.2, N)
p = ops.make_params(N, steps // 4, steps // 4, steps - 2 * (steps // 4), P=0.4, rhoE=0.18, seed=1)
for rep in range(2):
    E, fin = ops.big_integrate_f32(p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), dS, mG, mS, kernel=kernel)
    ms = ops.last_integrate_ms()
np.savez(out, E=E, fin=fin, us=ms * 1e3 / steps)
'''


def main():
    la, lb = sys.argv[1], sys.argv[2]
    N = sys.argv[3] if len(sys.argv) > 3 else "1000"
    B = sys.argv[4] if len(sys.argv) > 4 else "4096"
    steps = sys.argv[5] if len(sys.argv) > 5 else "3000"
    kernel = sys.argv[6] if len(sys.argv) > 6 else "bf3"
    res = []
    for k, lib in enumerate((la, lb)):
        out = f"/tmp/lib_ab_{k}.npz"
        env = dict(os.environ, NREM_LIB_PATH=os.path.abspath(lib))
        subprocess.check_call([sys.executable, "-c", CHILD % ROOT, N, B, steps, kernel, out], env=env)
        res.append(np.load(out))
    same = bool(np.array_equal(res[0]["E"], res[1]["E"]) and np.array_equal(res[0]["fin"], res[1]["fin"]))
    print(json.dumps({"N": int(N), "B": int(B), "steps": int(steps), "kernel": kernel, "bit_identical": same,
                      "us_per_step_A": float(res[0]["us"]), "us_per_step_B": float(res[1]["us"]),
                      "max_abs_diff_E": float(np.max(np.abs(res[0]["E"] - res[1]["E"])))}))


if __name__ == "__main__":
    main()
