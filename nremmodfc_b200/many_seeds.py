"""run_many_seeds.py on the GPU path: 50 seeds x 4 states at the per-state optima of one modality, FC + HMA
integration / segregation, in the reference's pickle schema.

Replaces the loop of run_many_seeds.py:105-136 (one `wc.run()` + `simBOLD` + `np.corrcoef` + HMA per (seed, state)) by ONE
batched sweep call plus host-side HMA (2 ms per matrix).  The optima are the ones the authors list at
run_many_seeds.py:34-47; `local_G = G + ach_dist*deltaG`, `local_sigmaE = sigmaE + na_dist*deltasigmaE` (:115-116).
"""
import itertools
import pickle

import numpy as np

from . import HMA, ops, sweep

STATES = ("W", "N1", "N2", "N3")
# (G, delta_G, sigmaE, delta_sigmaE) — run_many_seeds.py:34-47
OPTIMALS = {
    "homo": {"W": (0.16, 0.0, 7.68, 0.0), "N1": (0.16, 0.04, 7.68, 0.0), "N2": (0.16, 0.0, 7.68, 0.0), "N3": (0.16, -0.04, 7.68, 0.04)},
    "map": {"W": (0.16, -0.02, 7.68, -0.02), "N1": (0.16, 0.18, 7.68, -0.02), "N2": (0.16, 0.02, 7.68, -0.04), "N3": (0.16, 0.02, 7.68, -0.12)},
    "shuf": {"W": (0.16, 0.0, 7.68, 0.0), "N1": (0.16, 0.0, 7.68, 0.04), "N2": (0.16, 0.0, 7.68, 0.0), "N3": (0.16, 0.0, 7.68, -0.04)},
}


def run_many_seeds(p, CM, emp, ach_dist, na_dist, modality="map", seeds=range(50), optimals=None, **plan_kw):
    """Returns the reference's dict  (seed, state) -> {Hin_sim, Hse_sim, Hin_node_sim, Hse_node_sim, sFC}.

    ach_dist / na_dist are the raw maps (they are normalised to mean 1 here, run_many_seeds.py:60-61); for the
    homogeneous modality pass vectors of ones."""
    opt = (optimals or OPTIMALS)[modality]
    ach = np.asarray(ach_dist, dtype=np.float64)
    na = np.asarray(na_dist, dtype=np.float64)
    ach, na = ach / ach.mean(), na / na.mean()
    sims = list(itertools.product(list(seeds), STATES))                       # run_many_seeds.py:101
    G0 = np.array([opt[st][0] for _, st in sims])
    dG = np.array([opt[st][1] for _, st in sims])
    s0 = np.array([opt[st][2] for _, st in sims])
    ds = np.array([opt[st][3] for _, st in sims])
    streams = np.array([(int(sd) << 8) | STATES.index(st) for sd, st in sims], dtype=np.uint64)
    out = sweep.sweep_gof(p, CM, emp, G0, dG, s0, ds, streams, mapG=ach[None], mapS=na[None], want_fc=True, **plan_kw)
    save = {}
    for k, (sd, st) in enumerate(sims):
        sFC = out["fc"][k]
        num, size, _ = HMA.Functional_HP(sFC)                                 # clips sFC in place, as the reference does (:130,136)
        Hin, Hse = HMA.Balance(sFC, num, size)
        Hin_n, Hse_n = HMA.nodal_measures(sFC, num, size)
        save[(sd, st)] = {"Hin_sim": Hin, "Hse_sim": Hse, "Hin_node_sim": Hin_n, "Hse_node_sim": Hse_n, "sFC": sFC}
    return save


def dump(save, path):
    """run_many_seeds.py:144-146."""
    with open(path, "wb") as f:
        pickle.dump(save, f)
