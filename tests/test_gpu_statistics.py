"""Full-length statistical parity (north_star check 2): with in-kernel Philox noise, the GoF statistics over
seeds must fall inside the reference's seed-to-seed confidence interval.

The reference's committed sweep tables (output/*.txt -> tests/golden/sweep_cell_stats.npz: per-cell mean, sd, n
over 50 seeds) are the only pin of the whole pipeline including BOLDModel and SSIM.
Nine 128-simulation tiles (three per modality: homogeneous, NA/ACh maps, shuffled maps) run the full
1 + 400 + 600 s horizon concurrently on nine SMs: ~1 minute of GPU time.
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

CELLS = [(0.0, 0.0), (0.18, -0.02), (-0.1, -0.2), (0.02, -0.12)]      # (delta_G, delta_sigma)
SEEDS_PER_CELL = 96
COLS = ["ssimW", "ssimN1", "ssimN2", "ssimN3", "corrW", "corrN1", "corrN2", "corrN3", "eW", "eN1", "eN2", "eN3", "mean"]


def test_gof_statistics_match_committed_tables(aal90):
    from nremmodfc_b200 import ops, sweep
    stats = load_golden("sweep_cell_stats.npz")
    cols = list(stats["cols"])
    dGv, dSv = stats["delta_G"], stats["delta_sigma"]
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    ones = np.ones(90)
    norm = lambda m: m / m.mean()                                       # whole_sweep_both_maps.py:54,62
    mapG = np.stack([ones, norm(aal90["map_ACh"]), norm(aal90["map_ACh_shuf"])])
    mapS = np.stack([ones, norm(aal90["map_NA"]), norm(aal90["map_NA_shuf"])])
    mods = ["homo", "map", "shuf"]
    dG, dS, mid, seed = [], [], [], []
    for m in range(3):
        for (g, s) in CELLS:
            for k in range(SEEDS_PER_CELL):
                dG.append(g); dS.append(s); mid.append(m); seed.append(k)
    dG, dS, mid, seed = map(np.asarray, (dG, dS, mid, seed))
    B = len(dG)
    assert B == 9 * 128
    streams = (np.arange(B, dtype=np.uint64) << np.uint64(8)) | seed.astype(np.uint64)
    p = ops.make_params(90, 10_000, 4_000_000, 6_000_000, P=0.4, rhoE=0.18, seed=424242)   # whole_sweep_both.py:39-50
    out = sweep.sweep_gof(p, aal90["SC"], emp, np.full(B, 0.16), dG, np.full(B, 7.68), dS, streams, mapG=mapG, mapS=mapS,
                          map_id=mid.astype(np.int32), kernel="auto", bold_f32=True)
    gof = out["gof"]                                                     # [B, 4 states, (corr, euc, ssim, new)]
    table = np.concatenate([gof[:, :, 2], gof[:, :, 0], gof[:, :, 1], out["mean"][:, None]], axis=1)   # COLS order
    assert np.isfinite(table).all()
    report, worst = [], 0.0
    for m, mod in enumerate(mods):
        for (g, s) in CELLS:
            i, j = int(np.argmin(np.abs(dGv - g))), int(np.argmin(np.abs(dSv - s)))
            sel = (mid == m) & (dG == g) & (dS == s)
            mine, n_mine = table[sel].mean(0), sel.sum()
            sd_mine = table[sel].std(0, ddof=1)
            for c, name in enumerate(COLS):
                k = cols.index(name)
                ref, sd_ref, n_ref = stats[f"{mod}_mean"][i, j, k], stats[f"{mod}_sd"][i, j, k], stats[f"{mod}_n"][i, j]
                se = np.sqrt(sd_ref ** 2 / n_ref + sd_mine[c] ** 2 / n_mine)
                z = abs(mine[c] - ref) / (se + 1e-3)                    # 1e-3: the tables are rounded to 4 decimals
                worst = max(worst, z)
                report.append((z, mod, g, s, name, float(mine[c]), float(ref), float(sd_ref)))
    report.sort(reverse=True)
    for r in report[:8]:
        print("z=%.2f %s dG=%.2f dS=%.2f %s ours=%.4f ref=%.4f (sd %.4f)" % r)
    # 156 comparisons: |z| < 4.5 everywhere (P ~ 1e-3 for a false alarm), and no systematic bias
    assert worst < 4.5
    assert np.mean([r[0] for r in report]) < 1.6
