"""Full-length statistical parity (north_star check 2): with in-kernel Philox noise, the GoF statistics over
seeds must fall inside the reference's seed-to-seed confidence interval.

The reference's committed sweep tables (output/*.txt -> tests/golden/sweep_cell_stats.npz: per-cell mean, sd, n
over 50 seeds) are the only pin of the whole pipeline including BOLDModel and SSIM.
One wave of the GPU (147 tiles of 128 simulations on 148 SMs, the wall time of a single tile): 3 modalities
(homogeneous, NA/ACh maps, shuffled maps) x 64 cells (an 8 x 8 sub-grid of the 20 x 20 table) x 98 seeds = 18 816
simulations at the full 1 + 400 + 600 s horizon, and ALL 16 value columns of the reference's output row
(whole_sweep_both.py:112-116) including sync / meta (utils.kuramoto) and the Welch peak frequency.
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

GRID = [0, 3, 5, 8, 10, 13, 15, 18]                                   # indices into the table's 20 delta_G / delta_sigma values
SEEDS_PER_CELL = 98
COLS = ["ssimW", "ssimN1", "ssimN2", "ssimN3", "corrW", "corrN1", "corrN2", "corrN3", "eW", "eN1", "eN2", "eN3", "sync", "meta",
        "mean", "peakfreq"]


def test_gof_statistics_match_committed_tables(aal90):
    from nremmodfc_b200 import ops, sweep
    stats = load_golden("sweep_cell_stats.npz")
    assert list(stats["cols"]) == COLS
    dGv, dSv = stats["delta_G"], stats["delta_sigma"]
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    ones = np.ones(90)
    norm = lambda m: m / m.mean()                                       # whole_sweep_both_maps.py:54,62
    mapG = np.stack([ones, norm(aal90["map_ACh"]), norm(aal90["map_ACh_shuf"])])
    mapS = np.stack([ones, norm(aal90["map_NA"]), norm(aal90["map_NA_shuf"])])
    mods = ["homo", "map", "shuf"]
    ci, cj, mid, seed = [], [], [], []
    for m in range(3):
        for i in GRID:
            for j in GRID:
                for k in range(SEEDS_PER_CELL):
                    ci.append(i); cj.append(j); mid.append(m); seed.append(k)
    ci, cj, mid, seed = map(np.asarray, (ci, cj, mid, seed))
    B = len(ci)
    assert B == 147 * 128                                                # one wave; every 128-tile holds one modality
    dG, dS = dGv[ci], dSv[cj]
    streams = (np.arange(B, dtype=np.uint64) << np.uint64(8)) | seed.astype(np.uint64)
    p = ops.make_params(90, 10_000, 4_000_000, 6_000_000, P=0.4, rhoE=0.18, seed=424242)   # whole_sweep_both.py:39-50
    out = sweep.sweep_gof(p, aal90["SC"], emp, np.full(B, 0.16), dG, np.full(B, 7.68), dS, streams, mapG=mapG, mapS=mapS,
                          map_id=mid.astype(np.int32), kernel="auto", bold_f32=True, peakfreq=True)
    gof = out["gof"]                                                     # [B, 4 states, (corr, euc, ssim, new)]
    table = np.concatenate([gof[:, :, 2], gof[:, :, 0], gof[:, :, 1], out["sync"][:, None], out["meta"][:, None],
                            out["mean"][:, None], out["peakfreq"][:, None]], axis=1)               # COLS order
    assert table.shape == (B, 16) and np.isfinite(table).all()
    report, zs = [], {c: [] for c in COLS}
    for m, mod in enumerate(mods):
        for i in GRID:
            for j in GRID:
                sel = (mid == m) & (ci == i) & (cj == j)
                mine, n_mine = table[sel].mean(0), sel.sum()
                sd_mine = table[sel].std(0, ddof=1)
                for c, name in enumerate(COLS):
                    ref, sd_ref, n_ref = stats[f"{mod}_mean"][i, j, c], stats[f"{mod}_sd"][i, j, c], stats[f"{mod}_n"][i, j]
                    # the tables are rounded to 4 decimals; the peak frequency is quantised at fs/nperseg = 0.125 Hz
                    se = np.sqrt(sd_ref ** 2 / n_ref + sd_mine[c] ** 2 / n_mine) + (0.07 if name == "peakfreq" else 1e-3)
                    z = (mine[c] - ref) / se
                    zs[name].append(z)
                    report.append((abs(z), mod, dGv[i], dSv[j], name, float(mine[c]), float(ref), float(sd_ref)))
    report.sort(reverse=True)
    for r in report[:10]:
        print("|z|=%.2f %s dG=%.2f dS=%.2f %s ours=%.4f ref=%.4f (sd %.4f)" % r)
    allz = np.concatenate([np.asarray(v) for v in zs.values()])
    print("columns: " + ", ".join(f"{c} mean z {np.mean(zs[c]):+.2f} sd {np.std(zs[c]):.2f}" for c in COLS))
    print(f"{len(allz)} comparisons: mean z {allz.mean():+.3f}, sd(z) {allz.std():.3f}, |z|<3 in {np.mean(np.abs(allz) < 3) * 100:.2f} %, max |z| {np.abs(allz).max():.2f}")
    # 3072 comparisons (192 cells x 16 columns).  Our cell means use 98 seeds, the reference's 50, so a correct pipeline gives
    # z ~ N(0, <= 1) (the additive rounding terms make it slightly under-dispersed; columns of one cell are correlated).
    assert 0.6 < allz.std() < 1.3
    assert abs(allz.mean()) < 0.15
    assert np.mean(np.abs(allz) < 3) > 0.99 and np.abs(allz).max() < 5.5
    for c in COLS:                                                        # no column is systematically off
        assert abs(np.mean(zs[c])) < 0.45, (c, np.mean(zs[c]))
