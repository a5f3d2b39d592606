"""Output-row format (whole_sweep_both.py:112-116) and the consumers' reduction (heatmaps.py:28-58)."""
import numpy as np

from nremmodfc_b200 import table


def _fake_out(B, rng):
    return {"gof": rng.random((B, 4, 4)) + 0.1, "sync": rng.random(B), "meta": rng.random(B), "mean": rng.random(B)}


def test_row_format_and_collapse(tmp_path):
    rng = np.random.default_rng(0)
    out = _fake_out(6, rng)
    rows = table.rows_from_sweep(out, rank=3, seeds=[0, 0, 1, 1, 2, 2], delta_G=[-0.1, 0.02] * 3, delta_sigma=[0.0, -0.04] * 3,
                                 peakfreq=np.full(6, 4.875))
    assert rows.shape == (6, 20)
    p = tmp_path / "rank3"
    table.write_rank_file(p, rows[:3])
    table.write_rank_file(p, rows[3:])                              # appends without a second header
    lines = open(p).read().splitlines()
    assert lines[0].split("\t") == table.COLUMNS and len(lines) == 7
    f = lines[1].split("\t")
    assert f[0] == "3" and f[1] == "0" and f[2] == "-0.1000" and f[3] == "0.0000" and f[-1] == "4.8750"
    assert f[4] == f"{out['gof'][0, 0, 2]:.4f}" and f[8] == f"{out['gof'][0, 0, 0]:.4f}" and f[12] == f"{out['gof'][0, 0, 1]:.4f}"
    c = tmp_path / "collapsed.txt"
    table.collapse([p], c)
    tab = table.read_table(c)
    assert list(tab) == table.COLUMNS and len(tab["seed"]) == 6
    assert np.allclose(tab["eW"], np.round(out["gof"][:, 0, 1], 4))


def test_euccorr_optima_recovers_planted_minimum(tmp_path):
    rng = np.random.default_rng(1)
    dG = np.round(np.linspace(-0.1, 0.3, 20, endpoint=False), 4)
    dS = np.round(np.linspace(-0.2, 0.2, 20, endpoint=False), 4)
    g, s = np.meshgrid(dG, dS, indexing="ij")
    g, s = np.tile(g.ravel(), 5), np.tile(s.ravel(), 5)
    tab = {"delta_G": g, "delta_sigma": s}
    for k, st in enumerate(table.STATES):
        target = (dG[5 + k], dS[10 - k])
        tab[f"e{st}"] = 8 + 40 * ((g - target[0]) ** 2 + (s - target[1]) ** 2) + 0.002 * rng.normal(size=g.size)
        tab[f"corr{st}"] = np.full(g.size, 0.5)
    opt = table.euccorr_optima(tab)
    for k, st in enumerate(table.STATES):
        assert opt[st][:2] == (float(dG[5 + k]), float(dS[10 - k]))
