"""The oracle against the reference's own outputs (tests/golden/, made by make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import bold_oracle, philox, wc_oracle


def _noise(seed, steps, N=90):
    # numba's seeded stream == RandomState(seed).normal(0, sqdtD, (steps, N))  (asserted in make_golden.py)
    return np.random.RandomState(int(seed)).normal(0, 0.2, size=(steps, N))


@pytest.mark.parametrize("case", ["homo", "map"])
def test_wc_numpy_and_c_match_reference_run(case, aal90, oracle_lib):
    g = load_golden(f"wc_short_{case}.npz")
    n1, n2, n3, nrec = [int(v) for v in g["n"]]
    p = wc_oracle.params(P=float(g["P"]), rhoE=float(g["rhoE"]))
    nz = _noise(g["seed"], n1 + n2 + n3)
    Yc = oracle_lib.wc_run(aal90["SC"], g["G"], g["sigmaE"], n1, n2, n3, nrec, noise=nz, p=p)
    Yn = wc_oracle.run(aal90["SC"], g["G"], g["sigmaE"], n1, n2, n3, nrec, noise=nz, p=p)
    ref = g["Y"]
    rows = g["rows"]
    # 1.5 s of simulated time: summation-order differences stay ~1e-12 (SURVEY item 4)
    assert np.max(np.abs(Yc[rows] - ref) / np.abs(ref)) < 1e-9
    assert np.max(np.abs(Yn[rows] - ref) / np.abs(ref)) < 1e-9


def test_chain_matches_reference(aal90, oracle_lib):
    g = load_golden("chain_homo.npz")
    n1, n2, n3, nrec = [int(v) for v in g["n"]]
    p = wc_oracle.params(P=float(g["P"]), rhoE=float(g["rhoE"]))
    nz = _noise(g["seed"], n1 + n2 + n3)
    E_t = oracle_lib.wc_run(aal90["SC"], float(g["G"]), float(g["sigmaE"]), n1, n2, n3, nrec, noise=nz, p=p, want="E")
    ref = g["E_rows"]
    assert np.max(np.abs(E_t[::100] - ref) / np.abs(ref)) < 1e-6      # 20 s of chaos, FP64: order effects ~1e-10
    bold = bold_oracle.sim_bold(E_t, 90, int(g["BOLD_downsamp"]))
    assert bold.shape == g["BOLD"].shape
    # the band-pass is ill-conditioned: 1e-10 input differences come out ~1e-8 relative
    assert np.max(np.abs(bold - g["BOLD"])) < 1e-6 * np.max(np.abs(g["BOLD"]))
    FC = bold_oracle.fc(bold)
    assert np.max(np.abs(FC - g["FC"])) < 1e-6
    for k, s in enumerate(("W", "N1", "N2", "N3")):
        m = bold_oracle.get_all_metrics(FC, aal90[s])
        assert np.allclose(m, g["gof"][k], atol=1e-6)
    sync, meta = bold_oracle.kuramoto(bold)
    assert abs(sync - float(g["sync"])) < 1e-6 and abs(meta - float(g["meta"])) < 1e-6
    assert bold_oracle.welch_peak(E_t) == float(g["peakfreq"])
    assert abs(FC.mean() - float(g["mean"])) < 1e-7


def test_bold_c_equals_numpy(oracle_lib):
    rng = np.random.default_rng(0)
    rE = 0.15 + 0.1 * rng.random((400, 7))
    assert np.allclose(oracle_lib.bold_sim(rE, 0.04), bold_oracle.bold_sim(rE, dt=0.04), rtol=1e-12, atol=1e-15)


def test_philox_kat_and_c_equals_numpy(oracle_lib):
    # Random123 known-answer vectors (kat_vectors) for philox4x32 with 10 rounds and with 7, the stream's round count
    ctrs = [((0, 0, 0, 0), (0, 0)), ((0xffffffff,) * 4, (0xffffffff,) * 2),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))]
    kat = {10: [(0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd),
                (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)],
           7: [(0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48), (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662),
               (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)]}
    assert philox.ROUNDS == 7
    for rounds, exps in kat.items():
        for (ctr, key), exp in zip(ctrs, exps):
            out = philox.philox4x32(*[np.uint32(c) for c in ctr], *key, rounds=rounds)
            assert tuple(int(o) for o in out) == exp
    z = philox.normals(12345, np.array([7, 8], dtype=np.uint64), 99, 90)
    assert np.array_equal(z[1], oracle_lib.philox_normals(12345, 8, 99, 90))
    zz = philox.normals(1, np.arange(100, dtype=np.uint64)[:, None], np.arange(300, dtype=np.uint64)[None, :], 90)
    assert abs(zz.mean()) < 3e-3 and abs(zz.std() - 1) < 3e-3 and abs((zz ** 4).mean() - 3) < 0.03


def test_wc_philox_numpy_equals_c(aal90, oracle_lib):
    p = wc_oracle.params(P=0.4, rhoE=0.18)
    Yn = wc_oracle.run(aal90["SC"], 0.16, 7.68, 50, 100, 200, seed=5, streams=[3, 9], p=p)
    Yc = oracle_lib.wc_run(aal90["SC"], 0.16, 7.68, 50, 100, 200, seed=5, stream=9, p=p)
    assert np.allclose(Yn[1], Yc, rtol=1e-12, atol=0)


def test_committed_tables_statistics():
    s = load_golden("sweep_cell_stats.npz")
    cols = list(s["cols"])
    i, j = 5, 10                              # (delta_G, delta_sigma) = (0, 0)
    assert abs(float(s["delta_G"][i])) < 1e-9 and abs(float(s["delta_sigma"][j])) < 1e-9
    assert abs(s["homo_mean"][i, j, cols.index("corrW")] - 0.474) < 0.005
    assert abs(s["homo_mean"][i, j, cols.index("eW")] - 8.39) < 0.01
