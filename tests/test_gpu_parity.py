"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (north_star): trajectories within 1e-6 relative and FC within 1e-4 absolute of the
reference's float64 path when the reference's own noise stream is injected.  The network is
chaotic (one e-fold per ~1.5 s), so trajectory parity is only definable over a
short horizon (<= 20 s of simulated time here); full-length parity is statistical
(test_gpu_statistics.py).
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _noise(seed, steps, N=90):
    return np.random.RandomState(int(seed)).normal(0, 0.2, size=(steps, N))


@pytest.fixture(scope="module")
def wcmod(aal90):
    """The drop-in module configured the way whole_sweep_both.py:39-41 does."""
    import nremmodfc_b200.netwWilsonCowanPlastic as wc
    wc.P = 0.4
    wc.rhoE = 0.18
    wc.CM = aal90["SC"]
    return wc


def _set_times(wc, t1, t2, tstop):
    wc.tTrans1, wc.tTrans2, wc.tstop = t1, t2, tstop             # whole_sweep_both.py:43-50
    wc.timeTrans1 = np.arange(0, t1, wc.dtSim)
    wc.timeTrans2 = np.arange(0, t2, wc.dtSim)
    wc.timeSim = np.arange(0, tstop, wc.dtSim)
    wc.time = np.arange(0, tstop, wc.dt)


@pytest.mark.parametrize("case", ["homo", "map"])
def test_run_matches_reference_trajectory(case, wcmod):
    wc = wcmod
    g = load_golden(f"wc_short_{case}.npz")
    _set_times(wc, 0.05, 0.45, 1.0)
    n1, n2, n3, nrec = [int(v) for v in g["n"]]
    assert (len(wc.timeTrans1), len(wc.timeTrans2), len(wc.timeSim), len(wc.time)) == (n1, n2, n3, nrec)
    wc.G = g["G"] if g["G"].ndim else float(g["G"])
    wc.sigmaE = g["sigmaE"] if g["sigmaE"].ndim else float(g["sigmaE"])
    wc.noise = _noise(g["seed"], n1 + n2 + n3)
    wc.run.recompile()
    Y = wc.run()
    wc.noise = None
    assert Y.shape == (nrec, 3, 90) and Y.dtype == np.float64 and Y.flags["C_CONTIGUOUS"]
    ref = g["Y"]
    assert np.max(np.abs(Y[g["rows"]] - ref) / np.abs(ref)) < 1e-6          # north_star: 1e-6 relative


def test_chain_matches_reference(wcmod, aal90):
    """run() -> simBOLD() -> corrcoef -> get_all_metrics on the reference's 1+2+17 s golden chain."""
    from nremmodfc_b200 import ops, utils
    wc = wcmod
    g = load_golden("chain_homo.npz")
    n1, n2, n3, nrec = [int(v) for v in g["n"]]
    _set_times(wc, 1, 2, 17)
    wc.G, wc.sigmaE = float(g["G"]), float(g["sigmaE"])
    wc.noise = _noise(g["seed"], n1 + n2 + n3)
    wc.run.recompile()
    tray = wc.run()
    wc.noise = None
    E_t = tray[:, 0, :]
    assert np.max(np.abs(E_t[::100] - g["E_rows"]) / np.abs(g["E_rows"])) < 1e-6
    assert np.max(np.abs(tray[-1] - g["final"]) / np.abs(g["final"])) < 1e-6
    BOLD = wc.simBOLD(E_t, nnodes=90, BOLD_downsamp=int(g["BOLD_downsamp"]))
    assert BOLD.shape == g["BOLD"].shape
    assert np.max(np.abs(BOLD - g["BOLD"])) < 1e-6 * np.max(np.abs(g["BOLD"]))
    sFC = ops.fc(BOLD)
    assert np.max(np.abs(sFC - g["FC"])) < 1e-4                               # north_star: FC 1e-4 absolute
    assert np.max(np.abs(sFC - g["FC"])) < 1e-6                               # what we actually achieve
    for k, s in enumerate(("W", "N1", "N2", "N3")):
        m = utils.get_all_metrics(sFC, aal90[s], data_range=1)
        assert np.allclose(m, g["gof"][k], atol=1e-6)


def test_wc_f64_philox_matches_oracle(aal90, oracle_lib):
    from nremmodfc_b200 import ops
    from oracle import wc_oracle
    n1, n2, n3 = 60, 140, 400
    p = ops.make_params(90, n1, n2, n3, P=0.4, rhoE=0.18, seed=0xABCDEF0123)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    streams = np.array([0, 5, 2 ** 40 + 3], dtype=np.uint64)
    G = np.stack([np.full(90, 0.16), 0.16 + 0.1 * aal90["map_ACh"] / aal90["map_ACh"].mean(), np.full(90, 0.3)])
    sg = np.stack([np.full(90, 7.68), 7.68 - 0.1 * aal90["map_NA"] / aal90["map_NA"].mean(), np.full(90, 7.5)])
    Y, fin = ops.wc_run(p, aal90["SC"], G, sg, B=3, streams=streams)
    for b in range(3):
        Yo = oracle_lib.wc_run(aal90["SC"], G[b], sg[b], n1, n2, n3, seed=0xABCDEF0123, stream=int(streams[b]), p=po)
        assert np.max(np.abs(Y[b] - Yo) / np.abs(Yo)) < 1e-9
        fo = oracle_lib.wc_run(aal90["SC"], G[b], sg[b], n1, n2, n3, seed=0xABCDEF0123, stream=int(streams[b]), p=po, want="final")
        assert np.max(np.abs(fin[b] - fo) / np.abs(fo)) < 1e-9


def test_wc_f64_noise_warps_are_bit_identical(aal90, monkeypatch):
    """The float64 run() kernel draws the Philox noise of step t + 1 in a second set of warps while the first set integrates step t
    (csrc/wc_f64.cuh, NOISE_WARPS): same arithmetic per draw, so Y_t and the final state equal the single-set kernel bit for bit —
    for the shared-memory SC path (N = 90), a node count that is not a multiple of 32 with per-node vectors, and the transposed
    global-memory SC path (N = 200)."""
    from nremmodfc_b200 import ops
    rng = np.random.default_rng(3)
    for N in (90, 45, 200):
        SC = aal90["SC"] if N == 90 else _random_sc(N, N)
        p = ops.make_params(N, 40, 60, 120, P=0.4, rhoE=0.18, seed=77)
        G = rng.uniform(0.1, 0.3, (3, N))
        sg = rng.uniform(7.0, 8.0, (3, N))
        kw = dict(B=3, streams=[5, 6, 2 ** 40 + 7], want_Y=True, node_params={"P": 0.4 + 0.05 * rng.random(N)} if N == 45 else None)
        monkeypatch.setenv("NREM_F64_NOISE_WARPS", "0")
        Y0, f0 = ops.wc_run(p, SC, G, sg, **kw)
        monkeypatch.setenv("NREM_F64_NOISE_WARPS", "1")
        Y1, f1 = ops.wc_run(p, SC, G, sg, **kw)
        assert Y0.shape == (3, 6, 3, N) and np.isfinite(Y0).all()
        assert np.array_equal(Y0, Y1) and np.array_equal(f0, f1), N


def test_wc_derivative_matches_oracle(aal90):
    from nremmodfc_b200 import ops
    from oracle import wc_oracle
    rng = np.random.default_rng(3)
    X = np.stack([rng.random(90) * 0.5, rng.random(90) * 0.5, 2 + rng.random(90)])
    nz = rng.normal(0, 0.2, 90)
    p = ops.make_params(90, 0, 0, 0, P=0.4, rhoE=0.18)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    d = ops.wc_derivative(p, aal90["SC"], X, 0.16, 7.68, noise=nz, tau_ip=2.0)
    dE, dI, da = wc_oracle.derivative(X[0], X[1], X[2], aal90["SC"], 0.16, 7.68, nz, 2.0, po)
    assert np.allclose(d, np.stack([dE, dI, da]), rtol=1e-12, atol=1e-13)


def _synthetic_E(T, N, seed):
    """E_t-like input: positive, band-limited fluctuations around 0.18 plus white noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(T)[:, None] * 0.002
    slow = 0.03 * np.sin(2 * np.pi * (0.02 + 0.05 * rng.random(N))[None, :] * t * 20 + rng.random(N)[None, :] * 6)
    walk = np.cumsum(rng.normal(0, 1, (T, N)), axis=0)
    walk = 0.02 * (walk - walk.mean(0)) / (walk.std(0) + 1e-9)
    return 0.18 + slow + walk + 0.05 * rng.random((T, N))


@pytest.mark.parametrize("T,N,ds,Neq", [(300000, 90, 1000, 2000), (30011, 90, 10, 2000), (2100, 7, 3, 2000), (5000, 33, 97, 0)])
def test_bold_filter_stage_matches_oracle(T, N, ds, Neq, oracle_lib):
    """BOLDModel.Sim + cut/filtfilt/decimate at the reference's full size and at ragged sizes."""
    from nremmodfc_b200 import ops
    from nremmodfc_b200.sweep import bandpass_ba
    from oracle import bold_oracle
    E = _synthetic_E(T, N, seed=T)
    bold = ops.bold_sim(E, 0.04)
    bo = oracle_lib.bold_sim(E, 0.04)
    assert np.max(np.abs(bold - bo)) < 1e-12 * max(1.0, np.max(np.abs(bo)))
    b, a = bandpass_ba(0.04)
    y = ops.filtfilt_decimate(bo, b, a, Neq=Neq, ds=ds)
    yo = bold_oracle.filt_decimate(bo, ds, Neq, 0.04)
    assert y.shape == yo.shape
    assert np.max(np.abs(y - yo)) < 1e-6 * np.max(np.abs(yo))
    if y.shape[0] > 8:
        assert np.max(np.abs(ops.fc(y) - bold_oracle.fc(yo))) < 1e-6


def test_filter_is_linear_and_batched():
    """Size-independent property at the reference's full length: filt(a x + b y) = a filt(x) + b filt(y)."""
    from nremmodfc_b200 import ops
    from nremmodfc_b200.sweep import bandpass_ba
    rng = np.random.default_rng(1)
    x = np.cumsum(rng.normal(size=(2, 300000, 5)), axis=1) * 1e-3
    b, a = bandpass_ba(0.04)
    fx = ops.filtfilt_decimate(x, b, a)
    fz = ops.filtfilt_decimate(2.0 * x[0] - 3.0 * x[1], b, a)
    assert fx.shape == (2, 298, 5)
    assert np.max(np.abs(fz - (2.0 * fx[0] - 3.0 * fx[1]))) < 1e-9 * np.max(np.abs(fz))


def test_short_input_is_rejected():
    from nremmodfc_b200 import ops
    from nremmodfc_b200.sweep import bandpass_ba
    b, a = bandpass_ba(0.04)
    with pytest.raises(ValueError):
        ops.filtfilt_decimate(np.zeros((2010, 3)), b, a)


@pytest.mark.parametrize("N,J", [(90, 298), (90, 11), (17, 64), (128, 40), (129, 298), (300, 298), (1000, 77)])
def test_fc_matches_numpy(N, J):
    from nremmodfc_b200 import ops
    rng = np.random.default_rng(N + J)
    x = rng.normal(size=(3, J, N)) @ (np.eye(N) + 0.3 * rng.normal(size=(N, N)))
    fc = ops.fc(x)
    for b in range(3):
        assert np.max(np.abs(fc[b] - np.corrcoef(x[b].T))) < 1e-12
    assert np.array_equal(fc, np.transpose(fc, (0, 2, 1)))


@pytest.mark.parametrize("N", [90, 30, 7, 128, 200, 419])
def test_gof_matches_oracle(N, aal90):
    from nremmodfc_b200 import ops
    from oracle import bold_oracle
    rng = np.random.default_rng(N)
    if N == 90:
        emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    else:
        emp = np.stack([np.corrcoef(rng.normal(size=(N, 50)) + rng.normal(size=(1, 50))) for _ in range(2)])
    sims = np.stack([np.corrcoef(rng.normal(size=(N, 40)) + 0.7 * rng.normal(size=(1, 40))) for _ in range(5)])
    g, m = ops.gof(sims, emp)
    for b in range(5):
        assert abs(m[b] - sims[b].mean()) < 1e-13
        for k in range(emp.shape[0]):
            assert np.allclose(g[b, k], bold_oracle.get_all_metrics(sims[b], emp[k]), rtol=1e-10, atol=1e-12)
    # identity: a matrix against itself
    g2, _ = ops.gof(emp[0], emp[0])
    assert np.allclose(g2[0, 0], [1.0, 0.0, 1.0, 0.0], atol=1e-12)


@pytest.mark.parametrize("kernel", ["fma", "tc", "tc3"])
def test_integrator_f32_short_horizon_vs_oracle(kernel, aal90, oracle_lib):
    """The float32 production integrator with in-kernel Philox against the float64 oracle on the SAME
    counter-based stream: a few hundred steps, so float32 rounding (not chaos) bounds the error."""
    from nremmodfc_b200 import ops
    from oracle import wc_oracle
    n1, n2, n3 = 50, 100, 200
    seed = 77
    p = ops.make_params(90, n1, n2, n3, P=0.4, rhoE=0.18, seed=seed)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    B = 131                                                    # ragged: 2 tiles, the second almost empty
    rng = np.random.default_rng(0)
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    mG = np.stack([np.ones(90), aal90["map_ACh"] / aal90["map_ACh"].mean()])
    mS = np.stack([np.ones(90), aal90["map_NA"] / aal90["map_NA"].mean()])
    map_id = np.r_[np.zeros(128, np.int32), np.ones(3, np.int32)]
    streams = rng.integers(0, 2 ** 62, B).astype(np.uint64)
    E, fin = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, map_id, streams, kernel=kernel)
    assert E.shape == (10, 90, B) and fin.shape == (3, 90, B)
    for b in (0, 1, 31, 32, 77, 127, 128, 130):
        m = map_id[b]
        Yo, fo = wc_oracle.run(aal90["SC"], 0.16 + dG[b] * mG[m], 7.68 + ds[b] * mS[m], n1, n2, n3, seed=seed,
                               streams=[int(streams[b])], p=po, return_final=True)
        tol = 5e-3 if kernel == "tc" else 2e-4           # "tc" rounds the coupling operands to TF32 (2^-11)
        assert np.max(np.abs(E[:, :, b] - Yo[0, :, 0, :]) / np.abs(Yo[0, :, 0, :])) < tol
        assert np.max(np.abs(fin[:, :, b] - fo[0]) / np.abs(fo[0])) < tol


@pytest.mark.parametrize("kernel", ["fma", "tc", "tc3"])
def test_integrator_is_deterministic_and_layout_independent(kernel, aal90):
    """Results depend on (seed, stream, parameters) only — not on the position inside the batch."""
    from nremmodfc_b200 import ops
    p = ops.make_params(90, 100, 200, 400, P=0.4, rhoE=0.18, seed=3)
    B = 256
    streams = np.arange(B, dtype=np.uint64)
    dG = np.linspace(-0.1, 0.3, B)
    E1, f1 = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), np.zeros(B), streams=streams, kernel=kernel)
    E2, f2 = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), np.zeros(B), streams=streams, kernel=kernel)
    assert np.array_equal(E1, E2) and np.array_equal(f1, f2)
    perm = np.random.default_rng(0).permutation(B)[:100]
    E3, f3 = ops.integrate_f32(p, aal90["SC"], np.full(100, 0.16), dG[perm], np.full(100, 7.68), np.zeros(100),
                               streams=streams[perm], kernel=kernel)
    assert np.array_equal(E3, E1[:, :, perm]) and np.array_equal(f3, f1[:, :, perm])


@pytest.mark.parametrize("kernel,bold_f32", [("fma", False), ("fma", True), ("tc", True), ("tc3", False)])
def test_sweep_pipeline_vs_oracle(kernel, bold_f32, aal90, oracle_lib):
    """Whole fused pipeline on a shortened run (1 s of recording).

    The band-pass keeps ~1e-3 of a 1 s signal, so FC of such a short run amplifies float32
    trajectory rounding by ~1e3; the pipeline is therefore checked in two legs: (a) the
    integrator's E samples against the float64 oracle on the same Philox stream, (b) the
    BOLD -> filter -> FC -> GoF stages of the sweep against the oracle fed with the SAME samples."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle, wc_oracle
    n1, n2, n3 = 200, 800, 10000
    p = ops.make_params(90, n1, n2, n3, P=0.4, rhoE=0.18, seed=9)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    B = 140
    dG, ds = np.linspace(-0.1, 0.3, B), np.linspace(0.2, -0.2, B)
    streams = np.arange(B, dtype=np.uint64) * 7 + 1
    G0, s0 = np.full(B, 0.16), np.full(B, 7.68)
    out = sweep.sweep_gof(p, aal90["SC"], emp, G0, dG, s0, ds, streams, want_fc=True,
                          kernel=kernel, bold_f32=bold_f32, Neq=100, bold_downsamp=10, chunk_samples=37)
    assert out["gof"].shape == (B, 4, 4) and out["fc"].shape == (B, 90, 90)
    Eg, _ = ops.integrate_f32(p, aal90["SC"], G0, dG, s0, ds, streams=streams, kernel=kernel)
    for k in (0, 63, 127, 128, 139):
        Eo = oracle_lib.wc_run(aal90["SC"], 0.16 + dG[k], 7.68 + ds[k], n1, n2, n3, seed=9, stream=int(streams[k]), p=po, want="E")
        # (a) float32 vs float64 on the same stream.  A float32 numpy emulation of the reference loop already
        # differs from float64 by 2e-5 after 0.1 s and 3e-3 after 1.1 s at G = 0.16 (chaotic transient; on the GPU
        # 5.5e-5 / 1.2e-2 for "fma" and "tc3" alike) and reaches O(1) at the strongly coupled end of the grid, so
        # only the first recorded row (0.1 s) is a meaningful trajectory check here.
        rel = np.abs(Eg[:, :, k] - Eo) / np.maximum(np.abs(Eo), 0.05)
        assert rel[0].max() < (5e-3 if kernel == "tc" else 5e-4)
        E = Eg[:, :, k].astype(np.float64)
        FC = bold_oracle.fc(bold_oracle.filt_decimate(oracle_lib.bold_sim(E, 0.04), 10, 100, 0.04))
        # (b) float64 BOLD state: the stages agree to 1e-6.  float32 BOLD state: this run keeps 1 s of signal, of which the band-pass
        # passes ~1e-3, so float32 rounding of the Balloon-Windkessel state is amplified ~1e3-fold in FC — a property of the
        # shortened test, not of the kernel: at the reference's length the same kernel is held to 1e-4
        # (test_full_length_tail_matches_oracle below).
        assert np.max(np.abs(FC - out["fc"][k])) < (2e-2 if bold_f32 else 1e-6)
        g = np.array([bold_oracle.get_all_metrics(out["fc"][k], emp[j]) for j in range(4)])
        assert np.allclose(g, out["gof"][k], atol=1e-9)
        assert abs(out["mean"][k] - out["fc"][k].mean()) < 1e-12
        sync, meta = bold_oracle.kuramoto(bold_oracle.filt_decimate(oracle_lib.bold_sim(E, 0.04), 10, 100, 0.04))
        tol = 5e-2 if bold_f32 else 1e-6
        assert abs(out["sync"][k] - sync) < tol and abs(out["meta"][k] - meta) < tol


@pytest.mark.parametrize("passes,tol", [(1, 2e-3), (3, 2e-6)])
def test_tcgen05_contraction(passes, tol, aal90):
    """E[128,96] x SC^T through tcgen05.mma/TMEM against float64: TF32 (~2^-11) and 3xTF32 (~2^-21)."""
    from nremmodfc_b200 import ops
    rng = np.random.default_rng(passes)
    E = np.zeros((128, 96), np.float32)
    E[:, :90] = rng.random((128, 90), dtype=np.float32)
    SC = np.zeros((96, 96), np.float32)
    SC[:90, :90] = aal90["SC"]
    out = ops.selftest_tc_coupling(E, SC, passes=passes)
    exact = E.astype(np.float64) @ SC.astype(np.float64).T
    assert np.max(np.abs(out - exact)) < tol * np.max(np.abs(exact))


def test_full_length_tail_matches_oracle(aal90, oracle_lib):
    """Deterministic parity of the kernels the bench times, at the reference's sizes: 300 000 stored samples x 90 nodes,
    Neq = 2000, decimation 1000 (netwWilsonCowanPlastic.py:140-158, whole_sweep_both.py:79-96).

    E samples come from the production integrator itself (float32, full recording phase).  They go through (1) the fused sweep
    (`SweepPlan.run` on the same Philox streams: integrator -> bold_filter_chunk_kernel<float> -> filt_backward -> FC -> GoF ->
    Kuramoto -> Welch, the benched configuration), (2) the stored-samples hook of the plan with float32 and (3) float64
    Balloon-Windkessel state, and are compared with the oracle chain (C Balloon-Windkessel, SciPy filtfilt, np.corrcoef,
    utils.get_all_metrics, utils.kuramoto, scipy.signal.welch) fed with the SAME samples.
    north_star: FC within 1e-4 absolute."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle
    n1, n2, n3 = 10_000, 200_000, 6_000_000                    # shortened transient, FULL recording phase
    p = ops.make_params(90, n1, n2, n3, P=0.4, rhoE=0.18, seed=31)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    B = 2
    G0, s0 = np.full(B, 0.16), np.full(B, 7.68)
    dG, ds = np.array([0.0, 0.18]), np.array([0.0, -0.02])
    streams = np.array([11, 12], dtype=np.uint64)
    Eg, _ = ops.integrate_f32(p, aal90["SC"], G0, dG, s0, ds, streams=streams, kernel="auto")
    assert Eg.shape == (300_000, 90, B)
    plan = sweep.SweepPlan(p, B, bold_f32=True, peakfreq=True)
    runs = {"fused f32": plan.run(aal90["SC"], emp, G0, dG, s0, ds, streams, want_fc=True)}
    plan.begin(aal90["SC"], G0, dG, s0, ds, streams)
    for r0 in range(0, 300_000, 10_000):
        plan.feed_samples(Eg[r0:r0 + 10_000])
    runs["fed f32"] = plan.finish(emp, want_fc=True)
    plan.close()
    plan = sweep.SweepPlan(p, B, bold_f32=False, peakfreq=True)
    plan.begin(aal90["SC"], G0, dG, s0, ds, streams)
    for r0 in range(0, 300_000, 10_000):
        plan.feed_samples(Eg[r0:r0 + 10_000])
    runs["fed f64"] = plan.finish(emp, want_fc=True)
    plan.close()
    assert np.array_equal(runs["fused f32"]["fc"], runs["fed f32"]["fc"])          # same kernels, same samples
    for k in range(B):
        E = Eg[:, :, k].astype(np.float64)
        bold = bold_oracle.filt_decimate(oracle_lib.bold_sim(E, 0.04), 1000, 2000, 0.04)
        assert bold.shape == (298, 90)
        FC = bold_oracle.fc(bold)
        sync, meta = bold_oracle.kuramoto(bold)
        gof = np.array([bold_oracle.get_all_metrics(FC, emp[j]) for j in range(4)])
        peak = bold_oracle.welch_peak(E)
        for name, out in runs.items():
            tol_fc = 1e-6 if name == "fed f64" else 1e-4
            err = np.max(np.abs(out["fc"][k] - FC))
            print(f"{name} sim {k}: max |FC - oracle| = {err:.2e}, sync {abs(out['sync'][k] - sync):.1e}, meta {abs(out['meta'][k] - meta):.1e}")
            assert err < tol_fc, (name, k, err)
            assert abs(out["sync"][k] - sync) < 1e-4 and abs(out["meta"][k] - meta) < 1e-4
            assert np.max(np.abs(out["gof"][k] - gof)) < (1e-5 if name == "fed f64" else 2e-3)      # eucl. distance sums 4005 entries
            assert abs(out["mean"][k] - FC.mean()) < tol_fc
            assert out["peakfreq"][k] == peak


def test_sliced_run_equals_one_call(aal90):
    """nrem_sweep_begin / advance / finish (what bench.py times slice by slice): cutting a run into pieces changes nothing, bit
    for bit, with one stream and with tile-group streams, homogeneous and with maps."""
    import torch
    from nremmodfc_b200 import ops, sweep
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    p = ops.make_params(90, 300, 1100, 5000, P=0.4, rhoE=0.18, seed=17)
    kw = dict(Neq=40, bold_downsamp=5, chunk_samples=20, bold_f32=True, peakfreq=True, welch_nperseg=80)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(2)
    for B, hetero in ((200, False), (200, True), ((sms + 2) * 128, False)):
        dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
        streams = rng.integers(0, 2 ** 60, B).astype(np.uint64)
        maps = dict(mapG=aal90["map_ACh"] / aal90["map_ACh"].mean(), mapS=aal90["map_NA"] / aal90["map_NA"].mean()) if hetero else {}
        plan = sweep.SweepPlan(p, B, **kw)
        one = plan.run(aal90["SC"], emp, 0.16, dG, 7.68, ds, streams, want_fc=True, **maps)
        total = plan.chunks_total
        assert total == 1 + 3 + 13                                   # ceil(300/400) + ceil(1100/400) + ceil(5000/400)
        plan.begin(aal90["SC"], 0.16, dG, 7.68, ds, streams, **maps)
        left, seen = total, []
        for n in (1, 2, 0, 5, 3, 100):
            left = plan.advance(n)
            seen.append(left)
        assert seen == [16, 14, 14, 9, 6, 0]
        two = plan.finish(emp, want_fc=True)
        plan.close()
        for key in ("gof", "fc", "mean", "sync", "meta", "peakfreq"):
            assert np.array_equal(one[key], two[key]), (B, hetero, key)
        assert np.isfinite(one["gof"]).all() and np.isfinite(one["peakfreq"]).all()
    # finishing a run that is not complete is an error, not a wrong table
    from nremmodfc_b200._lib import NremError
    plan = sweep.SweepPlan(p, 4, **kw)
    plan.begin(aal90["SC"], 0.16, np.zeros(4), 7.68, np.zeros(4), np.arange(4, dtype=np.uint64))
    plan.advance(3)
    with pytest.raises(NremError):
        plan.finish(emp)
    with pytest.raises(ValueError):
        plan.begin(aal90["SC"], 0.16, np.zeros(4), 7.68, np.zeros(4), np.arange(3, dtype=np.uint64))       # short streams array
    with pytest.raises(ValueError):
        plan.begin(aal90["SC"], 0.16, np.zeros(4), 7.68, np.zeros(4), np.arange(4, dtype=np.uint64), map_id=np.array([0, 0, 0, 1]))
    plan.close()


def test_sweep_with_more_tiles_than_sms(aal90, monkeypatch):
    """More 128-tiles than SMs switches the host scheduler to independent tile-group streams; the results must
    not depend on it (same simulations in a small single-stream batch give identical numbers), nor on whether the BOLD / filter
    launches run one chunk behind the integrator on their own streams (NREM_K2_OVERLAP, the default) or in line."""
    import torch
    from nremmodfc_b200 import ops, sweep
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = (sms + 3) * 128 - 5
    p = ops.make_params(90, 100, 300, 2400, P=0.4, rhoE=0.18, seed=11)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    rng = np.random.default_rng(5)
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    streams = rng.integers(0, 2 ** 60, B).astype(np.uint64)
    kw = dict(Neq=40, bold_downsamp=5, chunk_samples=50, bold_f32=True, kernel="tc3")   # (auto would give the small batch 16-simulation tiles)
    plan = sweep.SweepPlan(p, B, **kw)
    plan.set_profiling(True)
    big = plan.run(aal90["SC"], emp, np.full(B, 0.16), dG, np.full(B, 7.68), ds, streams)
    assert plan.profile()["tile_groups"] > 1
    plan.close()
    monkeypatch.setenv("NREM_K2_OVERLAP", "0")
    plan = sweep.SweepPlan(p, B, **kw)
    inline = plan.run(aal90["SC"], emp, np.full(B, 0.16), dG, np.full(B, 7.68), ds, streams)
    plan.close()
    monkeypatch.delenv("NREM_K2_OVERLAP")
    for key in ("gof", "mean", "sync", "meta"):
        assert np.array_equal(big[key], inline[key]), key
    pick = np.r_[0:40, B - 300:B]
    small = sweep.sweep_gof(p, aal90["SC"], emp, np.full(len(pick), 0.16), dG[pick], np.full(len(pick), 7.68), ds[pick], streams[pick], **kw)
    assert np.isfinite(big["gof"]).all()
    assert np.array_equal(big["gof"][pick], small["gof"]) and np.array_equal(big["mean"][pick], small["mean"])


def test_driver_loop_through_compat_modules(aal90, tmp_path):
    """The call sequence of whole_sweep_both.py:39-116 (set attributes, recompile, run, simBOLD, corrcoef,
    get_all_metrics, kuramoto, one TSV row) through the compat/ module names, on a shortened horizon."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    np.savetxt(tmp_path / "SC_opti_25julio.txt", aal90["SC"])
    os.makedirs(tmp_path / "empirical", exist_ok=True)
    for s in ("W", "N1", "N2", "N3"):
        np.savetxt(tmp_path / "empirical" / f"mean_mat_{s}_8dic24.txt", aal90[s])
    script = r'''
import numpy as np, os, itertools
import netwWilsonCowanPlastic as wc
import BOLDModel as BD
from scipy import signal
from skimage.metrics import structural_similarity as ssim
import utils
rank=int(os.environ['SLURM_ARRAY_TASK_ID']); threads=int(os.environ['SLURM_ARRAY_TASK_MAX'])+1
struct = np.loadtxt("SC_opti_25julio.txt")
states = ("W","N1","N2","N3")
empFCW,empFCN1,empFCN2,empFCN3 = [np.loadtxt(f"empirical/mean_mat_{s}_8dic24.txt") for s in states]
wc.P = 0.4; wc.rhoE = 0.18; wc.CM = struct
wc.tTrans1=0.05; wc.tTrans2=0.5
wc.timeTrans1=np.arange(0,wc.tTrans1,wc.dtSim); wc.timeTrans2=np.arange(0,wc.tTrans2,wc.dtSim)
tstop = 9; wc.tstop = tstop
wc.timeSim=np.arange(0,tstop,wc.dtSim); wc.time=np.arange(0,tstop,wc.dt)
rows = []
for sim,(seed,dG,dS) in enumerate(itertools.product([0,1],[0.0],[0.0, 0.02])):
    if sim%threads == rank:
        wc.sid = seed
        wc.G = 0.16 + dG; wc.sigmaE = 7.68 + dS; wc.nnodes = 90
        wc.run.recompile()
        tray = wc.run()
        E_t = tray[:,0,:]
        BOLD = wc.simBOLD(E_t,nnodes=90,BOLD_downsamp=100)
        sFC = np.corrcoef(BOLD.T)
        m = [utils.get_all_metrics(sFC,e,data_range=1) for e in (empFCW,empFCN1,empFCN2,empFCN3)]
        sync,meta = utils.kuramoto(BOLD)
        rows.append((seed, dS, m[0][0], m[0][1], ssim(sFC, empFCW, data_range=1), m[0][2], sync, meta, float(np.mean(sFC)), tray.shape, BOLD.shape))
for r in rows: print(repr(r))
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "compat"), root]), SLURM_ARRAY_TASK_ID="0", SLURM_ARRAY_TASK_MAX="0")
    r = subprocess.run([sys.executable, "-c", script], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = [eval(l) for l in r.stdout.strip().splitlines()]
    assert len(rows) == 4
    for row in rows:
        assert row[-2] == (4500, 3, 90) and row[-1] == (25, 90)
        assert abs(row[4] - row[5]) < 1e-12                     # skimage shim == gof kernel's ssim
        assert all(np.isfinite(row[2:9]))
    # sid is a real seed here: same (seed, cell) would reproduce; different seeds differ
    assert rows[0][2] != rows[2][2]


@pytest.mark.parametrize("J,N", [(298, 90), (25, 90), (11, 7), (64, 33)])
def test_kuramoto_matches_oracle(J, N):
    """utils.kuramoto (Hilbert phases via a length-J FFT in SciPy) against the direct circular-convolution kernel."""
    from nremmodfc_b200 import ops, utils
    from oracle import bold_oracle
    rng = np.random.default_rng(J)
    x = np.cumsum(rng.normal(size=(3, J, N)), axis=1) * 1e-4
    sync, meta = ops.kuramoto(x)
    for b in range(3):
        so, mo = bold_oracle.kuramoto(x[b])
        assert abs(sync[b] - so) < 1e-10 and abs(meta[b] - mo) < 1e-10
    assert np.allclose(utils.kuramoto(x[0]), bold_oracle.kuramoto(x[0]), atol=1e-10)


def test_welch_peak_frequency_matches_scipy(aal90):
    """peakfreq column (whole_sweep_both.py:90-95): Welch spectrum with SciPy's defaults on the stored E samples.
    8000 stored samples -> 3 overlapping 4000-sample segments; the oracle runs scipy.signal.welch on the very samples
    the integrator produced (float32), so the arg-max must agree exactly."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle
    n1, n2, n3 = 1000, 20000, 160000
    p = ops.make_params(90, n1, n2, n3, P=0.4, rhoE=0.18, seed=21)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    B = 133
    dG = np.linspace(-0.1, 0.3, B)
    streams = np.arange(B, dtype=np.uint64) + 77
    G0, s0, ds = np.full(B, 0.16), np.full(B, 7.68), np.zeros(B)
    out = sweep.sweep_gof(p, aal90["SC"], emp, G0, dG, s0, ds, streams, Neq=2000, bold_downsamp=100, peakfreq=True)
    assert np.isfinite(out["peakfreq"]).all()
    Eg, _ = ops.integrate_f32(p, aal90["SC"], G0, dG, s0, ds, streams=streams)
    for k in (0, 40, 127, 128, 132):
        assert out["peakfreq"][k] == bold_oracle.welch_peak(Eg[:, :, k].astype(np.float64))
    out2 = sweep.sweep_gof(p, aal90["SC"], emp, G0[:5], dG[:5], s0[:5], ds[:5], streams[:5], Neq=2000, bold_downsamp=100)
    assert np.isnan(out2["peakfreq"]).all()                     # spectrum not requested
    assert np.array_equal(out2["gof"], out["gof"][:5])


def test_config5_large_connectome_f64_path(oracle_lib):
    """BASELINE configs[4]: 1000-node random SC (the module's own placeholder distribution, netwWilsonCowanPlastic.py:64,
    zero diagonal, mean row sum scaled to AAL90's 2.5).  The tcgen05 sweep is limited to N <= 96; the float64 path takes
    any N <= 1024 (SC streamed transposed from L2), checked here against the oracle on the same Philox streams."""
    from nremmodfc_b200 import ops
    from oracle import wc_oracle
    N = 1000
    rng = np.random.default_rng(5)
    SC = rng.uniform(size=(N, N))
    np.fill_diagonal(SC, 0.0)
    SC *= 2.5 / SC.sum(axis=1).mean()
    n1, n2, n3 = 20, 30, 60
    p = ops.make_params(N, n1, n2, n3, P=0.4, rhoE=0.18, seed=99)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    G = np.stack([np.full(N, 0.16), np.linspace(0.0, 0.4, N), np.full(N, 0.3)])
    sg = np.stack([np.full(N, 7.68), np.full(N, 7.5), np.linspace(7.4, 7.9, N)])
    streams = np.array([3, 4, 5], dtype=np.uint64)
    Y, fin = ops.wc_run(p, SC, G, sg, B=3, streams=streams)
    assert Y.shape == (3, 3, 3, N)
    for b in range(3):
        Yo = oracle_lib.wc_run(SC, G[b], sg[b], n1, n2, n3, seed=99, stream=int(streams[b]), p=po)
        fo = oracle_lib.wc_run(SC, G[b], sg[b], n1, n2, n3, seed=99, stream=int(streams[b]), p=po, want="final")
        assert np.max(np.abs(Y[b] - Yo) / np.abs(Yo)) < 1e-9 and np.max(np.abs(fin[b] - fo) / np.abs(fo)) < 1e-9


@pytest.mark.parametrize("N", [96, 68, 30])
def test_integrator_other_parcellation_sizes(N, oracle_lib):
    """The sweep takes any connectome with 7 <= N <= 96 nodes (one 96-wide MMA tile); N = 90 merely gets the 18-node
    loop for its last chunk.  Same check as for AAL90: float32 kernel vs float64 oracle on the same Philox streams."""
    from nremmodfc_b200 import ops
    from oracle import wc_oracle
    rng = np.random.default_rng(N)
    SC = rng.uniform(size=(N, N)) * (rng.uniform(size=(N, N)) < 0.4)
    SC = (SC + SC.T) / 2
    np.fill_diagonal(SC, 0.0)
    SC *= 2.5 / SC.sum(axis=1).mean()
    n1, n2, n3 = 40, 80, 160
    p = ops.make_params(N, n1, n2, n3, P=0.4, rhoE=0.18, seed=4)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    B = 130
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    streams = rng.integers(0, 2 ** 40, B).astype(np.uint64)
    for kernel in ("auto", "fma"):
        E, fin = ops.integrate_f32(p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), ds, streams=streams, kernel=kernel)
        assert E.shape == (8, N, B)
        for b in (0, 64, 129):
            Yo, fo = wc_oracle.run(SC, 0.16 + dG[b], 7.68 + ds[b], n1, n2, n3, seed=4, streams=[int(streams[b])], p=po, return_final=True)
            assert np.max(np.abs(E[:, :, b] - Yo[0, :, 0, :]) / np.abs(Yo[0, :, 0, :])) < 2e-4
            assert np.max(np.abs(fin[:, :, b] - fo[0]) / np.abs(fo[0])) < 2e-4


def test_run_many_seeds_schema_and_hma(aal90):
    """run_many_seeds.py:105-146 on the batched path (shortened horizon): the reference's pickle schema, FCs clipped in
    place by HMA, and HMA numbers equal to a fresh evaluation of the stored FC."""
    from nremmodfc_b200 import HMA, many_seeds, ops
    p = ops.make_params(90, 200, 2000, 30000, P=0.4, rhoE=0.18, seed=3)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    save = many_seeds.run_many_seeds(p, aal90["SC"], emp, aal90["map_ACh"], aal90["map_NA"], modality="map", seeds=range(3),
                                     Neq=500, bold_downsamp=20)
    assert sorted(save) == sorted((s, st) for s in range(3) for st in ("W", "N1", "N2", "N3"))
    for key, rec in save.items():
        assert set(rec) == {"Hin_sim", "Hse_sim", "Hin_node_sim", "Hse_node_sim", "sFC"}
        assert rec["sFC"].shape == (90, 90) and rec["sFC"].min() >= 0.0 and rec["Hin_node_sim"].shape == (90,)
        num, size, _ = HMA.Functional_HP(rec["sFC"].copy())
        hin, hse = HMA.Balance(rec["sFC"].copy(), num, size)
        assert abs(hin - rec["Hin_sim"]) < 1e-12 and abs(hse - rec["Hse_sim"]) < 1e-12
        assert 0 < rec["Hin_sim"] < 1 and rec["Hse_sim"] > 0
    # different states use different (delta_G, delta_sigma): W and N1 of one seed must differ
    assert not np.allclose(save[(0, "W")]["sFC"], save[(0, "N1")]["sFC"])


def _random_sc(N, seed):
    """netwWilsonCowanPlastic.py:64 placeholder distribution, zero diagonal, mean row sum scaled to AAL90's 2.5."""
    rng = np.random.default_rng(seed)
    SC = rng.uniform(size=(N, N))
    np.fill_diagonal(SC, 0.0)
    return SC * (2.5 / SC.sum(axis=1).mean())


@pytest.mark.parametrize("N,B,kernel,hetero", [(300, 130, "tc3", False), (300, 130, "tc", True), (1000, 256, "tc3", True),
                                               (528, 128, "tc3", False), (300, 130, "tcb", False), (1000, 256, "tcb", True),
                                               (528, 128, "auto", True), (300, 130, "bf3", False), (1000, 256, "bf3", True),
                                               (64, 128, "bf3", True)])
def test_large_connectome_integrator_vs_oracle(N, B, kernel, hetero, oracle_lib):
    """BASELINE configs[4] fast path (csrc/wc_big.cuh: one launch per Euler step, tcgen05 GEMM of the whole batch with the
    node update fused onto the TMEM accumulator) against the float64 oracle on the same Philox streams: recorded E rows,
    final (E, I, a_ie) and the coupling SC.E of the first step.  N = 300 / 528 exercise partial node slices and K padding."""
    from nremmodfc_b200 import ops
    from oracle import wc_oracle
    SC = _random_sc(N, N)
    rng = np.random.default_rng(N + 1)
    n1, n2, n3 = 20, 30, 60
    p = ops.make_params(N, n1, n2, n3, P=0.4, rhoE=0.18, seed=99)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    mG = rng.uniform(0.5, 1.5, N) if hetero else None
    mS = rng.uniform(0.8, 1.2, N) if hetero else None
    streams = rng.integers(0, 2 ** 62, B).astype(np.uint64)
    E, fin, coup = ops.big_integrate_f32(p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, streams, kernel=kernel,
                                         want_coupling=True)
    assert E.shape == (3, N, B) and fin.shape == (3, N, B) and coup.shape == (N, B)
    tol = 5e-3 if kernel == "tc" else 2e-4
    c0 = po["E0"] * SC.sum(axis=1) if isinstance(po, dict) else None
    if c0 is not None:
        # the tensor core truncates when it adds each K = 8 partial product to the FP32 accumulator: ~0.5 ulp bias per MMA,
        # (N/8) x passes MMAs per output -> ~4e-6 at N = 300, ~1.3e-5 at N = 1000 (3xTF32 itself is ~1e-6)
        assert np.max(np.abs(coup - c0[:, None]) / c0[:, None]) < (2e-3 if kernel == "tc" else 3e-5)
    for b in (0, 31, 127, B - 1):
        G = 0.16 + dG[b] * (mG if hetero else 1.0)
        sg = 7.68 + ds[b] * (mS if hetero else 1.0)
        Yo = oracle_lib.wc_run(SC, np.broadcast_to(G, N).copy(), np.broadcast_to(sg, N).copy(), n1, n2, n3, seed=99,
                               stream=int(streams[b]), p=po)
        fo = oracle_lib.wc_run(SC, np.broadcast_to(G, N).copy(), np.broadcast_to(sg, N).copy(), n1, n2, n3, seed=99,
                               stream=int(streams[b]), p=po, want="final")
        assert np.max(np.abs(E[:, :, b] - Yo[:, 0, :]) / np.abs(Yo[:, 0, :])) < tol
        assert np.max(np.abs(fin[:, :, b] - fo) / np.abs(fo)) < tol


@pytest.mark.parametrize("N,hetero", [(200, False), (300, True)])
def test_sweep_beyond_128_nodes_vs_oracle(N, hetero, oracle_lib, monkeypatch):
    """Fused sweep for a parcellation beyond 128 nodes (netwWilsonCowanPlastic.py:64-68 allows any nnodes; BASELINE configs[4]): the plan
    resolves to the large-connectome integrator (one launch per Euler step, wc_big.cuh) and feeds the same BOLD -> filter -> FC -> GoF ->
    Kuramoto chain.  (a) E samples of the plan's integrator == nrem_big_integrate_f32 on the same streams, checked against the float64
    oracle; (b) the stages against the oracle fed with the SAME samples (float64 BOLD state: 1e-6); (c) FC / GoF computed in batches of
    simulations (the N x N matrices of a large batch do not fit at once) == all at once; (d) a sliced run == one call."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle, wc_oracle
    n1, n2, n3 = 100, 300, 6000
    p = ops.make_params(N, n1, n2, n3, P=0.4, rhoE=0.18, seed=12)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    rng = np.random.default_rng(N)
    SC = _random_sc(N, N + 7)
    emp = np.stack([np.corrcoef(rng.normal(size=(N, 80)) + rng.normal(size=(1, 80))) for _ in range(2)])
    B = 140
    dG, ds = np.linspace(-0.1, 0.3, B), np.linspace(0.2, -0.2, B)
    streams = np.arange(B, dtype=np.uint64) * 5 + 3
    G0, s0 = np.full(B, 0.16), np.full(B, 7.68)
    mG = rng.uniform(0.5, 1.5, N) if hetero else None
    mS = rng.uniform(0.8, 1.2, N) if hetero else None
    kw = dict(kernel="auto", bold_f32=False, Neq=100, bold_downsamp=10, chunk_samples=37)
    plan = sweep.SweepPlan(p, B, n_maps=1, K=2, **kw)
    assert plan.kernel_name == "bf3"
    out = plan.run(SC, emp, G0, dG, s0, ds, streams, mG, mS, None, True)
    assert out["gof"].shape == (B, 2, 4) and out["fc"].shape == (B, N, N) and np.isfinite(out["gof"]).all()
    Eg, _ = ops.big_integrate_f32(p, SC, G0, dG, s0, ds, mG, mS, streams, kernel="bf3")
    for k in (0, 127, 139):
        G = 0.16 + dG[k] * (mG if hetero else np.ones(N))
        sg = 7.68 + ds[k] * (mS if hetero else np.ones(N))
        Eo = oracle_lib.wc_run(SC, G, sg, n1, n2, n3, seed=12, stream=int(streams[k]), p=po, want="E")
        assert (np.abs(Eg[0, :, k] - Eo[0]) / np.maximum(np.abs(Eo[0]), 0.05)).max() < 5e-4         # (a) first recorded row, see test_sweep_pipeline_vs_oracle
        E = Eg[:, :, k].astype(np.float64)
        bold = bold_oracle.filt_decimate(oracle_lib.bold_sim(E, 0.04), 10, 100, 0.04)
        assert np.max(np.abs(bold_oracle.fc(bold) - out["fc"][k])) < 1e-6                           # (b)
        g = np.array([bold_oracle.get_all_metrics(out["fc"][k], emp[j]) for j in range(2)])
        assert np.allclose(g, out["gof"][k], atol=1e-9)
        assert abs(out["mean"][k] - out["fc"][k].mean()) < 1e-12
        sync, meta = bold_oracle.kuramoto(bold)
        assert abs(out["sync"][k] - sync) < 1e-6 and abs(out["meta"][k] - meta) < 1e-6
    # (d) the same run in three slices, without the FC output: (c) NREM_SWEEP_FC_BATCH=33 makes finish() work in five batches
    monkeypatch.setenv("NREM_SWEEP_FC_BATCH", "33")
    plan2 = sweep.SweepPlan(p, B, n_maps=1, K=2, **kw)
    plan2.begin(SC, G0, dG, s0, ds, streams, mG, mS, None)
    left = plan2.chunks_total
    while left > 0:
        left = plan2.advance(3)
    out2 = plan2.finish(emp)
    for key in ("gof", "mean", "sync", "meta"):
        assert np.array_equal(out[key], out2[key]), key
    plan.close(); plan2.close()


@pytest.mark.parametrize("N,B", [(300, 130), (136, 256)])
def test_large_connectome_every_node_parameter_as_a_vector(N, B, oracle_lib):
    """"Any of them can be redefined as a vector of length nnodes" (netwWilsonCowanPlastic.py:21) beyond 128 nodes: all twelve per-node
    vectors through the large-connectome integrator (per-node table kernels of wc_big.cuh) against the NumPy oracle whose expressions
    broadcast; scalars given as a table reproduce the scalar run bit for bit; and through a sweep plan (node_params=...)."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle, wc_oracle
    rng = np.random.default_rng(N)
    vec = {"a_ee": 3.5 + 0.2 * rng.random(N), "a_ei": 3.75 - 0.2 * rng.random(N), "a_ii": 0.1 * rng.random(N),
           "tauE": 0.010 * (1 + 0.1 * rng.random(N)), "tauI": 0.020 * (1 + 0.1 * rng.random(N)), "P": 0.4 + 0.05 * rng.random(N),
           "rhoE": 0.18 + 0.02 * rng.random(N), "rE": 0.5 + 0.05 * rng.random(N), "rI": 0.5 - 0.05 * rng.random(N),
           "mu": 1.0 + 0.05 * rng.random(N), "sigmaI": 4.0 + 0.3 * rng.random(N), "a_ie_0": 2.5 + 0.2 * rng.random(N)}
    SC = _random_sc(N, N + 3)
    n1, n2, n3 = 20, 30, 60
    p = ops.make_params(N, n1, n2, n3, seed=8)
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    mG, mS = rng.uniform(0.5, 1.5, N), rng.uniform(0.8, 1.2, N)
    streams = np.arange(B, dtype=np.uint64) + 3
    args = (p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), ds)
    E, fin = ops.big_integrate_f32(*args, mG, mS, streams, node_params=vec)
    po = wc_oracle.params(**vec)
    for b in (0, 77, B - 1):
        Yo, fo = wc_oracle.run(SC, 0.16 + dG[b] * mG, 7.68 + ds[b] * mS, n1, n2, n3, seed=8, streams=[int(streams[b])], p=po, return_final=True)
        assert np.max(np.abs(E[:, :, b] - Yo[0, :, 0, :]) / np.abs(Yo[0, :, 0, :])) < 2e-4
        assert np.max(np.abs(fin[:, :, b] - fo[0]) / np.abs(fo[0])) < 2e-4
    # homogeneous call (no maps) with a table; scalars as a table == the scalar kernel bit for bit; a real table changes the result
    E0, f0 = ops.big_integrate_f32(*args, None, None, streams)
    E1, f1 = ops.big_integrate_f32(*args, None, None, streams, node_params={"P": np.full(N, p.P), "tauE": np.full(N, p.tauE)})
    assert np.array_equal(E0, E1) and np.array_equal(f0, f1) and not np.allclose(E0, E)
    with pytest.raises(Exception):
        ops.big_integrate_f32(*args, None, None, streams, kernel="tc3", node_params=vec)
    # sweep plan beyond 128 nodes with a table: same samples -> same FC as the oracle chain
    p2 = ops.make_params(N, 100, 300, 4000, seed=8)
    emp = np.stack([np.corrcoef(rng.normal(size=(N, 80)) + rng.normal(size=(1, 80))) for _ in range(2)])
    out = sweep.sweep_gof(p2, SC, emp, 0.16, dG, 7.68, ds, streams, mapG=mG[None], mapS=mS[None], want_fc=True, node_params=vec,
                          bold_f32=False, Neq=50, bold_downsamp=10, chunk_samples=64)
    Eg, _ = ops.big_integrate_f32(p2, SC, np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, streams, node_params=vec)
    for b in (0, B - 1):
        bold = bold_oracle.filt_decimate(oracle_lib.bold_sim(Eg[:, :, b].astype(np.float64), 0.04), 10, 50, 0.04)
        assert np.max(np.abs(bold_oracle.fc(bold) - out["fc"][b])) < 1e-6


def test_large_connectome_persistent_cluster_mode_is_bit_identical(monkeypatch):
    """NREM_BIG_PERSIST=1 runs the same step code inside one thread-block cluster per tile (barrier.cluster per step instead of a
    launch per step): the arithmetic is identical, so E samples and final state must match the per-step launches bit for bit."""
    from nremmodfc_b200 import ops
    N, B = 520, 200
    SC = _random_sc(N, 3)
    rng = np.random.default_rng(8)
    p = ops.make_params(N, 30, 4100, 80, P=0.4, rhoE=0.18, seed=5)                  # crosses an a_ie recombination (step 4096)
    args = (p, SC, np.full(B, 0.16), rng.uniform(-0.1, 0.3, B), np.full(B, 7.68), rng.uniform(-0.2, 0.2, B))
    kw = dict(mapG=rng.uniform(0.5, 1.5, N), mapS=rng.uniform(0.8, 1.2, N), streams=rng.integers(0, 2 ** 62, B).astype(np.uint64))
    monkeypatch.setenv("NREM_BIG_PERSIST", "0")
    E0, f0 = ops.big_integrate_f32(*args, **kw)
    monkeypatch.setenv("NREM_BIG_PERSIST", "1")
    E1, f1 = ops.big_integrate_f32(*args, **kw)
    assert E0.shape == (4, N, B) and np.isfinite(f0).all()
    assert np.array_equal(E0, E1) and np.array_equal(f0, f1)


@pytest.mark.parametrize("kernel,B", [("tcb", 1100), ("tc3", 256), ("tc", 384), ("bf3", 1100), ("bf3", 384)])
def test_large_connectome_cta_pair_mode_is_bit_identical(kernel, B, monkeypatch):
    """CTA pairs (tcgen05 cta_group::2: two 128-simulation tiles per M = 256 MMA, each CTA staging half of the SC tile) against
    single-CTA MMAs: same products, same accumulation order, so E samples, final state and the first coupling must be identical
    bit for bit.  B = 1100 is 9 tiles (padded to 10 for the pairs), B = 384 is 3 tiles (pairs forced, padded to 4)."""
    from nremmodfc_b200 import ops
    N = 520
    SC = _random_sc(N, 4)
    rng = np.random.default_rng(9)
    p = ops.make_params(N, 30, 100, 80, P=0.4, rhoE=0.18, seed=6)
    args = (p, SC, np.full(B, 0.16), rng.uniform(-0.1, 0.3, B), np.full(B, 7.68), rng.uniform(-0.2, 0.2, B))
    kw = dict(mapG=rng.uniform(0.5, 1.5, N), mapS=rng.uniform(0.8, 1.2, N), streams=rng.integers(0, 2 ** 62, B).astype(np.uint64),
              kernel=kernel, want_coupling=True)
    monkeypatch.setenv("NREM_BIG_PAIR", "0")
    E0, f0, c0 = ops.big_integrate_f32(*args, **kw)
    monkeypatch.setenv("NREM_BIG_PAIR", "1")
    E1, f1, c1 = ops.big_integrate_f32(*args, **kw)
    assert E0.shape == (4, N, B) and np.isfinite(f0).all() and np.isfinite(f1).all()
    assert np.array_equal(c0, c1), float(np.max(np.abs(c0 - c1)))
    assert np.array_equal(E0, E1) and np.array_equal(f0, f1)


@pytest.mark.parametrize("big_kernel", ["tc3", "bf3"])
def test_large_connectome_matches_small_path_statistics(big_kernel, aal90):
    """The per-step kernel and the register-resident kernel integrate the same model with the same noise: on AAL90 (N = 90,
    which both accept) their float32 trajectories agree to rounding over a short horizon, incl. an a_ie recombination (for bf3: the
    re-split of a_ie into a bf16 base and its FP32 remainder)."""
    from nremmodfc_b200 import ops
    p = ops.make_params(90, 100, 4200, 400, P=0.4, rhoE=0.18, seed=5)
    B = 128
    dG = np.linspace(-0.1, 0.02, B)
    st = np.arange(B, dtype=np.uint64) + 11
    E1, f1 = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), np.zeros(B), streams=st, kernel="tc3")
    E2, f2 = ops.big_integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), np.zeros(B), streams=st, kernel=big_kernel)
    assert np.max(np.abs(E1[0] - E2[0]) / np.abs(E1[0])) < 5e-3          # 0.43 s of chaotic float32 dynamics
    assert np.max(np.abs(f1[2] - f2[2]) / np.abs(f1[2])) < 5e-3


def test_large_connectome_argument_errors_and_ragged_batch(oracle_lib):
    """Error behaviour at the C ABI (negative status -> NremError with the library's message) and a batch that is not a
    multiple of the 128-simulation tile (padding simulations never leak into the outputs)."""
    from nremmodfc_b200 import ops
    from nremmodfc_b200._lib import NremError
    from oracle import wc_oracle
    N = 40
    SC = _random_sc(N, 1)
    p = ops.make_params(N, 5, 5, 20, P=0.4, rhoE=0.18, seed=3)
    with pytest.raises(NremError):
        ops.big_integrate_f32(ops.make_params(8, 5, 5, 20), _random_sc(8, 2), [0.16], [0.0], [7.68], [0.0])      # nnodes < 16
    with pytest.raises(NremError):
        ops.big_integrate_f32(p, SC, [0.16], [0.0], [7.68], [0.0], kernel="fma")                                   # no FMA variant
    with pytest.raises(ValueError):
        ops.big_integrate_f32(p, SC[:, :-1], [0.16], [0.0], [7.68], [0.0])
    E, fin = ops.big_integrate_f32(p, SC, np.full(3, 0.16), [0.0, 0.1, 0.2], np.full(3, 7.68), [0.0, 0.0, -0.1],
                                   streams=np.array([7, 8, 9], dtype=np.uint64), kernel="tc3")
    assert E.shape == (1, N, 3) and fin.shape == (3, N, 3)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    for b, (g, s) in enumerate([(0.16, 7.68), (0.26, 7.68), (0.36, 7.58)]):
        fo = oracle_lib.wc_run(SC, np.full(N, g), np.full(N, s), 5, 5, 20, seed=3, stream=7 + b, p=po, want="final")
        assert np.max(np.abs(fin[:, :, b] - fo) / np.abs(fo)) < 2e-5


def test_every_node_parameter_as_a_vector(aal90):
    """"Any of them can be redefined as a vector of length nnodes" (netwWilsonCowanPlastic.py:21): the float64 path takes per-node
    vectors for all eleven node parameters; checked against the NumPy oracle (whose expressions broadcast) and through the drop-in
    module surface."""
    from nremmodfc_b200 import ops
    from nremmodfc_b200 import netwWilsonCowanPlastic as wc
    from oracle import wc_oracle
    rng = np.random.default_rng(21)
    N = 90
    vec = {"a_ee": 3.5 + 0.2 * rng.random(N), "a_ei": 3.75 - 0.2 * rng.random(N), "a_ii": 0.1 * rng.random(N),
           "tauE": 0.010 * (1 + 0.1 * rng.random(N)), "tauI": 0.020 * (1 + 0.1 * rng.random(N)), "P": 0.4 + 0.05 * rng.random(N),
           "rhoE": 0.18 + 0.02 * rng.random(N), "rE": 0.5 + 0.05 * rng.random(N), "rI": 0.5 - 0.05 * rng.random(N),
           "mu": 1.0 + 0.05 * rng.random(N), "sigmaI": 4.0 + 0.3 * rng.random(N)}
    n1, n2, n3 = 50, 100, 200
    p = ops.make_params(N, n1, n2, n3, seed=8)
    G, sg = 0.16 + 0.05 * rng.random(N), 7.68 + 0.2 * rng.random(N)
    Y, fin = ops.wc_run(p, aal90["SC"], G, sg, B=1, streams=[5], node_params=vec)
    po = wc_oracle.params(**vec)
    Yo, fo = wc_oracle.run(aal90["SC"], G, sg, n1, n2, n3, seed=8, streams=[5], p=po, return_final=True)
    assert np.max(np.abs(Y[0] - Yo[0]) / np.abs(Yo[0])) < 1e-9 and np.max(np.abs(fin[0] - fo[0]) / np.abs(fo[0])) < 1e-9
    # a subset of names keeps the scalars of p for the others; it must differ from the all-scalar run
    Y1, _ = ops.wc_run(p, aal90["SC"], G, sg, B=1, streams=[5], node_params={"P": vec["P"]})
    Y0, _ = ops.wc_run(p, aal90["SC"], G, sg, B=1, streams=[5])
    assert not np.allclose(Y1, Y0) and not np.allclose(Y1, Y)
    with pytest.raises(ValueError):
        ops.wc_run(p, aal90["SC"], G, sg, node_params={"G0": vec["P"]})
    with pytest.raises(ValueError):
        ops.wc_run(p, aal90["SC"], G, sg, node_params={"P": vec["P"][:7]})
    # drop-in surface: module attributes redefined as vectors
    saved = {k: getattr(wc, k) for k in list(vec) + ["CM", "G", "sigmaE", "sid", "replicate", "timeTrans1", "timeTrans2", "timeSim", "time"]}
    try:
        for k, v in vec.items():
            setattr(wc, k, v)
        wc.CM, wc.G, wc.sigmaE, wc.sid, wc.replicate = aal90["SC"], G, sg, 8, 5
        wc.timeTrans1, wc.timeTrans2, wc.timeSim = np.arange(n1), np.arange(n2), np.arange(n3)
        wc.time = np.arange((n3 + 19) // 20)
        Yd = wc.run()
    finally:
        for k, v in saved.items():
            setattr(wc, k, v)
    assert np.array_equal(Yd, Y[0])


# ---- node-lane integrator (csrc/wc_node.cuh): small tiles, N <= 128, per-node parameter tables -------------------------------

@pytest.mark.parametrize("hetero", [False, True])
def test_node_lane_kernel_is_bit_identical_to_the_128_simulation_kernel(hetero, aal90):
    """The low-latency kernel (a thread = one node x 8 or 4 simulations, transposed tcgen05 contraction, 32 / 16 simulations per
    CTA) performs the same float32 operations on the same Philox stream as the throughput kernel (a thread = one simulation x 24
    nodes): E samples and final state must agree bit for bit, across an a_ie recombination (global step 4096), for a ragged
    batch, homogeneous and with NA/ACh maps."""
    from nremmodfc_b200 import ops
    p = ops.make_params(90, 300, 4000, 1200, P=0.4, rhoE=0.18, seed=12)
    B = 200                                                    # run_many_seeds.py: 50 seeds x 4 states
    rng = np.random.default_rng(4)
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    streams = rng.integers(0, 2 ** 62, B).astype(np.uint64)
    kw = dict(streams=streams)
    if hetero:
        kw.update(mapG=np.stack([np.ones(90), aal90["map_ACh"] / aal90["map_ACh"].mean()]),
                  mapS=np.stack([np.ones(90), aal90["map_NA"] / aal90["map_NA"].mean()]),
                  map_id=np.r_[np.zeros(128, np.int32), np.ones(B - 128, np.int32)])
    ref = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, kernel="tc3", **kw)
    assert np.isfinite(ref[0]).all() and ref[0].shape == (60, 90, B)
    for kernel in ("node32", "node16"):
        got = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, kernel=kernel, **kw)
        dE = np.max(np.abs(got[0] - ref[0]))
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (kernel, dE)


@pytest.mark.parametrize("N,kernel", [(116, "node32"), (128, "node16"), (100, "auto"), (68, "node32")])
def test_node_lane_kernel_other_parcellations_vs_oracle(N, kernel, oracle_lib):
    """Connectomes up to 128 nodes (AAL116, Schaefer-100 ...; netwWilsonCowanPlastic.py:64-68 takes any len(CM)) on the node-lane
    kernel against the float64 oracle on the same Philox streams, and the fused sweep's FC / GoF tail at that size."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle, wc_oracle
    rng = np.random.default_rng(N)
    SC = rng.uniform(size=(N, N)) * (rng.uniform(size=(N, N)) < 0.4)
    SC = (SC + SC.T) / 2
    np.fill_diagonal(SC, 0.0)
    SC *= 2.5 / SC.sum(axis=1).mean()
    n1, n2, n3 = 40, 80, 160
    p = ops.make_params(N, n1, n2, n3, P=0.4, rhoE=0.18, seed=4)
    po = wc_oracle.params(P=0.4, rhoE=0.18)
    B = 70
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    mG, mS = rng.uniform(0.5, 1.5, (1, N)), rng.uniform(0.8, 1.2, (1, N))
    streams = rng.integers(0, 2 ** 40, B).astype(np.uint64)
    E, fin = ops.integrate_f32(p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, streams=streams, kernel=kernel)
    assert E.shape == (8, N, B)
    for b in (0, 33, 69):
        Yo, fo = wc_oracle.run(SC, 0.16 + dG[b] * mG[0], 7.68 + ds[b] * mS[0], n1, n2, n3, seed=4, streams=[int(streams[b])], p=po, return_final=True)
        assert np.max(np.abs(E[:, :, b] - Yo[0, :, 0, :]) / np.abs(Yo[0, :, 0, :])) < 2e-4
        assert np.max(np.abs(fin[:, :, b] - fo[0]) / np.abs(fo[0])) < 2e-4
    # fused sweep at this size: stages after the integrator vs the oracle fed with the same samples
    p2 = ops.make_params(N, 200, 800, 8000, P=0.4, rhoE=0.18, seed=5)
    emp = np.stack([np.corrcoef(rng.normal(size=(N, 60)) + rng.normal(size=(1, 60))) for _ in range(4)])
    plan = sweep.SweepPlan(p2, B, kernel=kernel, bold_f32=False, Neq=100, bold_downsamp=10, chunk_samples=64)
    assert plan.kernel_name in ("node32", "node16")
    out = plan.run(SC, emp, 0.16, dG, 7.68, ds, streams, mapG=mG, mapS=mS, want_fc=True)
    plan.close()
    Eg, _ = ops.integrate_f32(p2, SC, np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG, mS, streams=streams, kernel=plan_kernel(kernel, N))
    for b in (0, 69):
        bold = bold_oracle.filt_decimate(oracle_lib.bold_sim(Eg[:, :, b].astype(np.float64), 0.04), 10, 100, 0.04)
        assert np.max(np.abs(bold_oracle.fc(bold) - out["fc"][b])) < 1e-6
        g = np.array([bold_oracle.get_all_metrics(out["fc"][b], emp[j]) for j in range(4)])
        assert np.allclose(g, out["gof"][b], atol=1e-9)


def plan_kernel(kernel, N):
    return "node16" if kernel == "auto" else kernel            # auto with 70 simulations: 16-simulation tiles


def test_sweep_takes_every_node_parameter_as_a_vector(aal90, oracle_lib):
    """"Any of them can be redefined as a vector of length nnodes" (netwWilsonCowanPlastic.py:21) on the batched path: all twelve
    per-node vectors (the eleven model parameters and a_ie_0) through `node_params`, against the NumPy oracle whose expressions
    broadcast; then through the fused sweep (same samples -> same FC as the oracle chain)."""
    from nremmodfc_b200 import ops, sweep
    from oracle import bold_oracle, wc_oracle
    rng = np.random.default_rng(21)
    N = 90
    vec = {"a_ee": 3.5 + 0.2 * rng.random(N), "a_ei": 3.75 - 0.2 * rng.random(N), "a_ii": 0.1 * rng.random(N),
           "tauE": 0.010 * (1 + 0.1 * rng.random(N)), "tauI": 0.020 * (1 + 0.1 * rng.random(N)), "P": 0.4 + 0.05 * rng.random(N),
           "rhoE": 0.18 + 0.02 * rng.random(N), "rE": 0.5 + 0.05 * rng.random(N), "rI": 0.5 - 0.05 * rng.random(N),
           "mu": 1.0 + 0.05 * rng.random(N), "sigmaI": 4.0 + 0.3 * rng.random(N), "a_ie_0": 2.5 + 0.2 * rng.random(N)}
    n1, n2, n3 = 50, 100, 200
    p = ops.make_params(N, n1, n2, n3, seed=8)
    B = 40
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    mG, mS = aal90["map_ACh"] / aal90["map_ACh"].mean(), aal90["map_NA"] / aal90["map_NA"].mean()
    streams = np.arange(B, dtype=np.uint64) + 3
    E, fin = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG[None], mS[None], streams=streams,
                               kernel="auto", node_params=vec)
    po = wc_oracle.params(**vec)
    for b in (0, 17, 39):
        Yo, fo = wc_oracle.run(aal90["SC"], 0.16 + dG[b] * mG, 7.68 + ds[b] * mS, n1, n2, n3, seed=8, streams=[int(streams[b])], p=po,
                               return_final=True)
        assert np.max(np.abs(E[:, :, b] - Yo[0, :, 0, :]) / np.abs(Yo[0, :, 0, :])) < 2e-4
        assert np.max(np.abs(fin[:, :, b] - fo[0]) / np.abs(fo[0])) < 2e-4
    # scalars given as a "vector table" reproduce the scalar run bit for bit; a table changes the result
    E0, _ = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG[None], mS[None], streams=streams, kernel="node16")
    E1, _ = ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG[None], mS[None], streams=streams, kernel="node16",
                              node_params={"P": np.full(N, p.P)})
    assert np.array_equal(E0, E1) and not np.allclose(E0, E)
    with pytest.raises(Exception):
        ops.integrate_f32(p, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, streams=streams, kernel="tc3", node_params=vec)
    # fused sweep
    p2 = ops.make_params(N, 200, 800, 8000, seed=8)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    out = sweep.sweep_gof(p2, aal90["SC"], emp, 0.16, dG, 7.68, ds, streams, mapG=mG[None], mapS=mS[None], want_fc=True, node_params=vec,
                          bold_f32=False, Neq=100, bold_downsamp=10, chunk_samples=64)
    Eg, _ = ops.integrate_f32(p2, aal90["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, mG[None], mS[None], streams=streams,
                              kernel="auto", node_params=vec)
    for b in (0, 39):
        bold = bold_oracle.filt_decimate(oracle_lib.bold_sim(Eg[:, :, b].astype(np.float64), 0.04), 10, 100, 0.04)
        assert np.max(np.abs(bold_oracle.fc(bold) - out["fc"][b])) < 1e-6


def test_small_batch_uses_small_tiles_and_matches_the_large_batch(aal90):
    """kernel="auto": a batch that cannot fill the SMs with 128-simulation tiles runs on 16- / 32-simulation tiles; the result
    table is identical to the same simulations inside a big batch on the throughput kernel."""
    import torch
    from nremmodfc_b200 import ops, sweep
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    p = ops.make_params(90, 100, 300, 2400, P=0.4, rhoE=0.18, seed=11)
    kw = dict(Neq=40, bold_downsamp=5, chunk_samples=50, bold_f32=True)
    rng = np.random.default_rng(6)
    Bbig = 66 * sms * 2
    dG, ds = rng.uniform(-0.1, 0.3, Bbig), rng.uniform(-0.2, 0.2, Bbig)
    streams = rng.integers(0, 2 ** 60, Bbig).astype(np.uint64)
    big = sweep.SweepPlan(p, Bbig, **kw)
    assert big.kernel_name == "tc3"
    ref = big.run(aal90["SC"], emp, 0.16, dG, 7.68, ds, streams)
    big.close()
    for B, name in ((200, "node16"), (20 * sms, "node32")):
        plan = sweep.SweepPlan(p, B, **kw)
        assert plan.kernel_name == name
        out = plan.run(aal90["SC"], emp, 0.16, dG[:B], 7.68, ds[:B], streams[:B])
        plan.close()
        assert np.array_equal(out["gof"], ref["gof"][:B]) and np.array_equal(out["sync"], ref["sync"][:B])
