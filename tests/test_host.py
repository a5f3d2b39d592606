"""CPU-only tests: the C-ABI library loads and exports what include/nremfc.h declares, host-side
sweep logic, the drop-in module surface, and the 2-rank gloo gather.  No compute calls (no GPU)."""
import itertools
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built():
    from nremmodfc_b200 import build
    return build.build_library()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "nremfc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nrem_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 16
    from nremmodfc_b200 import _lib
    assert declared == set(_lib.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.nrem_abi_version() == 2
    nm = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", nm), f"{name} is not an exported text symbol"


def test_library_is_built_for_sm100a(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nremmodfc_b200 import _lib, ops
    assert _lib.lib.nrem_device_count() == 0
    with pytest.raises(_lib.NremError):
        ops.fc(np.zeros((10, 4)))
    import nremmodfc_b200.netwWilsonCowanPlastic as wc
    with pytest.raises(_lib.NremError):
        wc.run()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under nremmodfc_b200/ or compat/ may import, include or load it
    (comments may cite oracle/philox.py as the definition of the noise stream)."""
    bad = []
    for top in ("nremmodfc_b200", "compat"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle|#include.*oracle|liboracle", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad


def test_module_surface_defaults(built):
    """Same attribute names and defaults as netwWilsonCowanPlastic.py:23-68."""
    import importlib
    import nremmodfc_b200.netwWilsonCowanPlastic as wc
    wc = importlib.reload(wc)
    expect = dict(a_ee=3.5, a_ie_0=2.5, a_ei=3.75, a_ii=0, tauE=0.010, tauI=0.020, P=0.4, Q=0, rhoE=0.14, tau_ip=2,
                  rE=0.5, rI=0.5, mu=1, sigmaE=4, sigmaI=4, tTrans1=600, tTrans2=600, tstop=600, dt=0.002,
                  dtSim=0.0001, downsamp=20, D=0.002, sid=12, G=0.7, nnodes=90, N=90)
    for k, v in expect.items():
        assert getattr(wc, k) == v, k
    assert abs(wc.sqdtD - 0.2) < 1e-15
    assert wc.CM.shape == (90, 90)
    assert len(wc.timeSim) == 6000000 and len(wc.time) == 300000
    assert wc.run.recompile() is None and wc.wilsonCowan.recompile() is None
    assert callable(wc.run) and callable(wc.simBOLD) and callable(wc.S)
    assert abs(wc.S(1.0, 4.0, 1.0) - 0.5) < 1e-15


def test_make_params_validation(built):
    from nremmodfc_b200 import ops
    p = ops.make_params(90, 1, 2, 3, P=0.4, rhoE=0.18, seed=2 ** 63 + 5)
    assert (p.nnodes, p.n1, p.n2, p.n3, p.downsamp) == (90, 1, 2, 3, 20)
    assert p.seed == 2 ** 63 + 5 and abs(p.sqdtD - 0.2) < 1e-15
    with pytest.raises(ValueError):
        ops.make_params(90, 1, 2, 3, nonsense=1)
    with pytest.raises(ValueError):
        ops.make_params(90, 1, 2, 3, tauE=np.ones(90))


def test_product_grid_order(built):
    from nremmodfc_b200 import sweep
    seeds = np.arange(3)
    dG = np.linspace(-0.1, 0.3, 4, endpoint=False)
    dS = np.linspace(-0.2, 0.2, 5, endpoint=False)
    s, g, sg = sweep.product_grid(seeds, dG, dS)
    ref = list(itertools.product(seeds, dG, dS))                  # whole_sweep_both.py:61
    assert len(s) == 60
    for k, (a, b, c) in enumerate(ref):
        assert (s[k], g[k], sg[k]) == (a, b, c)


def test_shard_ids_match_reference_round_robin(built):
    from nremmodfc_b200 import sweep
    n, world = 1003, 8
    seen = []
    for r in range(world):
        ids = sweep.shard_ids(n, r, world)
        assert all(i % world == r for i in ids)                   # whole_sweep_both.py:64
        seen.append(ids)
    assert sorted(np.concatenate(seen)) == list(range(n))
    cont = [sweep.shard_ids(n, r, world, contiguous=True) for r in range(world)]
    assert np.array_equal(np.concatenate(cont), np.arange(n))


def test_pad_by_map(built):
    from nremmodfc_b200 import sweep
    map_id = np.array([0] * 130 + [2] * 5 + [1] * 128)
    src, valid = sweep.pad_by_map(map_id)
    assert len(src) % 128 == 0 and len(src) == 256 + 128 + 128
    tiles = map_id[src].reshape(-1, 128)
    assert all(len(set(t)) == 1 for t in tiles)
    assert sorted(src[valid]) == list(range(len(map_id)))


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from nremmodfc_b200 import sweep
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n = 37
ids = sweep.shard_ids(n, rank, world)
rows = np.stack([ids * 10.0 + c for c in range(5)], axis=1)        # row content is a function of the global id
table = sweep.gather_rows(ids, rows, n)
expect = np.stack([np.arange(n) * 10.0 + c for c in range(5)], axis=1)
assert np.array_equal(table, expect), table
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gather_gloo(built, tmp_path):
    """World-size-2 run of the sharding + final gather (the only exchange on this path)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for r, pr in enumerate(procs):
        out, _ = pr.communicate(timeout=180)
        assert pr.returncode == 0, out
        assert f"ok {r}" in out


def test_compat_modules_alias_the_package(built):
    """`import netwWilsonCowanPlastic as wc` (whole_sweep_both.py:10) must return the package's module object."""
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import netwWilsonCowanPlastic as wc, BOLDModel as BD, utils\n"
            "from skimage.metrics import structural_similarity as ssim\n"
            "import nremmodfc_b200.netwWilsonCowanPlastic as m\n"
            "assert wc is m and callable(BD.Sim) and callable(utils.get_all_metrics) and callable(utils.kuramoto) and callable(ssim)\n"
            "wc.P = 0.123\nassert m.P == 0.123\nprint('ok')\n") % (os.path.join(ROOT, "compat"), ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_c_abi_argument_errors(built):
    """Bad arguments come back as negative status codes with a message (no exceptions, no CUDA needed)."""
    import ctypes as C
    from nremmodfc_b200 import _lib, ops
    lib = _lib.lib
    p = ops.make_params(90, 1, 1, 20)
    one = C.c_void_p(8)                                        # a non-null dummy "device pointer": never dereferenced
    assert lib.nrem_fc_f64(None, 1, 10, 90, one, None) == -1 and b"null" in lib.nrem_last_error()
    assert lib.nrem_fc_f64(one, 1, 10, 40000, one, None) == -1 and b"N <= 32768" in lib.nrem_last_error()
    assert lib.nrem_gof_f64(one, one, 1, 4, 3, 1.0, one, None, None) == -1
    assert lib.nrem_gof_f64(one, one, 1, 4, 40000, 1.0, one, None, None) == -1 and b"N <= 32768" in lib.nrem_last_error()
    assert lib.nrem_bold_sim_f64(one, 0, 10, 90, 0.04, one, None) == -1
    assert lib.nrem_wc_run_f64(None, one, one, one, None, None, 1, 1, 1, None, one, None) == -1
    assert lib.nrem_wc_run_f64(C.byref(p), one, one, one, None, one, 3, 5, 1, None, one, None) == -1      # noise_batch not in {1, B}
    hb = (C.c_double * 5)(1, 0, -2, 0, 1)
    ha = (C.c_double * 5)(2, 0, 0, 0, 0)
    assert lib.nrem_filtfilt_decimate_f64(one, 1, 40, 3, 0, 1, hb, ha, one, one, None) == -1 and b"a[0]" in lib.nrem_last_error()
    ha = (C.c_double * 5)(1, -0.5, 0, 0, 0)                    # real poles: not two complex-conjugate pairs
    assert lib.nrem_filtfilt_decimate_f64(one, 1, 40, 3, 0, 1, hb, ha, one, one, None) == -3
    assert lib.nrem_filt_scratch_bytes(1, 40, 3, 20, 1) == -1      # fewer than 32 samples after the cut
    o = _lib.SweepOpts()
    plan = C.c_void_p()
    pp = ops.make_params(9000, 1, 1, 20)
    assert lib.nrem_sweep_create(C.byref(pp), C.byref(o), 4, 1, 4, C.byref(plan)) == -1 and b"nnodes <= 8192" in lib.nrem_last_error()
    pp = ops.make_params(200, 1, 1, 20)
    o.kernel, o.bold_downsamp = 5, 10                          # beyond 128 nodes only the large-connectome integrator (0 / 7) ...
    assert lib.nrem_sweep_create(C.byref(pp), C.byref(o), 4, 1, 4, C.byref(plan)) == -1 and b"large-connectome" in lib.nrem_last_error()
    o.kernel = 0                                               # ... which takes one map pair per sweep
    assert lib.nrem_sweep_create(C.byref(pp), C.byref(o), 4, 2, 4, C.byref(plan)) == -1 and b"one (mapG, mapS) pair" in lib.nrem_last_error()
    pw = ops.make_params(120, 1, 1, 20)
    o.kernel, o.bold_downsamp = 3, 10                          # the 128-simulation tcgen05 kernel is one 96-wide MMA tile
    assert lib.nrem_sweep_create(C.byref(pw), C.byref(o), 4, 1, 4, C.byref(plan)) == -1 and b"node-lane" in lib.nrem_last_error()
    o.kernel = 4
    assert lib.nrem_sweep_create(C.byref(pp), C.byref(o), 4, 1, 4, C.byref(plan)) == -1
    assert lib.nrem_sweep_begin(None, *([one] * 7), None, one, 1, None) == -1
    assert lib.nrem_sweep_advance(None, 1, None, None) == -1 and lib.nrem_sweep_finish(None, one, one, None, None, None) == -1
    assert lib.nrem_sweep_feed_samples(None, one, 1, None) == -1 and lib.nrem_sweep_set_node_params(None, None, None) == -1
    assert lib.nrem_sweep_chunks_total(None) == -1 and lib.nrem_sweep_kernel(None) == -1
    assert lib.nrem_sweep_run(None, *([one] * 7), None, one, one, one, one, None, None) == -1
    assert lib.nrem_selftest_tc_coupling(one, one, one, 2, 0, 0, 0, 0, 0, None) == -1


def test_shuffled_symmetric_maps_follow_the_reference_rule(aal90):
    """empirical/retrieve_AALmaps.py:62-75: left hemisphere permuted, right hemisphere mirrored.  The reference's own shuffled
    VAChT map (committed fixture) satisfies exactly this relation to the unshuffled one."""
    from nremmodfc_b200 import sweep
    o, s = aal90["map_ACh"], aal90["map_ACh_shuf"]
    perm = [int(np.argmin(np.abs(o[:45] - v))) for v in s[:45]]
    assert sorted(perm) == list(range(45))
    assert all(abs(s[89 - i] - o[89 - perm[i]]) < 1e-12 for i in range(45))
    maps, idx = sweep.shuffled_symmetric_maps(o, 5, seed=3)
    assert maps.shape == (5, 90) and idx.shape == (5, 90)
    for k in range(5):
        assert sorted(idx[k]) == list(range(90)) and sorted(idx[k, :45]) == list(range(45))
        assert all(idx[k, 89 - i] == 89 - idx[k, i] for i in range(45))
        assert np.array_equal(maps[k], o[idx[k]]) and abs(maps[k].mean() - o.mean()) < 1e-12
    m2, _ = sweep.shuffled_symmetric_maps(o, 5, seed=3)
    assert np.array_equal(maps, m2) and not np.array_equal(maps[0], maps[1])
    with pytest.raises(ValueError):
        sweep.shuffled_symmetric_maps(np.ones(7), 1, 0)


def test_many_seeds_batch_assembly(monkeypatch, aal90):
    """run_many_seeds.py:105-136 as one batch: (seed, state) order of itertools.product, per-state optima of :34-47, maps normalised
    to mean 1, unique stream ids, and the reference's record layout — checked on the host with the GPU call replaced."""
    from nremmodfc_b200 import many_seeds, sweep
    seen = {}

    def fake_sweep_gof(p, CM, emp, G0, dG, s0, ds, streams, mapG=None, mapS=None, want_fc=False, **kw):
        seen.update(G0=G0, dG=dG, s0=s0, ds=ds, streams=streams, mapG=mapG, mapS=mapS, want_fc=want_fc, kw=kw)
        rng = np.random.default_rng(0)
        fc = rng.uniform(-0.2, 0.9, (len(G0), 90, 90))
        fc = (fc + fc.transpose(0, 2, 1)) / 2
        for f in fc:
            np.fill_diagonal(f, 1.0)
        return {"fc": fc}

    monkeypatch.setattr(sweep, "sweep_gof", fake_sweep_gof)
    emp = np.stack([aal90[s] for s in ("W", "N1", "N2", "N3")])
    save = many_seeds.run_many_seeds(None, aal90["SC"], emp, aal90["map_ACh"], aal90["map_NA"], modality="map", seeds=range(3), Neq=7)
    assert list(save) == [(s, st) for s in range(3) for st in ("W", "N1", "N2", "N3")]
    assert seen["want_fc"] and seen["kw"] == {"Neq": 7}
    opt = many_seeds.OPTIMALS["map"]
    assert np.allclose(seen["G0"], 0.16) and np.allclose(seen["s0"], 7.68)
    assert np.allclose(seen["dG"], [opt[st][1] for _ in range(3) for st in ("W", "N1", "N2", "N3")])
    assert np.allclose(seen["ds"], [opt[st][3] for _ in range(3) for st in ("W", "N1", "N2", "N3")])
    assert abs(seen["mapG"].mean() - 1) < 1e-12 and abs(seen["mapS"].mean() - 1) < 1e-12 and seen["mapG"].shape == (1, 90)
    assert len(set(int(x) for x in seen["streams"])) == 12
    rec = save[(1, "N2")]
    assert set(rec) == {"Hin_sim", "Hse_sim", "Hin_node_sim", "Hse_node_sim", "sFC"}
    assert rec["sFC"].min() >= 0.0 and rec["Hin_node_sim"].shape == (90,)           # HMA clips negatives in place, as the reference does
    assert many_seeds.OPTIMALS["homo"]["N3"] == (0.16, -0.04, 7.68, 0.04) and many_seeds.OPTIMALS["shuf"]["N1"] == (0.16, 0.0, 7.68, 0.04)


def _staged():
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "whole_sweep_both.py"))


@pytest.mark.skipif(not _staged(), reason="baseline/_ref is not staged (build() copies it where /root/reference exists)")
def test_unmodified_drivers_reach_the_cuda_library_through_compat(built, tmp_path):
    """whole_sweep_both.py / whole_sweep_both_maps.py / run_many_seeds.py, byte-identical copies, with compat/ on PYTHONPATH: every
    import and data path resolves and the first `wc.run()` lands in the CUDA library — which, on a box without a GPU, must refuse
    (NremError) instead of computing anything on the CPU.  (With a GPU the same scripts produce rows: tests/test_gpu_drivers.py.)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_drivers.py")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_drivers as drv
    cwd = drv._workdir(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), SLURM_ARRAY_TASK_ID="0", SLURM_ARRAY_TASK_MAX="199")
    procs = [subprocess.Popen([sys.executable, s], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for s in ("whole_sweep_both.py", "whole_sweep_both_maps.py", "run_many_seeds.py")]
    for pr in procs:
        out, _ = pr.communicate(timeout=300)
        assert pr.returncode != 0
        assert "NremError: no CUDA device visible" in out and "wc.run()" in out, out[-1500:]


@pytest.mark.skipif(not _staged(), reason="baseline/_ref is not staged")
def test_reference_arm_runs_the_unmodified_numba_module(built):
    """bench.py --impl reference / cpu_baseline: the reference's own netwWilsonCowanPlastic.py (numba) from baseline/_ref, with the
    oracle's shims for its two absent dependencies; and the oracle's restatement agrees with it on a seeded stream."""
    pytest.importorskip("numba")
    code = r'''
import sys, numpy as np, numba
sys.path.insert(0, %r)
import bench
from oracle import refshim, wc_oracle
bench._ref_init()
rec, run, tail = bench._ref_sim(0.001)
assert rec >= 0 and run > 0
wc = bench._REF["wc"]
assert wc.__file__.startswith(refshim.REF_DIR) and hasattr(wc.run, "recompile")
@numba.njit
def nseed(s):
    np.random.seed(s)
wc.timeTrans1, wc.timeTrans2, wc.timeSim, wc.time = np.arange(30.), np.arange(50.), np.arange(120.), np.arange(6.)
wc.run.recompile(); nseed(11)
Y = wc.run()
nz = np.random.RandomState(11).normal(0, 0.2, size=(200, 90))
Yo = wc_oracle.run(wc.CM, 0.16, 7.68, 30, 50, 120, noise=nz, p=wc_oracle.params(P=0.4, rhoE=0.18))
assert Y.shape == Yo.shape == (6, 3, 90) and np.max(np.abs(Y - Yo) / np.abs(Yo)) < 1e-11
print("ok")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_node_param_table_validation(built):
    """{name: vector} -> [NREM_NODE_PARAMS, N] table of the per-node parameter kernels (netwWilsonCowanPlastic.py:21): names not given keep
    the scalar of the parameter block, unknown names and wrong lengths are rejected before anything reaches the device."""
    from nremmodfc_b200 import ops
    p = ops.make_params(200, 1, 1, 20, P=0.37)
    t = ops.node_param_table(p, {"tauE": np.linspace(0.009, 0.011, 200), "a_ie_0": 3.0})
    assert t.shape == (len(ops.NODE_PARAMS), 200)
    assert np.all(t[ops.NODE_PARAMS.index("P")] == 0.37) and np.all(t[ops.NODE_PARAMS.index("a_ie_0")] == 3.0)
    assert t[ops.NODE_PARAMS.index("tauE"), -1] == 0.011
    with pytest.raises(ValueError):
        ops.node_param_table(p, {"G": np.ones(200)})              # G / sigmaE go through the maps, not through the table
    with pytest.raises(ValueError):
        ops.node_param_table(p, {"tauE": np.ones(90)})


def test_oracle_metrics_beyond_128_nodes():
    """The oracle side of the large-parcellation tests: FC / GoF restatements are size-agnostic (identity: corr 1, distance 0, SSIM 1)."""
    from oracle import bold_oracle
    rng = np.random.default_rng(3)
    fc = bold_oracle.fc(rng.normal(size=(298, 200)) + rng.normal(size=(298, 1)))
    assert fc.shape == (200, 200) and np.allclose(np.diag(fc), 1.0)
    assert np.allclose(bold_oracle.get_all_metrics(fc, fc), [1.0, 0.0, 1.0, 0.0], atol=1e-12)
