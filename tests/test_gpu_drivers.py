"""The reference's own driver scripts, UNMODIFIED, on the GPU path (north_star: "whole_sweep_both.py, whole_sweep_both_maps.py
and run_many_seeds.py drive it unchanged").

`__graft_entry__.build()` copies the three scripts and their data verbatim from /root/reference into the git-ignored
baseline/_ref/ (oracle/refshim.py:stage_reference); here byte-identical copies are started from a scratch directory, with compat/ on PYTHONPATH (module
names netwWilsonCowanPlastic / BOLDModel / utils / HMA / skimage resolve to the CUDA-backed drop-ins), the two SLURM variables
in the environment and a working directory that satisfies their relative paths.  Every script runs its first simulation at the
reference's FULL length (1 + 400 + 600 s, ~27 s through the float64 one-CTA kernel); the three processes run concurrently.
"""
import os
import pickle
import shutil
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
HEADER = "rank\tseed\tdelta_G\tdelta_sigma\tssimW\tssimN1\tssimN2\tssimN3\tcorrW\tcorrN1\tcorrN2\tcorrN3\teW\teN1\teN2\teN3\tsync\tmeta\tmean\tpeakfreq"


def _workdir(tmp_path):
    """tmp/analyze_empirical + tmp/a/cwd: run_many_seeds.py:79 reads ../../analyze_empirical/ relative to the CWD."""
    cwd = tmp_path / "a" / "cwd"
    os.makedirs(cwd / "output" / "temp")
    os.makedirs(cwd / "empirical")
    os.makedirs(tmp_path / "analyze_empirical")
    shutil.copy(os.path.join(REF, "SC_opti_25julio.txt"), cwd)
    for script in ("whole_sweep_both.py", "whole_sweep_both_maps.py", "run_many_seeds.py"):
        shutil.copy(os.path.join(REF, script), cwd)              # verbatim; run from a directory WITHOUT the reference's own modules,
                                                                 # so that `import netwWilsonCowanPlastic` resolves through compat/
    shutil.copytree(os.path.join(REF, "empirical", "maps"), cwd / "empirical" / "maps")
    for s in ("W", "N1", "N2", "N3"):
        src = os.path.join(REF, "empirical", f"mean_mat_{s}_8dic24.txt")
        shutil.copy(src, cwd / "empirical")                                     # whole_sweep_both.py:37
        shutil.copy(src, cwd)                                                   # whole_sweep_both_maps.py:41 (no empirical/ prefix)
        shutil.copy(src, tmp_path / "analyze_empirical" / f"mean_arctanhrho_filtered_{s}.txt")   # run_many_seeds.py:79 (loaded, never used)
    return cwd


def _wait_for_rows(path, proc, timeout):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if os.path.isfile(path):
            lines = open(path).read().splitlines()
            if len(lines) >= 2 and lines[1].count("\t") == 19:
                return lines
        if proc.poll() is not None:
            break
        time.sleep(1.0)
    return open(path).read().splitlines() if os.path.isfile(path) else []


def test_unmodified_reference_drivers_run_on_the_gpu_path(tmp_path):
    if not os.path.isfile(os.path.join(REF, "whole_sweep_both.py")):
        pytest.skip("baseline/_ref is not staged (build() copies it where /root/reference exists)")
    cwd = _workdir(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), SLURM_ARRAY_TASK_ID="0")
    logs = {n: open(tmp_path / f"{n}.log", "w") for n in ("wsb", "wsm", "rms")}
    # rank 0 of 200 ranks: run_many_seeds.py then runs exactly one (seed, state) and writes its pickle
    procs = {
        "wsb": subprocess.Popen([sys.executable, "whole_sweep_both.py"], cwd=cwd, env=dict(env, SLURM_ARRAY_TASK_MAX="0"),
                                stdout=logs["wsb"], stderr=subprocess.STDOUT),
        "wsm": subprocess.Popen([sys.executable, "whole_sweep_both_maps.py"], cwd=cwd, env=dict(env, SLURM_ARRAY_TASK_MAX="0"),
                                stdout=logs["wsm"], stderr=subprocess.STDOUT),
        "rms": subprocess.Popen([sys.executable, "run_many_seeds.py"], cwd=cwd, env=dict(env, SLURM_ARRAY_TASK_MAX="199"),
                                stdout=logs["rms"], stderr=subprocess.STDOUT),
    }
    try:
        out_wsb = cwd / "output" / "temp" / "sweep_delta_homoW_fromG0.16_sigma7.68_maps_0_0_9dic24_50iter_from0_rank0"
        out_wsm = cwd / "output" / "temp" / "sweep_deltaSHUFFLED_from_homoW_fromG0.16_sigma7.68_maps_2_2_9dic24_25iter_from25_rank0.txt"
        rows = {"wsb": _wait_for_rows(out_wsb, procs["wsb"], 420), "wsm": _wait_for_rows(out_wsm, procs["wsm"], 420)}
        rc = procs["rms"].wait(timeout=420)
    finally:
        for p in procs.values():
            if p.poll() is None:
                p.kill()
        for f in logs.values():
            f.close()
    tail = lambda n: open(tmp_path / f"{n}.log").read()[-2000:]
    for n in ("wsb", "wsm"):
        assert len(rows[n]) >= 2, tail(n)
        assert rows[n][0] == HEADER
        v = rows[n][1].split("\t")
        assert len(v) == 20 and v[0] == "0" and v[1] == ("0" if n == "wsb" else "25")
        x = np.array([float(t) for t in v[2:]])
        assert np.isfinite(x).all()
        assert x[0] == -0.1 and x[1] == (-1.0 if n == "wsb" else -0.2)            # first grid cell of each script
        assert np.all(np.abs(x[2:10]) <= 1.0) and np.all(x[10:14] > 0)           # ssim, corr in [-1, 1]; distances positive
        assert 0 < x[14] <= 1 and x[15] >= 0 and 0 <= x[17] <= 250               # sync, meta, peakfreq
        assert all(len(t.split(".")[1]) == 4 for t in v[2:])                       # :.4f formatting
    assert rc == 0, tail("rms")
    with open(cwd / "output" / "temp" / "run_50seeds_output_map_16dic_rank0.pickle", "rb") as f:
        save = pickle.load(f)
    assert list(save) == [(0, "W")]
    rec = save[(0, "W")]
    assert set(rec) == {"Hin_sim", "Hse_sim", "Hin_node_sim", "Hse_node_sim", "sFC"}
    assert rec["sFC"].shape == (90, 90) and rec["sFC"].min() >= 0.0 and np.allclose(np.diag(rec["sFC"]), 1.0)
    assert 0 < rec["Hin_sim"] < 1 and rec["Hse_sim"] > 0 and rec["Hin_node_sim"].shape == (90,)
