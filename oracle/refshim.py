"""Runs the UNMODIFIED reference modules (netwWilsonCowanPlastic.py, utils.py) from baseline/_ref.  TEST INFRASTRUCTURE ONLY.

baseline/_ref holds verbatim copies of the reference's files, made by `__graft_entry__.build()` when /root/reference is present
(git-ignored: the copies never enter the history; they travel to the GPU box with the snapshot).  Two imports of the reference
cannot be satisfied anywhere: `BOLDModel` (not in the reference tree, not on PyPI) and `skimage` (not
installed).  oracle/ref_shims/ provides them from the oracle's restatements so that the reference's own code runs unmodified
around them; bench.py's `--impl reference` / `cpu_baseline` leg uses this to time the reference's numba path on the host cores.
"""
import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(HERE, "ref_shims")
REFERENCE = "/root/reference"

# every file the reference's hot-path drivers read (whole_sweep_both.py:34-37, whole_sweep_both_maps.py:37-65, run_many_seeds.py:52-79)
FILES = ["netwWilsonCowanPlastic.py", "utils.py", "HMA.py", "whole_sweep_both.py", "whole_sweep_both_maps.py", "run_many_seeds.py",
         "SC_opti_25julio.txt"] + [f"empirical/mean_mat_{s}_8dic24.txt" for s in ("W", "N1", "N2", "N3")] + \
        [f"empirical/maps/{n}.npy" for n in ("DIST_VAChT_feobv_hc18_aghourian", "DIST_LC_proj", "SHUFFLED_SYMM_DIST_VAChT_feobv_hc18_aghourian",
                                            "SHUFFLED_SYMM_DIST_LC_proj")]


def stage_reference():
    """Copy the reference's files verbatim into baseline/_ref (no-op where /root/reference does not exist, e.g. on the GPU box).
    Returns the number of files copied."""
    if not os.path.isdir(REFERENCE):
        return 0
    n = 0
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    return n


def check_available():
    """Raises when the numba reference cannot run here (no staged copy, or numba missing)."""
    if not os.path.exists(os.path.join(REF_DIR, "netwWilsonCowanPlastic.py")):
        raise FileNotFoundError("baseline/_ref/netwWilsonCowanPlastic.py is not staged (run __graft_entry__.build() where /root/reference exists)")
    import numba  # noqa: F401


def import_reference():
    """-> (netwWilsonCowanPlastic, utils): the reference's own modules, imported from baseline/_ref."""
    check_available()
    for p in (SHIMS, REF_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [REF_DIR, SHIMS]
    for name in ("netwWilsonCowanPlastic", "utils", "BOLDModel", "skimage", "skimage.metrics"):
        sys.modules.pop(name, None)
    wc = importlib.import_module("netwWilsonCowanPlastic")
    utils = importlib.import_module("utils")
    assert os.path.dirname(os.path.abspath(wc.__file__)) == REF_DIR, wc.__file__
    return wc, utils
