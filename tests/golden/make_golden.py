"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, numba, scipy):

    python tests/golden/make_golden.py

What it does
------------
* ../../data/aal90_inputs.npz — the reference's DATA inputs (SC, 4 empirical FCs, NA/ACh maps); input data, not a golden output.
* wc_short_*.npz        — trajectories of the reference's own ``run()`` (numba) with its
                          noise stream seeded inside numba (``np.random.seed`` in an @njit
                          function; that stream equals RandomState(s).normal(0, 0.2, (steps, N))
                          row by row — asserted below).
* chain_*.npz           — reference ``run()`` -> ``simBOLD`` -> ``np.corrcoef`` ->
                          ``utils.get_all_metrics`` / ``utils.kuramoto`` / welch on a 1+2+17 s run (20 s: FP64 summation-order chaos stays < 1e-9 that long).
* sweep_cell_stats.npz  — per-cell mean / sd / n of the reference's committed sweep tables
                          (output/*.txt), the statistical pin of the whole pipeline.

Two imports of the reference cannot be satisfied: ``BOLDModel`` (not in the
repo, not on PyPI) and ``skimage``.  They are shimmed with the oracle's restatements
(oracle/bold_oracle.py) so the reference's own code runs unmodified around them.  Consequently
the chain fixture pins WC integration, cut/filter/decimate, corrcoef, corr/euclid/new_metric
against reference code, and BOLDModel/SSIM only against the oracle itself (parity unpinned).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import bold_oracle, cwrap  # noqa: E402

# ---- shims for the two unsatisfiable imports -------------------------------------------------
bd = types.ModuleType("BOLDModel")
bd.Sim = lambda rE, nnodes, dt: cwrap.bold_sim(np.ascontiguousarray(rE), dt)
sys.modules["BOLDModel"] = bd
sk = types.ModuleType("skimage")
skm = types.ModuleType("skimage.metrics")
skm.structural_similarity = lambda a, b, data_range=1.0: bold_oracle.ssim(a, b, data_range)
sk.metrics = skm
sys.modules["skimage"] = sk
sys.modules["skimage.metrics"] = skm

sys.path.insert(0, REF)
import numba  # noqa: E402
import netwWilsonCowanPlastic as wc  # noqa: E402  (the unmodified reference module)
import utils as refutils  # noqa: E402
from scipy import signal  # noqa: E402


@numba.njit
def nseed(s):
    np.random.seed(s)


@numba.njit
def ndraw(n, N):
    out = np.empty((n, N))
    for i in range(n):
        out[i] = np.random.normal(0, 0.2, size=N)
    return out


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KB")


# ---- 1. data inputs -------------------------------------------------------------------------
struct = np.loadtxt(f"{REF}/SC_opti_25julio.txt")
emp = {s: np.loadtxt(f"{REF}/empirical/mean_mat_{s}_8dic24.txt") for s in ("W", "N1", "N2", "N3")}
maps = {n: np.load(f"{REF}/empirical/maps/{n}.npy").astype(np.float64) for n in
        ("DIST_VAChT_feobv_hc18_aghourian", "DIST_LC_proj",
         "SHUFFLED_SYMM_DIST_VAChT_feobv_hc18_aghourian", "SHUFFLED_SYMM_DIST_LC_proj")}
save(os.path.join("..", "..", "data", "aal90_inputs.npz"), SC=struct, W=emp["W"], N1=emp["N1"], N2=emp["N2"], N3=emp["N3"],
     map_ACh=maps["DIST_VAChT_feobv_hc18_aghourian"], map_NA=maps["DIST_LC_proj"],
     map_ACh_shuf=maps["SHUFFLED_SYMM_DIST_VAChT_feobv_hc18_aghourian"], map_NA_shuf=maps["SHUFFLED_SYMM_DIST_LC_proj"])

# ---- noise-stream equivalence (numba seeded stream == RandomState) ---------------------------
nseed(5)
assert np.array_equal(ndraw(50, 90), np.random.RandomState(5).normal(0, 0.2, size=(50, 90)))
print("numba stream == RandomState stream: OK")


def set_times(t1, t2, tstop):
    wc.tTrans1, wc.tTrans2, wc.tstop = t1, t2, tstop                 # whole_sweep_both.py:43-50
    wc.timeTrans1 = np.arange(0, t1, wc.dtSim)
    wc.timeTrans2 = np.arange(0, t2, wc.dtSim)
    wc.timeSim = np.arange(0, tstop, wc.dtSim)
    wc.time = np.arange(0, tstop, wc.dt)
    return len(wc.timeTrans1), len(wc.timeTrans2), len(wc.timeSim), len(wc.time)


wc.P = 0.4                                                           # whole_sweep_both.py:39-41
wc.rhoE = 0.18
wc.CM = struct

# ---- 2. short trajectories ------------------------------------------------------------------
m1 = maps["DIST_VAChT_feobv_hc18_aghourian"] / maps["DIST_VAChT_feobv_hc18_aghourian"].mean()
m2 = maps["DIST_LC_proj"] / maps["DIST_LC_proj"].mean()
cases = {
    "homo": dict(seed=7, G=0.16, sigmaE=7.68),
    "map": dict(seed=11, G=0.16 + 0.18 * m1, sigmaE=7.68 - 0.02 * m2),   # whole_sweep_both_maps.py:104-108
}
for name, c in cases.items():
    n1, n2, n3, nrec = set_times(0.05, 0.45, 1.0)
    wc.G, wc.sigmaE = c["G"], c["sigmaE"]
    wc.run.recompile()
    nseed(c["seed"])
    Y = wc.run()
    assert Y.shape == (nrec, 3, 90)
    save(f"wc_short_{name}.npz", Y=Y[::5], rows=np.arange(nrec)[::5], seed=c["seed"], G=np.asarray(c["G"]),
         sigmaE=np.asarray(c["sigmaE"]), n=np.array([n1, n2, n3, nrec]), P=0.4, rhoE=0.18)

# ---- 3. full chain on a 1 + 2 + 17 s run ----------------------------------------------------
n1, n2, n3, nrec = set_times(1, 2, 17)
wc.G, wc.sigmaE = 0.16, 7.68
wc.run.recompile()
nseed(3)
tray = wc.run()
E_t = tray[:, 0, :]
BOLD = wc.simBOLD(E_t, nnodes=90, BOLD_downsamp=100)                 # netwWilsonCowanPlastic.py:140-158
sFC = np.corrcoef(BOLD.T)                                            # whole_sweep_both.py:81
gof = np.array([refutils.get_all_metrics(sFC, emp[s], data_range=1) for s in ("W", "N1", "N2", "N3")])
freqs, fftPow = signal.welch(E_t.T, fs=1 / wc.dt, nperseg=4000)      # whole_sweep_both.py:90-95
meanpow = fftPow.mean(axis=0)
peak = freqs[np.where(meanpow == meanpow.max())[0][0]]
sync, meta = refutils.kuramoto(BOLD)
save("chain_homo.npz", seed=3, n=np.array([n1, n2, n3, nrec]), G=0.16, sigmaE=7.68, P=0.4, rhoE=0.18,
     E_rows=E_t[::100], final=tray[-1], BOLD=BOLD, FC=sFC, gof=gof, mean=sFC.mean(), peakfreq=peak, sync=sync, meta=meta,
     BOLD_downsamp=100)
print("chain gof (corr, euc, ssim, new) x W,N1,N2,N3:\n", gof)

# ---- 4. committed sweep tables -> per-cell statistics ----------------------------------------
import pandas as pd  # noqa: E402

files = {
    "homo": "sweep_delta_homoW_fromG0.16_sigma7.68_maps_0_0_9dic24_50iter.txt",
    "map": "sweep_deltamaps_from_homoW_fromG0.16_sigma7.68_maps_1_1_9dic24_50iter.txt",
    "shuf": "sweep_deltaSHUFFLED_from_homoW_fromG0.16_sigma7.68_maps_2_2_9dic24_50iter.txt",
}
cols = ["ssimW", "ssimN1", "ssimN2", "ssimN3", "corrW", "corrN1", "corrN2", "corrN3", "eW", "eN1", "eN2", "eN3",
        "sync", "meta", "mean", "peakfreq"]
dG = np.round(np.linspace(-0.1, 0.3, 20, endpoint=False), 4)
dS = np.round(np.linspace(-0.2, 0.2, 20, endpoint=False), 4)
out = {}
for mod, fn in files.items():
    df = pd.read_csv(f"{REF}/output/{fn}")
    df = df.drop_duplicates(subset=["seed", "delta_G", "delta_sigma"])
    mean = np.full((20, 20, len(cols)), np.nan, dtype=np.float32)
    sd = np.full((20, 20, len(cols)), np.nan, dtype=np.float32)
    cnt = np.zeros((20, 20), dtype=np.int32)
    for (g, s), grp in df.groupby(["delta_G", "delta_sigma"]):
        i = int(np.argmin(np.abs(dG - g)))
        j = int(np.argmin(np.abs(dS - s)))
        assert abs(dG[i] - g) < 1e-6 and abs(dS[j] - s) < 1e-6
        mean[i, j] = grp[cols].mean().to_numpy()
        sd[i, j] = grp[cols].std(ddof=1).to_numpy()
        cnt[i, j] = len(grp)
    out[f"{mod}_mean"], out[f"{mod}_sd"], out[f"{mod}_n"] = mean, sd, cnt
    print(mod, "cells", (cnt > 0).sum(), "n range", cnt.min(), cnt.max(), "cell(0,0) eW", mean[5, 10, 8], sd[5, 10, 8])
save("sweep_cell_stats.npz", cols=np.array(cols), delta_G=dG, delta_sigma=dS, **out)

# ---- 5. HMA known answers (output/emp_15inds_output_16dic.pickle) ------------------------------
import pickle  # noqa: E402

import HMA as refHMA  # noqa: E402  (the reference's own, pure NumPy)

with open(f"{REF}/output/emp_15inds_output_16dic.pickle", "rb") as fh:
    emp_out = pickle.load(fh)
keys = sorted(emp_out.keys())[:: max(1, len(emp_out) // 4)][:4]
cn = [refHMA.Functional_HP(emp_out[k]["sFC"].copy())[0] for k in keys]       # also pins the module counts per level
save("hma_kat.npz", sFC=np.stack([emp_out[k]["sFC"] for k in keys]),
     Hin=np.array([emp_out[k]["Hin_sim"] for k in keys]), Hse=np.array([emp_out[k]["Hse_sim"] for k in keys]),
     Hin_node=np.stack([emp_out[k]["Hin_node_sim"] for k in keys]), Hse_node=np.stack([emp_out[k]["Hse_node_sim"] for k in keys]),
     keys=np.array([str(k) for k in keys]), Clus_num=np.array(cn))
