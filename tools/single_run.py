"""BASELINE configs[0]: one full-length Wilson-Cowan + BOLD run on AAL90 (G = 0.16, sigma = 7.68, one seed) through the
drop-in module surface (set attributes, run(), simBOLD(), corrcoef, get_all_metrics) — timing of the compatibility path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT]
import netwWilsonCowanPlastic as wc  # noqa: E402
import utils  # noqa: E402
from scipy import signal  # noqa: E402

d = np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"))
wc.P, wc.rhoE, wc.CM = 0.4, 0.18, d["SC"]
wc.tTrans1, wc.tTrans2, tstop = 1, 400, 600
wc.timeTrans1 = np.arange(0, wc.tTrans1, wc.dtSim)
wc.timeTrans2 = np.arange(0, wc.tTrans2, wc.dtSim)
wc.tstop = tstop
wc.timeSim = np.arange(0, tstop, wc.dtSim)
wc.time = np.arange(0, tstop, wc.dt)
wc.G, wc.sigmaE, wc.sid = 0.16, 7.68, int(sys.argv[1]) if len(sys.argv) > 1 else 0
t0 = time.perf_counter()
wc.run.recompile()
tray = wc.run()
t1 = time.perf_counter()
E_t = tray[:, 0, :]
BOLD = wc.simBOLD(E_t, nnodes=90)
sFC = np.corrcoef(BOLD.T)
m = utils.get_all_metrics(sFC, d["W"], data_range=1)
sync, meta = utils.kuramoto(BOLD)
t2 = time.perf_counter()
freqs, p = signal.welch(E_t.T, fs=1 / wc.dt, nperseg=4000)
peak = freqs[np.argmax(p.mean(axis=0))]
print(f"run() {t1 - t0:.2f} s (Y_t {tray.nbytes / 1e6:.0f} MB), simBOLD+FC+GoF+kuramoto {t2 - t1:.2f} s")
print(f"corrW {m[0]:.4f} (ref 0.474+-0.023)  eW {m[1]:.3f} (8.39+-0.28)  ssimW {m[2]:.4f} (0.416+-0.026)  mean {sFC.mean():.4f} (0.483+-0.020)  "
      f"sync {sync:.4f} (0.600+-0.016)  meta {meta:.4f} (0.213+-0.008)  peak {peak:.3f} Hz (4.81+-0.06)")
