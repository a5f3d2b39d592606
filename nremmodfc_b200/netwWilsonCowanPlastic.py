"""Drop-in for the reference module netwWilsonCowanPlastic.py.

Same surface: set module attributes, ``run.recompile()``, ``run()``, ``simBOLD()``
(reference lines 23-68 for the attributes, 86-137 ``run``, 140-158 ``simBOLD``), so
whole_sweep_both.py / whole_sweep_both_maps.py / run_many_seeds.py drive it unchanged.  Underneath
every call goes to the CUDA library; there is no numba and no CPU path.

Differences (all supersets of the reference behaviour):
  * attributes are read at CALL time; ``recompile()`` is a no-op kept for compatibility
    (numba froze globals at JIT time and the drivers paid ~3 s of LLVM per simulation);
  * the noise stream is counter-based Philox keyed by (``sid``, replicate).  By default (``replicate = None``) the replicate id
    is the number of ``run()`` calls made so far in this process, so consecutive calls draw fresh, independent noise exactly
    as the reference does (its numba generator is never re-seeded: ``sid`` is a replicate label there) —
    the 400 (dG, dsigma) cells a driver runs for one ``sid`` are statistically independent.  Set ``replicate`` to an integer to
    pin the stream (same ``sid`` + ``replicate`` -> same trajectory), or ``noise`` to an array [steps, N] to inject one;
  * ``run()`` takes a length-N vector for any node parameter (reference line 21); the batched sweep (``sweep.SweepPlan``)
    takes per-node ``G`` and ``sigmaE``, the only ones the drivers vary per node.
"""
import gc

import numpy as np

from . import ops
from . import BOLDModel as BD
from .sweep import bandpass_ba

###  MODEL PARAMETERS  (reference lines 23-38)
a_ee = 3.5
a_ie_0 = 2.5
a_ei = 3.75
a_ii = 0
tauE = 0.010
tauI = 0.020
P = 0.4
Q = 0
rhoE = 0.14
tau_ip = 2
rE, rI = 0.5, 0.5
mu = 1
sigmaE = 4
sigmaI = 4

### Time units are seconds (reference lines 41-52)
tTrans1 = 600
tTrans2 = 600
tstop = 600
dt = 0.002
dtSim = 0.0001
downsamp = int(dt / dtSim)
timeTrans1 = np.arange(0, tTrans1, dtSim)
timeTrans2 = np.arange(0, tTrans2, dtSim)
timeSim = np.arange(0, tstop, dtSim)
time = np.arange(0, tstop, dt)

# Noise factor (reference lines 55-59)
D = 0.002
sqdtD = D / np.sqrt(dtSim)
sid = 12
noise = None            # extension: inject [len(timeTrans1)+len(timeTrans2)+len(timeSim), N] values (already scaled)
replicate = None        # extension: Philox stream id of the next run; None = number of run() calls so far (fresh noise per call)
_calls = 0

# network parameters (reference lines 61-68)
G = 0.7
CM = np.random.RandomState(sid).uniform(size=(90, 90))
nnodes = len(CM)
N = len(CM)


def S(x, sigma, mu):
    """Reference lines 72-74."""
    return 1 / (1 + np.exp(-(np.asarray(x, dtype=np.float64) - mu) * sigma))


def _node_vectors(nn):
    """Attributes that were redefined as length-N vectors (reference line 21) -> {name: vector}."""
    g = globals()
    out = {}
    for name in ops.NODE_PARAMS:
        v = np.asarray(g[name], dtype=np.float64)
        if v.ndim == 1 and v.shape[0] == nn:
            out[name] = v
        elif v.ndim != 0:
            raise ValueError(f"{name} must be a scalar or a vector of length {nn}, got shape {v.shape}")
    return out


def _params(n1, n2, n3, nn):
    g = globals()
    sc = {name: float(np.mean(g[name])) for name in ops.NODE_PARAMS}        # vectors travel separately (_node_vectors)
    return ops.make_params(nn, n1, n2, n3, dtSim=g["dtSim"], sqdtD=g["sqdtD"],
                           downsamp=int(g["dt"] / g["dtSim"]), seed=int(g["sid"]), **sc)


class _Recompilable:
    def recompile(self):
        """No-op: parameters are read at call time (the reference re-JITs here, whole_sweep_both.py:75)."""
        return None


class _Run(_Recompilable):
    def __call__(self, verbose=False):
        """Reference lines 86-137 -> Y_t [len(time), 3, N] float64 (E, I, a_ie before every downsamp-th step)."""
        g = globals()
        cm = np.asarray(g["CM"], dtype=np.float64)
        nn = len(cm)
        p = _params(len(g["timeTrans1"]), len(g["timeTrans2"]), len(g["timeSim"]), nn)
        rep = g["_calls"] if g["replicate"] is None else int(g["replicate"])
        g["_calls"] += 1
        Y, _ = ops.wc_run(p, cm, g["G"], g["sigmaE"], B=1, streams=[rep], noise=g["noise"],
                          nrec=len(g["time"]), want_Y=True, node_params=_node_vectors(nn) or None)
        return Y[0]


class _WilsonCowan(_Recompilable):
    def __call__(self, t, X, sigmaE, mu, tau_ip, G):
        """Reference lines 77-83: derivative of (E, I, a_ie) with a fresh noise draw."""
        g = globals()
        cm = np.asarray(g["CM"], dtype=np.float64)
        nn = len(cm)
        vec = _node_vectors(nn)
        if vec or np.ndim(mu) != 0:
            raise ValueError("wilsonCowan(): per-node vectors for " + ", ".join(sorted(vec) or ["mu"]) + " are only supported by run()")
        p = _params(0, 0, 0, nn)
        p.mu = float(mu)
        nz = np.random.normal(0, g["sqdtD"], size=nn)
        return ops.wc_derivative(p, cm, X, G, sigmaE, noise=nz, tau_ip=tau_ip)


run = _Run()
wilsonCowan = _WilsonCowan()


def simBOLD(E_t, nnodes=90, BOLD_downsamp=1000):
    """Reference lines 140-158: E_t [T, nnodes] -> filtered, decimated BOLD."""
    g = globals()
    E_t = np.ascontiguousarray(E_t, dtype=np.float64)
    if E_t.ndim != 2 or E_t.shape[1] != nnodes:
        raise ValueError(f"E_t must be [T, {nnodes}], got {E_t.shape}")
    BOLD_dt = g["dt"] * g["downsamp"]
    BOLD_signals = BD.Sim(E_t, nnodes, BOLD_dt)
    b, a = bandpass_ba(BOLD_dt)
    BOLD = ops.filtfilt_decimate(BOLD_signals, b, a, Neq=2000, ds=int(BOLD_downsamp))
    gc.collect()
    return BOLD
