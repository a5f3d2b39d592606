"""ctypes loader for oracle/liboracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import wc_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class WCParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("a_ee", "a_ie_0", "a_ei", "a_ii", "tauE", "tauI", "P", "rhoE", "rE", "rI", "mu", "sigmaI",
                 "dtSim", "sqdtD", "E0", "I0")] + [("tau_ip", C.c_double * 3)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.orc_wc_run.restype = C.c_int
        L.orc_wc_run.argtypes = [C.POINTER(WCParams), dp, C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                 C.c_int64, dp, C.c_uint64, C.c_uint64, dp, dp, dp, C.c_int]
        L.orc_bold_sim.restype = C.c_int
        L.orc_bold_sim.argtypes = [dp, C.c_int64, C.c_int, C.c_double, dp]
        L.orc_philox_normals.restype = None
        L.orc_philox_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, dp]
        L.orc_philox4x32.restype = None
        L.orc_philox4x32.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_uint32)]
        _LIB = L
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def c_params(p):
    cp = WCParams()
    for n, _ in WCParams._fields_:
        if n == "tau_ip":
            cp.tau_ip = (C.c_double * 3)(*p["tau_ip"])
        else:
            setattr(cp, n, float(p[n]))
    return cp


def wc_run(CM, G, sigmaE, n1, n2, n3, nrec=None, noise=None, seed=0, stream=0, p=None, want="Y", fast_rng=False):
    """One simulation.  want: "Y" -> [nrec,3,N], "E" -> [nrec,N], "final" -> [3,N].
    fast_rng=True (timed CPU baseline only) draws the noise with xoshiro + polar method instead of Philox."""
    p = wc_oracle.params() if p is None else p
    CM = np.ascontiguousarray(CM, dtype=np.float64)
    N = CM.shape[0]
    G = np.ascontiguousarray(np.broadcast_to(np.asarray(G, dtype=np.float64), (N,)))
    sg = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmaE, dtype=np.float64), (N,)))
    ds = int(p["dt"] / p["dtSim"])
    if nrec is None:
        nrec = (n3 + ds - 1) // ds
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        assert noise.shape == (n1 + n2 + n3, N)
    Y = np.zeros((nrec, 3, N)) if want == "Y" else None
    Eo = np.zeros((nrec, N)) if want == "E" else None
    fin = np.zeros((3, N))
    cp = c_params(p)
    rc = lib().orc_wc_run(C.byref(cp), _dp(CM), N, _dp(G), _dp(sg), n1, n2, n3, ds, nrec, _dp(noise),
                          int(seed), int(stream), _dp(Y), _dp(Eo), _dp(fin), int(bool(fast_rng)))
    assert rc == 0
    return {"Y": Y, "E": Eo, "final": fin}[want]


def bold_sim(rE, dt=0.04):
    rE = np.ascontiguousarray(rE, dtype=np.float64)
    out = np.empty_like(rE)
    lib().orc_bold_sim(_dp(rE), rE.shape[0], rE.shape[1], float(dt), _dp(out))
    return out


def philox_normals(seed, stream, step, N):
    z = np.empty(N)
    lib().orc_philox_normals(int(seed), int(stream), int(step), N, _dp(z))
    return z
