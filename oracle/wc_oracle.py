"""CPU restatement (numpy, float64) of the reference Wilson-Cowan integrator.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/netwWilsonCowanPlastic.py:
  * S(x, sigma, mu)                     lines 72-74
  * wilsonCowan(t, X, sigmaE, mu, tau_ip, G)   lines 77-83
  * run(verbose)                        lines 86-137  (three phases, explicit
    Euler-Maruyama, state stored BEFORE the update when i % downsamp == 0)
with the parameter defaults of lines 23-68.

Differences from the reference, all deliberate:
  * parameters are passed explicitly instead of being frozen numba globals;
  * a leading batch axis is allowed (state is [B, N]) so that a few
    simulations can be advanced together;
  * the noise is an argument: either an injected array (the reference's own
    stream is bit-identical to RandomState(s).normal(0, sqdtD, (steps, N)),
    SURVEY.md item 3) or the counter-based stream of oracle/philox.py.
"""
import numpy as np

from . import philox

DEFAULTS = dict(
    a_ee=3.5, a_ie_0=2.5, a_ei=3.75, a_ii=0.0,      # netwWilsonCowanPlastic.py:23-25
    tauE=0.010, tauI=0.020,                         # :27
    P=0.4,                                          # :29
    rhoE=0.14,                                      # :32
    rE=0.5, rI=0.5, mu=1.0, sigmaE=4.0, sigmaI=4.0,  # :35-38
    dt=0.002, dtSim=0.0001,                         # :45-46
    D=0.002,                                        # :55  (sqdtD = D/sqrt(dtSim) = 0.2, :56)
    G=0.7,                                          # :61
    E0=0.1, I0=0.1,                                 # :90-91
    tau_ip=(0.05, 1.0, 2.0),                        # :101, :111, :118
)


def S(x, sigma, mu):
    """netwWilsonCowanPlastic.py:72-74."""
    return 1.0 / (1.0 + np.exp(-(x - mu) * sigma))


def params(**over):
    p = dict(DEFAULTS)
    p.update(over)
    p.setdefault("sqdtD", p["D"] / np.sqrt(p["dtSim"]))
    return p


def derivative(E, I, a_ie, CM, G, sigmaE, noise, tau_ip, p):
    """netwWilsonCowanPlastic.py:77-83.  E, I, a_ie, noise: [..., N]; G, sigmaE scalar or [..., N]."""
    coup = E @ CM.T                       # np.dot(CM, E) per simulation
    dE = (-E + (1 - p["rE"] * E) * S(p["a_ee"] * E - a_ie * I + G * coup + p["P"] + noise, sigmaE, p["mu"])) / p["tauE"]
    dI = (-I + (1 - p["rI"] * I) * S(p["a_ei"] * E - p["a_ii"] * I, p["sigmaI"], p["mu"])) / p["tauI"]
    da = (I * (E - p["rhoE"])) / tau_ip
    return dE, dI, da


def run(CM, G, sigmaE, n1, n2, n3, nrec=None, noise=None, seed=0, streams=None, p=None, return_final=False):
    """netwWilsonCowanPlastic.py:86-137.

    CM      [N, N] float64
    G, sigmaE  scalar, [N] or [B, N]
    n1,n2,n3  len(timeTrans1), len(timeTrans2), len(timeSim)
    nrec    len(time) (rows of Y_t); default ceil(n3 / downsamp)
    noise   None -> philox stream (seed, streams[B]); else array [n1+n2+n3, N] or [B, n1+n2+n3, N]
            already scaled by sqdtD (i.e. exactly what np.random.normal(0, sqdtD, N) returned).
    Returns Y_t [B, nrec, 3, N] (or [nrec, 3, N] when no batch axis was implied).
    """
    p = params() if p is None else p
    CM = np.asarray(CM, dtype=np.float64)
    N = CM.shape[0]
    G = np.asarray(G, dtype=np.float64)
    sigmaE = np.asarray(sigmaE, dtype=np.float64)
    batched = (G.ndim == 2) or (sigmaE.ndim == 2) or (noise is not None and np.ndim(noise) == 3) or (streams is not None)
    if streams is not None:
        B = len(streams)
    elif G.ndim == 2:
        B = G.shape[0]
    elif sigmaE.ndim == 2:
        B = sigmaE.shape[0]
    elif noise is not None and np.ndim(noise) == 3:
        B = noise.shape[0]
    else:
        B = 1
    if streams is None:
        streams = np.arange(B, dtype=np.uint64)
    streams = np.asarray(streams, dtype=np.uint64)
    downsamp = int(p["dt"] / p["dtSim"])                    # :120  (int() truncation as in the reference)
    if nrec is None:
        nrec = (n3 + downsamp - 1) // downsamp
    dtSim = p["dtSim"]
    E = np.full((B, N), p["E0"])
    I = np.full((B, N), p["I0"])
    a = np.full((B, N), p["a_ie_0"])
    Y = np.zeros((B, nrec, 3, N))
    step = 0
    for phase, nsteps in enumerate((n1, n2, n3)):
        tau_ip = p["tau_ip"][phase]
        for i in range(nsteps):
            if phase == 2 and i % downsamp == 0:            # :129-130
                Y[:, i // downsamp, 0] = E
                Y[:, i // downsamp, 1] = I
                Y[:, i // downsamp, 2] = a
            if noise is None:
                nz = p["sqdtD"] * philox.normals(seed, streams, step, N)
            else:
                nz = noise[..., step, :]
            dE, dI, da = derivative(E, I, a, CM, G, sigmaE, nz, tau_ip, p)
            E = E + dtSim * dE                              # Var += dtSim*wilsonCowan(...)
            I = I + dtSim * dI
            a = a + dtSim * da
            step += 1
    if return_final:
        final = np.stack([E, I, a], axis=1)
        return (Y if batched else Y[0]), (final if batched else final[0])
    return Y if batched else Y[0]
