"""Drop-in for the GoF part of the reference's utils.py (lines 18-50)."""
import numpy as np
from scipy import signal

from . import ops


def cohen_d(x, y):
    """utils.py:18-22."""
    nx, ny = len(x), len(y)
    dof = nx + ny - 2
    return (np.mean(x) - np.mean(y)) / np.sqrt(((nx - 1) * np.std(x, ddof=1) ** 2 + (ny - 1) * np.std(y, ddof=1) ** 2) / dof)


def flat_FC(FC):
    """utils.py:24-26 — row-major strict upper triangle."""
    FC = np.asarray(FC)
    return FC[np.triu_indices(len(FC), k=1)]


def get_all_metrics(sFC, empFC, data_range=1):
    """utils.py:42-50 -> (corr, euc, ssim, new_metric), computed by the gof kernel."""
    g, _ = ops.gof(sFC, empFC, data_range)
    return tuple(float(v) for v in g[0, 0])


def new_metric(flat1, flat2):
    """utils.py:28-31."""
    flat1, flat2 = np.asarray(flat1, dtype=np.float64), np.asarray(flat2, dtype=np.float64)
    return float(1 - np.corrcoef(flat1, flat2)[0, 1] + (flat1.mean() - flat2.mean()) ** 2)


def kuramoto(sign):
    """utils.py:34-40 — Kuramoto order parameter of the Hilbert phases: (mean, std) over time (kuramoto kernel)."""
    sign = np.asarray(sign, dtype=np.float64)
    if sign.ndim != 2:
        raise ValueError("sign must be [time, nodes]")
    if sign.shape[0] > 1024:                       # longer than the kernel's one-thread-per-time-point layout: SciPy on the host
        k = np.abs(np.mean(np.exp(1j * np.angle(signal.hilbert(sign, axis=0))), axis=1))
        return float(k.mean()), float(k.std())
    return ops.kuramoto(sign)


# node groups of the reference (utils.py:52-57)
thal = [38, 51]
subL, subR = [35, 36, 37, 38], [51, 52, 53, 54]
sub = subL + subR
cortex = [i for i in range(90) if i not in sub]
