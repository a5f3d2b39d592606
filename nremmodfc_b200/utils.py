"""Drop-in for the GoF part of the reference's utils.py (lines 18-50)."""
import numpy as np

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
