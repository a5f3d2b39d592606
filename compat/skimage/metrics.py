import numpy as np

from nremmodfc_b200 import ops


def structural_similarity(im1, im2, data_range=None, **kw):
    """scikit-image defaults only (7x7 uniform window, sample covariance, K1=0.01, K2=0.03) — utils.py:48."""
    if kw:
        raise NotImplementedError(f"only the default structural_similarity is implemented, got {sorted(kw)}")
    if data_range is None:
        raise ValueError("data_range must be given for float images")
    g, _ = ops.gof(np.asarray(im1, dtype=np.float64), np.asarray(im2, dtype=np.float64), float(data_range))
    return float(g[0, 0, 2])
