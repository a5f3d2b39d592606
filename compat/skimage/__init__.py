"""Minimal stand-in so that `from skimage.metrics import structural_similarity` (whole_sweep_both.py:15,
utils.py:14) resolves when scikit-image is not installed; the SSIM itself is computed by the gof kernel."""
from . import metrics  # noqa: F401
