"""`import utils` -> nremmodfc_b200.utils (GoF part of the reference's utils.py)."""
import sys

import nremmodfc_b200.utils as _m

sys.modules[__name__] = _m
