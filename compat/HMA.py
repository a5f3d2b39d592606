"""`import HMA` (run_many_seeds.py:13) -> nremmodfc_b200.HMA."""
import sys

import nremmodfc_b200.HMA as _m

sys.modules[__name__] = _m
