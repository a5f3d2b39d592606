"""`import netwWilsonCowanPlastic as wc` -> nremmodfc_b200.netwWilsonCowanPlastic (same module object)."""
import sys

import nremmodfc_b200.netwWilsonCowanPlastic as _m

sys.modules[__name__] = _m
