"""`import BOLDModel as BD` -> nremmodfc_b200.BOLDModel."""
import sys

import nremmodfc_b200.BOLDModel as _m

sys.modules[__name__] = _m
