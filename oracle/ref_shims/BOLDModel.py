"""Stand-in for the lab's external `BOLDModel` module (call site netwWilsonCowanPlastic.py:144), which is not part of the reference
tree: the oracle's Balloon-Windkessel restatement (parity unpinned, see oracle/__init__.py).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cwrap  # noqa: E402


def Sim(rE, nnodes, dt):
    return cwrap.bold_sim(np.ascontiguousarray(rE, dtype=np.float64), dt)
