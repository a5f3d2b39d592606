"""Drop-in for the reference's external ``BOLDModel`` module (call site netwWilsonCowanPlastic.py:144).

The original is not part of the reference tree; the Balloon-Windkessel
form implemented by the CUDA library is stated in csrc/bold_filter.cuh.
"""
import numpy as np

from . import ops


def Sim(rE, nnodes, dt):
    """rE [T, nnodes] -> BOLD [T, nnodes] float64, one explicit-Euler step of size dt per row."""
    rE = np.asarray(rE, dtype=np.float64)
    if rE.ndim != 2 or rE.shape[1] != nnodes:
        raise ValueError(f"rE must be [T, {nnodes}], got {rE.shape}")
    return ops.bold_sim(rE, dt)
