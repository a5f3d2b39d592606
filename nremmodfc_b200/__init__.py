"""nremmodfc_b200 — B200-native Wilson-Cowan -> BOLD -> FC -> GoF hot path of NREMmodFC.

Layout
  csrc/                       hand-written sm_100a CUDA kernels + the C ABI (include/nremfc.h)
  _lib.py, ops.py             ctypes binding and NumPy-in / NumPy-out operators
  sweep.py                    batched G x sigma x seed x map sweeps, sharding and gather
  netwWilsonCowanPlastic.py   drop-in for the reference module of the same name
  BOLDModel.py, utils.py      drop-ins for the reference's BOLDModel.Sim and utils.get_all_metrics
"""
import os as _os

# sweeps with more tiles than SMs run one stream per tile group (csrc/nremfc_api.cu:integrate); give every
# stream its own hardware queue.  Only effective if set before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__all__ = ["ops", "sweep", "netwWilsonCowanPlastic", "BOLDModel", "utils"]
