"""ctypes binding of csrc/libnremfc.so (C ABI: include/nremfc.h).

There is no CPU fallback: importing works without a GPU (so the symbols can be checked), but every
compute call raises when no CUDA device is visible, and import fails loudly when the library has
not been built (``python -m nremmodfc_b200.build``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NREM_LIB_PATH: developer switch for kernel experiments (a variant build of the same C ABI, see build.build_variant)
LIB_PATH = os.environ.get("NREM_LIB_PATH") or os.path.join(_HERE, "csrc", "libnremfc.so")

ABI_SYMBOLS = [
    "nrem_abi_version", "nrem_last_error", "nrem_device_count", "nrem_wc_run_f64", "nrem_wc_run_f64_ex", "nrem_wc_derivative_f64",
    "nrem_bold_sim_f64", "nrem_filt_scratch_bytes", "nrem_filtfilt_decimate_f64", "nrem_fc_f64", "nrem_gof_f64", "nrem_kuramoto_f64",
    "nrem_sweep_create", "nrem_sweep_destroy", "nrem_sweep_device_bytes", "nrem_sweep_run",
    "nrem_sweep_set_node_params", "nrem_sweep_kernel", "nrem_sweep_begin", "nrem_sweep_chunks_total", "nrem_sweep_advance", "nrem_sweep_finish", "nrem_sweep_feed_samples",
    "nrem_sweep_integrate_f32", "nrem_sweep_integrate_f32_ex", "nrem_big_integrate_f32", "nrem_big_integrate_f32_ex", "nrem_launch_count", "nrem_selftest_tc_coupling", "nrem_measure_fma_peak", "nrem_last_integrate_ms", "nrem_sweep_set_profiling", "nrem_sweep_get_profile",
]


class NremError(RuntimeError):
    pass


class WCParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("a_ee", "a_ie_0", "a_ei", "a_ii", "tauE", "tauI", "P", "rhoE", "rE", "rI", "mu", "sigmaI",
                 "dtSim", "sqdtD", "E0", "I0")] + [
        ("tau_ip", C.c_double * 3), ("n1", C.c_int64), ("n2", C.c_int64), ("n3", C.c_int64),
        ("downsamp", C.c_int32), ("nnodes", C.c_int32), ("seed", C.c_uint64)]


class SweepOpts(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("bold_f32", C.c_int32), ("chunk_samples", C.c_int32), ("want_fc", C.c_int32),
                ("Neq", C.c_int64), ("bold_downsamp", C.c_int64), ("bold_dt", C.c_double),
                ("b", C.c_double * 5), ("a", C.c_double * 5),
                ("welch_nperseg", C.c_int32), ("reserved", C.c_int32), ("welch_fs", C.c_double)]


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -m nremmodfc_b200.build` "
                      "(there is no CPU fallback)")

lib = C.CDLL(LIB_PATH)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
lib.nrem_abi_version.restype = _i
lib.nrem_last_error.restype = C.c_char_p
lib.nrem_device_count.restype = _i
lib.nrem_launch_count.restype = _i64
lib.nrem_last_integrate_ms.restype = _d
lib.nrem_launch_count.argtypes = [_i]
lib.nrem_wc_run_f64.argtypes = [C.POINTER(WCParams), _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp, _vp, _vp]
lib.nrem_wc_run_f64_ex.argtypes = [C.POINTER(WCParams), _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp, _vp, _vp]
lib.nrem_wc_derivative_f64.argtypes = [C.POINTER(WCParams), _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp]
lib.nrem_bold_sim_f64.argtypes = [_vp, _i, _i64, _i, _d, _vp, _vp]
lib.nrem_filt_scratch_bytes.restype = _i64
lib.nrem_filt_scratch_bytes.argtypes = [_i, _i64, _i, _i64, _i64]
lib.nrem_filtfilt_decimate_f64.argtypes = [_vp, _i, _i64, _i, _i64, _i64, C.POINTER(_d), C.POINTER(_d), _vp, _vp, _vp]
lib.nrem_fc_f64.argtypes = [_vp, _i, _i64, _i, _vp, _vp]
lib.nrem_kuramoto_f64.argtypes = [_vp, _i, _i64, _i, _vp, _vp, _vp]
lib.nrem_gof_f64.argtypes = [_vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp]
lib.nrem_sweep_create.argtypes = [C.POINTER(WCParams), C.POINTER(SweepOpts), _i, _i, _i, C.POINTER(_vp)]
lib.nrem_sweep_destroy.argtypes = [_vp]
lib.nrem_sweep_device_bytes.restype = _i64
lib.nrem_sweep_device_bytes.argtypes = [_vp]
lib.nrem_sweep_run.argtypes = [_vp] + [_vp] * 7 + [C.POINTER(C.c_int32), _vp, _vp, _vp, _vp, _vp, _vp]
lib.nrem_sweep_begin.argtypes = [_vp] + [_vp] * 7 + [C.POINTER(C.c_int32), _vp, _i, _vp]
lib.nrem_sweep_set_node_params.argtypes = [_vp, _vp, _vp]
lib.nrem_sweep_kernel.argtypes = [_vp]
lib.nrem_sweep_chunks_total.restype = _i64
lib.nrem_sweep_chunks_total.argtypes = [_vp]
lib.nrem_sweep_advance.argtypes = [_vp, _i64, C.POINTER(_i64), _vp]
lib.nrem_sweep_finish.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
lib.nrem_sweep_feed_samples.argtypes = [_vp, _vp, _i64, _vp]
lib.nrem_big_integrate_f32.argtypes = [C.POINTER(WCParams), _i] + [_vp] * 8 + [_i, _i64, _vp, _vp, _vp, _vp]
lib.nrem_big_integrate_f32_ex.argtypes = [C.POINTER(WCParams), _i] + [_vp] * 9 + [_i, _i64, _vp, _vp, _vp, _vp]
lib.nrem_sweep_integrate_f32.argtypes = [C.POINTER(WCParams), _i] + [_vp] * 7 + [C.POINTER(C.c_int32), _vp, _i, _i, _i64, _vp, _vp, _vp]
lib.nrem_sweep_integrate_f32_ex.argtypes = [C.POINTER(WCParams), _i] + [_vp] * 7 + [C.POINTER(C.c_int32), _vp, _vp, _i, _i, _i64, _vp, _vp, _vp]
lib.nrem_measure_fma_peak.argtypes = [C.POINTER(_d), C.POINTER(_d)]
lib.nrem_sweep_set_profiling.argtypes = [_vp, _i]
lib.nrem_sweep_get_profile.argtypes = [_vp, C.POINTER(_d)]
lib.nrem_selftest_tc_coupling.argtypes = [_vp, _vp, _vp, _i] + [C.c_uint32] * 5 + [_vp]
for _n in ABI_SYMBOLS:
    getattr(lib, _n)          # AttributeError here = the library does not export what include/nremfc.h declares
    if _n not in ("nrem_last_integrate_ms", "nrem_last_error", "nrem_launch_count", "nrem_filt_scratch_bytes", "nrem_sweep_device_bytes", "nrem_sweep_chunks_total"):
        getattr(lib, _n).restype = _i

if lib.nrem_abi_version() != 2:
    raise ImportError("libnremfc.so ABI version mismatch")


def check(rc):
    if rc != 0:
        raise NremError(f"libnremfc error {rc}: {lib.nrem_last_error().decode()}")


def require_gpu():
    if lib.nrem_device_count() < 1:
        raise NremError("no CUDA device visible: nremmodfc_b200 has no CPU fallback")
