"""Host-side operators over the C ABI (include/nremfc.h).

PyTorch is used only as plumbing: device memory, pinned host buffers, streams.  Every function
takes and returns NumPy arrays (host) unless it says "device"; all arithmetic happens in the
CUDA library.  Reference call sites are cited per function.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SweepOpts, WCParams, check, lib

# defaults of netwWilsonCowanPlastic.py:23-57, :90-91, :101/:111/:118
WC_DEFAULTS = dict(a_ee=3.5, a_ie_0=2.5, a_ei=3.75, a_ii=0.0, tauE=0.010, tauI=0.020, P=0.4, rhoE=0.14,
                   rE=0.5, rI=0.5, mu=1.0, sigmaI=4.0, dtSim=1e-4, sqdtD=0.2, E0=0.1, I0=0.1,
                   tau_ip=(0.05, 1.0, 2.0), downsamp=20, seed=0)


def make_params(nnodes, n1, n2, n3, **over):
    d = dict(WC_DEFAULTS)
    unknown = set(over) - set(d)
    if unknown:
        raise ValueError(f"unknown Wilson-Cowan parameters: {sorted(unknown)}")
    d.update(over)
    p = WCParams()
    for k, v in d.items():
        if k == "tau_ip":
            if len(v) != 3:
                raise ValueError("tau_ip needs one value per phase (3)")
            p.tau_ip = (C.c_double * 3)(*[float(x) for x in v])
        elif k in ("downsamp",):
            p.downsamp = int(v)
        elif k == "seed":
            p.seed = int(v) & 0xFFFFFFFFFFFFFFFF
        else:
            if np.ndim(v) != 0:
                raise ValueError(f"{k} must be a scalar in nrem_wc_params (per-node vectors: node_params of wc_run; G and sigmaE everywhere)")
            setattr(p, k, float(v))
    p.nnodes, p.n1, p.n2, p.n3 = int(nnodes), int(n1), int(n2), int(n3)
    return p


def _device(device=None):
    _lib.require_gpu()
    return torch.device("cuda", torch.cuda.current_device() if device is None else device)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(x, dtype, dev):
    """Host array -> device tensor through a pinned staging buffer."""
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    a = np.ascontiguousarray(x)
    if not a.flags.writeable:          # e.g. arrays straight out of np.load(...): torch wants writable memory
        a = a.copy()
    h = torch.from_numpy(a)
    if h.dtype != dtype:
        h = h.to(dtype)
    if h.numel() > 0:
        h = h.pin_memory()
    return h.to(dev, non_blocking=True)


def _u64(x, dev):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64)).view(np.int64)
    return to_device(a, torch.int64, dev)


def _per_node(x, B, N, name):
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        a = np.full((B, N), float(a))
    elif a.ndim == 1:
        if a.shape[0] != N:
            raise ValueError(f"{name} must be a scalar or have length nnodes={N}, got {a.shape}")
        a = np.broadcast_to(a, (B, N))
    elif a.shape != (B, N):
        raise ValueError(f"{name} must have shape ({B}, {N}), got {a.shape}")
    return np.ascontiguousarray(a)


NODE_PARAMS = ("a_ee", "a_ei", "a_ii", "tauE", "tauI", "P", "rhoE", "rE", "rI", "mu", "sigmaI", "a_ie_0")     # NREM_NODE_PARAMS order


def node_param_table(p, node_params):
    """{name: scalar or length-N vector} -> [NREM_NODE_PARAMS, N] float64 table (names not given: the scalar of `p`)."""
    N = p.nnodes
    unknown = set(node_params) - set(NODE_PARAMS)
    if unknown:
        raise ValueError(f"not per-node parameters: {sorted(unknown)} (allowed: {NODE_PARAMS})")
    npar = np.empty((len(NODE_PARAMS), N))
    for k, name in enumerate(NODE_PARAMS):
        v = np.asarray(node_params.get(name, getattr(p, name)), dtype=np.float64)
        if v.ndim > 1 or (v.ndim == 1 and v.shape[0] != N):
            raise ValueError(f"{name} must be a scalar or a vector of length {N}")
        npar[k] = v
    return npar


def wc_run(p, CM, G, sigmaE, B=1, streams=None, noise=None, nrec=None, want_Y=True, node_params=None, device=None):
    """run() of netwWilsonCowanPlastic.py:86-137 for B simulations (float64).

    node_params: optional {name: length-N vector} for any of NODE_PARAMS ("Any of them can be redefined as a vector of
    length nnodes", netwWilsonCowanPlastic.py:21); names not given keep the scalar of `p`.
    Returns (Y [B, nrec, 3, N] or None, final [B, 3, N])."""
    dev = _device(device)
    N = p.nnodes
    CM = np.asarray(CM, dtype=np.float64)
    if CM.shape != (N, N):
        raise ValueError(f"CM must be ({N}, {N}), got {CM.shape}")
    G = _per_node(G, B, N, "G")
    sg = _per_node(sigmaE, B, N, "sigmaE")
    steps = p.n1 + p.n2 + p.n3
    if nrec is None:
        nrec = (p.n3 + p.downsamp - 1) // p.downsamp
    d_noise, nb = None, 1
    if noise is not None:
        noise = np.asarray(noise, dtype=np.float64)
        if noise.ndim == 2:
            noise = noise[None]
        if noise.shape[1:] != (steps, N) or noise.shape[0] not in (1, B):
            raise ValueError(f"noise must be [1 or {B}, {steps}, {N}], got {noise.shape}")
        nb = noise.shape[0]
        d_noise = to_device(noise, torch.float64, dev)
    npar = node_param_table(p, node_params) if node_params else None
    with torch.cuda.device(dev):
        d_CM, d_G, d_sg = (to_device(x, torch.float64, dev) for x in (CM, G, sg))
        d_np = None if npar is None else to_device(npar, torch.float64, dev)
        d_st = _u64(np.arange(B) if streams is None else streams, dev)
        d_Y = torch.zeros((B, max(nrec, 1), 3, N), dtype=torch.float64, device=dev) if want_Y and nrec > 0 else None   # rows beyond ceil(n3/downsamp) stay 0 (np.zeros, netwWilsonCowanPlastic.py:122)
        d_fin = torch.empty((B, 3, N), dtype=torch.float64, device=dev)
        check(lib.nrem_wc_run_f64_ex(C.byref(p), _ptr(d_CM), _ptr(d_G), _ptr(d_sg), _ptr(d_np), _ptr(d_st), _ptr(d_noise), nb, B,
                                     nrec, _ptr(d_Y), _ptr(d_fin), _stream()))
        Y = d_Y.cpu().numpy() if d_Y is not None else None
        return Y, d_fin.cpu().numpy()


def wc_derivative(p, CM, X, G, sigmaE, noise=None, tau_ip=2.0, device=None):
    """wilsonCowan(t, X, sigmaE, mu, tau_ip, G) of netwWilsonCowanPlastic.py:77-83 -> dX [3, N]."""
    dev = _device(device)
    N = p.nnodes
    X = np.asarray(X, dtype=np.float64)
    if X.shape != (3, N):
        raise ValueError(f"X must be (3, {N})")
    with torch.cuda.device(dev):
        d = [to_device(a, torch.float64, dev) for a in (np.asarray(CM, dtype=np.float64), X, _per_node(G, 1, N, "G")[0],
                                                        _per_node(sigmaE, 1, N, "sigmaE")[0])]
        d_nz = to_device(np.asarray(noise, dtype=np.float64), torch.float64, dev) if noise is not None else None
        out = torch.empty((3, N), dtype=torch.float64, device=dev)
        check(lib.nrem_wc_derivative_f64(C.byref(p), _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _ptr(d[3]), _ptr(d_nz),
                                         float(tau_ip), _ptr(out), _stream()))
        return out.cpu().numpy()


def _as_btn(x, name):
    x = np.asarray(x, dtype=np.float64)
    squeeze = x.ndim == 2
    if squeeze:
        x = x[None]
    if x.ndim != 3:
        raise ValueError(f"{name} must be [T, N] or [B, T, N]")
    return np.ascontiguousarray(x), squeeze


def bold_sim(rE, dt, device=None):
    """BOLDModel.Sim(rE, nnodes, dt) (call site netwWilsonCowanPlastic.py:144)."""
    dev = _device(device)
    x, sq = _as_btn(rE, "rE")
    B, T, N = x.shape
    with torch.cuda.device(dev):
        d_in = to_device(x, torch.float64, dev)
        d_out = torch.empty_like(d_in)
        check(lib.nrem_bold_sim_f64(_ptr(d_in), B, T, N, float(dt), _ptr(d_out), _stream()))
        out = d_out.cpu().numpy()
    return out[0] if sq else out


def filtfilt_decimate(bold, b, a, Neq=2000, ds=1000, device=None):
    """Cut + zero-phase band-pass + decimate of simBOLD (netwWilsonCowanPlastic.py:145-156)."""
    dev = _device(device)
    x, sq = _as_btn(bold, "bold")
    B, T, N = x.shape
    if T - Neq < 32:
        raise ValueError("need at least 32 samples after the Neq cut")
    J = (T - Neq + ds - 1) // ds
    hb = (C.c_double * 5)(*[float(v) for v in b])
    ha = (C.c_double * 5)(*[float(v) for v in a])
    with torch.cuda.device(dev):
        d_in = to_device(x, torch.float64, dev)
        nbytes = lib.nrem_filt_scratch_bytes(B, T, N, Neq, ds)
        scratch = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=dev)
        d_out = torch.empty((B, J, N), dtype=torch.float64, device=dev)
        check(lib.nrem_filtfilt_decimate_f64(_ptr(d_in), B, T, N, Neq, ds, hb, ha, _ptr(d_out), _ptr(scratch), _stream()))
        out = d_out.cpu().numpy()
    return out[0] if sq else out


def fc(bold, device=None):
    """np.corrcoef(BOLD.T) (whole_sweep_both.py:81) for [J, N] or [B, J, N]."""
    dev = _device(device)
    x, sq = _as_btn(bold, "bold")
    B, J, N = x.shape
    with torch.cuda.device(dev):
        d_in = to_device(x, torch.float64, dev)
        d_out = torch.empty((B, N, N), dtype=torch.float64, device=dev)
        check(lib.nrem_fc_f64(_ptr(d_in), B, J, N, _ptr(d_out), _stream()))
        out = d_out.cpu().numpy()
    return out[0] if sq else out


def gof(sFC, empFC, data_range=1.0, device=None):
    """utils.get_all_metrics (utils.py:42-50) for sFC [B,N,N] (or [N,N]) against empFC [K,N,N] (or [N,N]).

    Returns (gof [B, K, 4] = corr, euc, ssim, new_metric; mean FC [B])."""
    dev = _device(device)
    s = np.asarray(sFC, dtype=np.float64)
    e = np.asarray(empFC, dtype=np.float64)
    if s.ndim == 2:
        s = s[None]
    if e.ndim == 2:
        e = e[None]
    B, N, _ = s.shape
    K = e.shape[0]
    if s.shape[1:] != (N, N) or e.shape[1:] != (N, N):
        raise ValueError("sFC and empFC must be square matrices of the same size")
    with torch.cuda.device(dev):
        d_s, d_e = to_device(s, torch.float64, dev), to_device(e, torch.float64, dev)
        d_g = torch.empty((B, K, 4), dtype=torch.float64, device=dev)
        d_m = torch.empty((B,), dtype=torch.float64, device=dev)
        check(lib.nrem_gof_f64(_ptr(d_s), _ptr(d_e), B, K, N, float(data_range), _ptr(d_g), _ptr(d_m), _stream()))
        return d_g.cpu().numpy(), d_m.cpu().numpy()


def kuramoto(bold, device=None):
    """utils.kuramoto (utils.py:34-40) for [J, N] or [B, J, N] -> (sync, meta) per simulation."""
    dev = _device(device)
    x, sq = _as_btn(bold, "bold")
    B, J, N = x.shape
    with torch.cuda.device(dev):
        d_in = to_device(x, torch.float64, dev)
        d_out = torch.empty((B, 2), dtype=torch.float64, device=dev)
        d_g = torch.empty((J,), dtype=torch.float64, device=dev)
        check(lib.nrem_kuramoto_f64(_ptr(d_in), B, J, N, _ptr(d_out), _ptr(d_g), _stream()))
        out = d_out.cpu().numpy()
    return (float(out[0, 0]), float(out[0, 1])) if sq else (out[:, 0], out[:, 1])


KERNELS = {"auto": 0, "fma": 1, "tc": 2, "tc3": 3, "tcb": 4, "node32": 5, "node16": 6, "bf3": 7}
KERNEL_NAMES = {v: k for k, v in KERNELS.items()}


def integrate_f32(p, CM, G0, dG, sigma0, dsigma, mapG=None, mapS=None, map_id=None, streams=None, kernel="fma",
                  record=True, node_params=None, device=None):
    """Test hook: the sweep's float32 integrator alone (kernel: a name of KERNELS; node_params: per-node vectors, node-lane
    kernels only).  Returns (E samples [nrec, N, B] or None, final [3, N, B])."""
    dev = _device(device)
    N = p.nnodes
    G0 = np.atleast_1d(np.asarray(G0, dtype=np.float64))
    B = G0.shape[0]
    mapG = np.ones((1, N)) if mapG is None else np.atleast_2d(np.asarray(mapG, dtype=np.float64))
    mapS = np.ones((1, N)) if mapS is None else np.atleast_2d(np.asarray(mapS, dtype=np.float64))
    Bs = (B + 127) // 128 * 128
    nrec = (p.n3 + p.downsamp - 1) // p.downsamp
    mid = None if map_id is None else np.ascontiguousarray(map_id, dtype=np.int32)
    if mapG.shape[1] != N or mapS.shape != mapG.shape:
        raise ValueError(f"maps must be [n_maps, {N}]")
    if mid is not None and (mid.shape != (B,) or mid.min() < 0 or mid.max() >= mapG.shape[0]):
        raise ValueError(f"map_id must have shape ({B},) with values in [0, {mapG.shape[0]})")
    streams = np.arange(B, dtype=np.uint64) if streams is None else np.asarray(streams, dtype=np.uint64)
    if streams.shape != (B,):
        raise ValueError(f"streams must have shape ({B},), got {streams.shape}")
    with torch.cuda.device(dev):
        d = [to_device(np.broadcast_to(np.asarray(a, dtype=np.float64), (B,)).copy(), torch.float64, dev)
             for a in (G0, dG, sigma0, dsigma)]
        d_CM = to_device(np.asarray(CM, dtype=np.float64), torch.float64, dev)
        d_mG, d_mS = to_device(mapG, torch.float64, dev), to_device(mapS, torch.float64, dev)
        d_st = _u64(streams, dev)
        d_E = torch.empty((max(nrec, 1), N, Bs), dtype=torch.float32, device=dev) if record and nrec > 0 else None
        d_fin = torch.empty((3, N, Bs), dtype=torch.float32, device=dev)
        d_np = None if not node_params else to_device(node_param_table(p, node_params), torch.float64, dev)
        check(lib.nrem_sweep_integrate_f32_ex(C.byref(p), KERNELS[kernel], _ptr(d_CM), _ptr(d_mG), _ptr(d_mS), _ptr(d[0]),
                                              _ptr(d[1]), _ptr(d[2]), _ptr(d[3]),
                                              None if mid is None else mid.ctypes.data_as(C.POINTER(C.c_int32)),
                                              _ptr(d_st), _ptr(d_np), B, mapG.shape[0], nrec, _ptr(d_E), _ptr(d_fin), _stream()))
        E = d_E[:, :, :B].cpu().numpy() if d_E is not None else None
        return E, d_fin[:, :, :B].cpu().numpy()


def big_integrate_f32(p, CM, G0, dG, sigma0, dsigma, mapG=None, mapS=None, streams=None, kernel="auto", record=True,
                      want_coupling=False, node_params=None, device=None):
    """Large-connectome integrator (BASELINE configs[4]; csrc/wc_big.cuh): any 16 <= nnodes <= 8192, one launch per Euler
    step.  Returns (E samples [nrec, N, B] or None, final [3, N, B]) and, with want_coupling, SC.E of the first step [N, B].
    node_params: {name: length-N vector} for any of NODE_PARAMS (netwWilsonCowanPlastic.py:21; kernel "auto" / "bf3").
    `ops.last_integrate_ms()` gives the device time of the step launches."""
    dev = _device(device)
    N = p.nnodes
    G0 = np.atleast_1d(np.asarray(G0, dtype=np.float64))
    B = G0.shape[0]
    Bs = (B + 127) // 128 * 128
    nrec = (p.n3 + p.downsamp - 1) // p.downsamp
    with torch.cuda.device(dev):
        d = [to_device(np.broadcast_to(np.asarray(a, dtype=np.float64), (B,)).copy(), torch.float64, dev)
             for a in (G0, dG, sigma0, dsigma)]
        d_CM = to_device(np.asarray(CM, dtype=np.float64), torch.float64, dev)
        if tuple(d_CM.shape) != (N, N):
            raise ValueError(f"CM must be ({N}, {N})")
        d_mG = None if mapG is None else to_device(np.asarray(mapG, dtype=np.float64).reshape(N), torch.float64, dev)
        d_mS = None if mapS is None else to_device(np.asarray(mapS, dtype=np.float64).reshape(N), torch.float64, dev)
        d_st = _u64(np.arange(B) if streams is None else streams, dev)
        d_E = torch.empty((max(nrec, 1), N, Bs), dtype=torch.float32, device=dev) if record and nrec > 0 else None
        d_fin = torch.empty((3, N, Bs), dtype=torch.float32, device=dev)
        d_cp = torch.empty((N, Bs), dtype=torch.float32, device=dev) if want_coupling else None
        d_np = to_device(node_param_table(p, node_params), torch.float64, dev) if node_params else None
        check(lib.nrem_big_integrate_f32_ex(C.byref(p), KERNELS[kernel], _ptr(d_CM), _ptr(d_mG), _ptr(d_mS), _ptr(d[0]), _ptr(d[1]),
                                            _ptr(d[2]), _ptr(d[3]), _ptr(d_st), _ptr(d_np), B, nrec, _ptr(d_E), _ptr(d_fin), _ptr(d_cp),
                                            _stream()))
        E = d_E[:, :, :B].cpu().numpy() if d_E is not None else None
        fin = d_fin[:, :, :B].cpu().numpy()
        if want_coupling:
            return E, fin, d_cp[:, :B].cpu().numpy()
        return E, fin


def selftest_tc_coupling(E, SC, passes=1, lboA=0, sboA=0, lboB=0, sboB=0, idesc=0, device=None):
    """out[128, 96] = E[128, 96] @ SC[96, 96].T through the tcgen05 path of the integrator (diagnostic)."""
    dev = _device(device)
    E = np.ascontiguousarray(E, dtype=np.float32)
    SC = np.ascontiguousarray(SC, dtype=np.float32)
    if E.shape != (128, 96) or SC.shape != (96, 96):
        raise ValueError("E must be [128, 96] and SC [96, 96]")
    with torch.cuda.device(dev):
        d_E, d_S = to_device(E, torch.float32, dev), to_device(SC, torch.float32, dev)
        d_o = torch.zeros((128, 96), dtype=torch.float32, device=dev)
        check(lib.nrem_selftest_tc_coupling(_ptr(d_E), _ptr(d_S), _ptr(d_o), int(passes), lboA, sboA, lboB, sboB, idesc, _stream()))
        torch.cuda.synchronize(dev)
        return d_o.cpu().numpy()


def measure_fma_peak(device=None):
    """FP32 FMA peak of the device in TFLOP/s (register-only FMA chains; the integrator's roofline denominator)."""
    dev = _device(device)
    tf, ms = C.c_double(), C.c_double()
    with torch.cuda.device(dev):
        check(lib.nrem_measure_fma_peak(C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def last_integrate_ms():
    return float(lib.nrem_last_integrate_ms())


def launch_count(reset=False):
    return int(lib.nrem_launch_count(1 if reset else 0))
