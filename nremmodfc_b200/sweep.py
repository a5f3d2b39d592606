"""Batched G x sigma x seed x map sweeps — the fast path.

One call replaces the body of the reference's driver loops (whole_sweep_both.py:63-96,
whole_sweep_both_maps.py:99-132, run_many_seeds.py:105-136): for every simulation b
    G_i     = G0[b]     + dG[b]     * mapG[map_id[b]][i]        (whole_sweep_both_maps.py:104-105)
    sigma_i = sigma0[b] + dsigma[b] * mapS[map_id[b]][i]        (whole_sweep_both_maps.py:107-108)
it integrates the Wilson-Cowan network, runs BOLD -> cut -> band-pass -> decimate -> FC and
returns the goodness of fit against K empirical matrices, all on the GPU.

Multi-GPU: simulations are independent (the reference shards them with ``sim % threads == rank``,
whole_sweep_both.py:64), so ``shard_ids`` gives each rank its slice and ``gather_rows`` collects
the small result table once at the end — no per-step collective exists on this path.
"""
import ctypes as C
import itertools

import numpy as np
import torch
from scipy import signal

from . import ops
from ._lib import SweepOpts, check, lib

TILE = 128


def bandpass_ba(bold_dt):
    """The reference's filter design, netwWilsonCowanPlastic.py:152 (returns SciPy's (b, a))."""
    return signal.bessel(2, [2 * 0.01 * bold_dt, 2 * 0.1 * bold_dt], btype="bandpass")


def product_grid(seeds, delta_G_vals, delta_sigma_vals):
    """Flattened parameter product in the reference's order (whole_sweep_both.py:60-61).

    Returns (seed[B], delta_G[B], delta_sigma[B])."""
    sims = list(itertools.product(seeds, delta_G_vals, delta_sigma_vals))
    a = np.asarray(sims, dtype=np.float64).reshape(-1, 3)
    return a[:, 0].astype(np.int64), a[:, 1].copy(), a[:, 2].copy()


def shard_ids(n, rank, world, contiguous=False):
    """Simulation ids of one rank.  Default = the reference's round-robin (whole_sweep_both.py:64)."""
    ids = np.arange(n)
    if contiguous:
        return np.array_split(ids, world)[rank]
    return ids[ids % world == rank]


def gather_rows(local_ids, local_rows, n_total, group=None):
    """All-gather per-rank result rows into the full [n_total, C] table (every rank gets it).

    Works on any torch.distributed backend (NCCL on GPUs, gloo in CPU tests); without an
    initialised process group it just scatters the local rows."""
    import torch.distributed as dist
    local_rows = np.asarray(local_rows, dtype=np.float64)
    ncol = local_rows.shape[1]
    out = np.full((n_total, ncol), np.nan)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out[np.asarray(local_ids)] = local_rows
        return out
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    cap = (n_total + world - 1) // world
    buf = torch.full((cap, ncol + 1), -1.0, dtype=torch.float64, device=dev)
    k = len(local_ids)
    if k:
        buf[:k, 0] = torch.as_tensor(np.asarray(local_ids, dtype=np.float64), device=dev)
        buf[:k, 1:] = torch.as_tensor(local_rows, device=dev)
    allb = torch.empty((world, cap, ncol + 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allb.view(-1, ncol + 1), buf, group=group)
    allb = allb.view(-1, ncol + 1).cpu().numpy()
    valid = allb[:, 0] >= 0
    out[allb[valid, 0].astype(np.int64)] = allb[valid, 1:]
    return out


def shuffled_symmetric_maps(vec, K, seed):
    """K hemisphere-mirrored shuffles of a per-node map, the rule of empirical/retrieve_AALmaps.py:62-75 (BASELINE configs[3],
    "many shuffled maps"): the first half of the nodes is permuted at random and node N-1-i receives the value of node
    N-1-perm[i], so homotopic regions stay paired.  Returns (maps [K, N], index [K, N]) from numpy.random.default_rng(seed)."""
    vec = np.asarray(vec, dtype=np.float64)
    N = vec.shape[0]
    if N % 2:
        raise ValueError("a hemisphere-mirrored shuffle needs an even number of nodes")
    h = N // 2
    rng = np.random.default_rng(seed)
    idx = np.empty((K, N), dtype=np.int64)
    for k in range(K):
        perm = rng.permutation(h)
        idx[k, :h] = perm
        idx[k, N - 1 - np.arange(h)] = N - 1 - perm
    return vec[idx], idx


def pad_by_map(map_id):
    """Order/pad simulations so that every 128-tile holds one map id.

    Returns (src[Bp], valid[Bp]): src[k] = index of the simulation placed at padded slot k."""
    map_id = np.asarray(map_id)
    src, valid = [], []
    for m in np.unique(map_id):
        idx = np.nonzero(map_id == m)[0]
        padn = (-len(idx)) % TILE
        src.append(np.concatenate([idx, np.full(padn, idx[-1])]))
        valid.append(np.concatenate([np.ones(len(idx), bool), np.zeros(padn, bool)]))
    return np.concatenate(src), np.concatenate(valid)


class SweepPlan:
    """Owns the device scratch of one batch size; reusable across calls of ``run``."""

    def __init__(self, p, B, n_maps=1, K=4, kernel="auto", bold_f32=True, chunk_samples=0, Neq=2000, bold_downsamp=1000,
                 bold_dt=None, dt=0.002, peakfreq=False, welch_nperseg=4000, device=None):
        self.dev = ops._device(device)
        self.p, self.B, self.n_maps, self.K, self.N = p, int(B), int(n_maps), int(K), p.nnodes
        if bold_dt is None:
            bold_dt = dt * p.downsamp                     # netwWilsonCowanPlastic.py:144 (module dt * downsamp)
        b, a = bandpass_ba(bold_dt)
        o = SweepOpts()
        o.kernel, o.bold_f32, o.chunk_samples, o.want_fc = ops.KERNELS[kernel], int(bool(bold_f32)), int(chunk_samples), 0
        o.Neq, o.bold_downsamp, o.bold_dt = int(Neq), int(bold_downsamp), float(bold_dt)
        # peakfreq: Welch spectrum of the stored E samples (whole_sweep_both.py:90-95: fs = 1/dt, nperseg = 4000)
        o.welch_nperseg, o.welch_fs = (int(welch_nperseg) if peakfreq else 0), 1.0 / (p.dtSim * p.downsamp)
        o.b = (C.c_double * 5)(*b)
        o.a = (C.c_double * 5)(*a)
        self.opts = o
        self._plan = C.c_void_p()
        with torch.cuda.device(self.dev):
            check(lib.nrem_sweep_create(C.byref(p), C.byref(o), self.B, self.n_maps, self.K, C.byref(self._plan)))
        self.device_bytes = int(lib.nrem_sweep_device_bytes(self._plan))
        T = (p.n3 + p.downsamp - 1) // p.downsamp
        self.T, self.J = T, (T - int(Neq) + int(bold_downsamp) - 1) // int(bold_downsamp)
        self._held = None

    def set_profiling(self, on=True):
        check(lib.nrem_sweep_set_profiling(self._plan, 1 if on else 0))

    def profile(self):
        """Device times since the previous call (blocks until the enqueued work has finished):
        dict(total_ms, integrator_ms, integrator_launches, tile_groups)."""
        out = (C.c_double * 4)()
        check(lib.nrem_sweep_get_profile(self._plan, out))
        return {"total_ms": out[0], "integrator_ms": out[1], "integrator_launches": int(out[2]), "tile_groups": int(out[3])}

    def close(self):
        if self._plan:
            lib.nrem_sweep_destroy(self._plan)
            self._plan = C.c_void_p()
        self._held = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_name(self):
        """The integrator kernel the plan resolved to: "tc3" (128 simulations per CTA), "node32" / "node16" (node-lane kernel,
        32 / 16 simulations per CTA: small batches, nnodes > 96, per-node parameter tables), "fma", "tc"."""
        return ops.KERNEL_NAMES[int(lib.nrem_sweep_kernel(self._plan))]

    def set_node_params(self, node_params=None):
        """Per-node vectors for any of ops.NODE_PARAMS ("Any of them can be redefined as a vector of length nnodes",
        netwWilsonCowanPlastic.py:21) for the next runs: {name: length-N vector}; names not given keep the scalar of the plan's
        parameters.  None / {} goes back to scalars."""
        if not node_params:
            check(lib.nrem_sweep_set_node_params(self._plan, None, ops._stream()))
            return
        table = ops.node_param_table(self.p, node_params)
        with torch.cuda.device(self.dev):
            d = ops.to_device(table, torch.float64, self.dev)
            check(lib.nrem_sweep_set_node_params(self._plan, ops._ptr(d), ops._stream()))
            torch.cuda.current_stream().synchronize()          # d is freed on return

    @property
    def chunks_total(self):
        """Integrator launches (per tile group) of one whole run: the unit `advance` counts in."""
        return int(lib.nrem_sweep_chunks_total(self._plan))

    # -- device-resident calls (bench `value`): tensors already in HBM ----------------------------
    def begin_device(self, d_CM, d_mapG, d_mapS, d_G0, d_dG, d_s0, d_ds, map_id, d_streams, homogeneous=-1):
        """Stage the inputs of a run and rewind the plan (nrem_sweep_begin).  homogeneous: True/False when the caller knows
        whether every map entry is exactly 1, -1 to let the library find out (one stream synchronisation)."""
        mid = None
        if map_id is not None:
            mid = np.ascontiguousarray(map_id, dtype=np.int32)
            if mid.shape != (self.B,):
                raise ValueError(f"map_id must have shape ({self.B},), got {mid.shape}")
            if mid.min() < 0 or mid.max() >= self.n_maps:
                raise ValueError(f"map_id values must lie in [0, {self.n_maps})")
        for name, t, n in (("G0", d_G0, self.B), ("dG", d_dG, self.B), ("sigma0", d_s0, self.B), ("dsigma", d_ds, self.B),
                           ("streams", d_streams, self.B), ("CM", d_CM, self.N * self.N), ("mapG", d_mapG, self.n_maps * self.N),
                           ("mapS", d_mapS, self.n_maps * self.N)):
            if t.numel() != n:
                raise ValueError(f"{name} has {t.numel()} elements, expected {n}")
        hint = -1 if homogeneous in (-1, None) else int(bool(homogeneous))
        self._held = (d_CM, d_mapG, d_mapS, d_G0, d_dG, d_s0, d_ds, d_streams)       # alive until the kernels have read them
        with torch.cuda.device(self.dev):
            check(lib.nrem_sweep_begin(self._plan, ops._ptr(d_CM), ops._ptr(d_mapG), ops._ptr(d_mapS), ops._ptr(d_G0),
                                       ops._ptr(d_dG), ops._ptr(d_s0), ops._ptr(d_ds),
                                       None if mid is None else mid.ctypes.data_as(C.POINTER(C.c_int32)),
                                       ops._ptr(d_streams), hint, ops._stream()))

    def advance(self, max_chunks=None):
        """Enqueue up to max_chunks integrator launches (default: all that are left); returns how many are left."""
        left = C.c_int64()
        n = self.chunks_total if max_chunks is None else int(max_chunks)
        with torch.cuda.device(self.dev):
            check(lib.nrem_sweep_advance(self._plan, n, C.byref(left), ops._stream()))
        return int(left.value)

    def feed_samples(self, E):
        """Test hook: feed stored E samples [rows, N, B] (float32; what run() records) to the plan's BOLD / filter / spectrum
        kernels instead of integrating.  Consecutive calls continue the series; all but the last must bring a multiple of
        chunk_samples rows."""
        E = np.ascontiguousarray(E, dtype=np.float32)
        if E.ndim != 3 or E.shape[1:] != (self.N, self.B):
            raise ValueError(f"E must be [rows, {self.N}, {self.B}], got {E.shape}")
        Bs = (self.B + TILE - 1) // TILE * TILE
        with torch.cuda.device(self.dev):
            d = torch.zeros((E.shape[0], self.N, Bs), dtype=torch.float32, device=self.dev)
            d[:, :, :self.B] = ops.to_device(E, torch.float32, self.dev)
            check(lib.nrem_sweep_feed_samples(self._plan, ops._ptr(d), E.shape[0], ops._stream()))
            torch.cuda.current_stream().synchronize()          # d is freed on return

    def finish_device(self, d_emp, d_gof, d_extra, d_fc=None):
        if d_emp.numel() != self.K * self.N * self.N or d_gof.numel() != self.B * self.K * 4:
            raise ValueError("emp / gof have the wrong size")
        with torch.cuda.device(self.dev):
            check(lib.nrem_sweep_finish(self._plan, ops._ptr(d_emp), ops._ptr(d_gof), ops._ptr(d_extra), ops._ptr(d_fc), ops._stream()))

    def run_device(self, d_CM, d_mapG, d_mapS, d_G0, d_dG, d_s0, d_ds, map_id, d_streams, d_emp, d_gof, d_extra, d_fc=None,
                   homogeneous=-1):
        self.begin_device(d_CM, d_mapG, d_mapS, d_G0, d_dG, d_s0, d_ds, map_id, d_streams, homogeneous)
        self.advance()
        self.finish_device(d_emp, d_gof, d_extra, d_fc)

    # -- host calls (bench `e2e`, drivers): NumPy in, NumPy out -----------------------------------
    def begin(self, CM, G0, dG, sigma0, dsigma, streams, mapG=None, mapS=None, map_id=None):
        """Upload the inputs of one run (pinned H2D) and rewind the plan to Euler step 0."""
        B, N, dev = self.B, self.N, self.dev
        f64 = torch.float64
        mapG = np.ones((1, N)) if mapG is None else np.atleast_2d(np.asarray(mapG, dtype=np.float64))
        mapS = np.ones((1, N)) if mapS is None else np.atleast_2d(np.asarray(mapS, dtype=np.float64))
        if mapG.shape != (self.n_maps, N) or mapS.shape != (self.n_maps, N):
            raise ValueError(f"maps must be [{self.n_maps}, {N}]")
        CM = np.asarray(CM, dtype=np.float64)
        if CM.shape != (N, N):
            raise ValueError(f"CM must be [{N}, {N}], got {CM.shape}")
        streams = np.asarray(streams, dtype=np.uint64)
        if streams.shape != (B,):
            raise ValueError(f"streams must have shape ({B},), got {streams.shape}")
        per_sim = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (B,))) for a in (G0, dG, sigma0, dsigma)]
        homo = bool(np.all(mapG == 1.0) and np.all(mapS == 1.0))
        with torch.cuda.device(dev):
            d_CM = ops.to_device(CM, f64, dev)
            d_mG, d_mS = ops.to_device(mapG, f64, dev), ops.to_device(mapS, f64, dev)
            d_par = [ops.to_device(a, f64, dev) for a in per_sim]
            d_st = ops._u64(streams, dev)
            self.begin_device(d_CM, d_mG, d_mS, d_par[0], d_par[1], d_par[2], d_par[3], map_id, d_st, homogeneous=homo)
        self.h2d_bytes = sum(int(t.numel() * t.element_size()) for t in [d_CM, d_mG, d_mS, d_st] + d_par)

    def finish(self, emp, want_fc=False):
        """Backward filter pass, FC, GoF and observables of the completed run; downloads the result table (D2H)."""
        B, N, K, dev = self.B, self.N, self.K, self.dev
        f64 = torch.float64
        emp = np.asarray(emp, dtype=np.float64)
        if emp.shape != (K, N, N):
            raise ValueError(f"emp must be [{K}, {N}, {N}]")
        with torch.cuda.device(dev):
            d_emp = ops.to_device(emp, f64, dev)
            d_gof = torch.empty((B, K, 4), dtype=f64, device=dev)
            d_extra = torch.empty((B, 4), dtype=f64, device=dev)
            d_fc = torch.empty((B, N, N), dtype=f64, device=dev) if want_fc else None
            self.finish_device(d_emp, d_gof, d_extra, d_fc)
            extra = d_extra.cpu().numpy()
            out = {"gof": d_gof.cpu().numpy(), "mean": extra[:, 0], "sync": extra[:, 1], "meta": extra[:, 2], "peakfreq": extra[:, 3]}
            if want_fc:
                out["fc"] = d_fc.cpu().numpy()
        self._held = None
        self.h2d_bytes = getattr(self, "h2d_bytes", 0) + int(d_emp.numel() * 8)
        self.d2h_bytes = int(d_gof.numel() * 8 + B * 4 * 8 + (d_fc.numel() * 8 if want_fc else 0))
        return out

    def run(self, CM, emp, G0, dG, sigma0, dsigma, streams, mapG=None, mapS=None, map_id=None, want_fc=False, node_params=None):
        if node_params is not None:
            self.set_node_params(node_params)
        self.begin(CM, G0, dG, sigma0, dsigma, streams, mapG, mapS, map_id)
        self.advance()
        return self.finish(emp, want_fc)


def sweep_gof(p, CM, emp, G0, dG, sigma0, dsigma, streams, mapG=None, mapS=None, map_id=None, want_fc=False, node_params=None, **plan_kw):
    """One-shot convenience wrapper: build a plan for len(streams) simulations, run it, free it."""
    B = len(np.atleast_1d(streams))
    n_maps = 1 if mapG is None else np.atleast_2d(mapG).shape[0]
    plan = SweepPlan(p, B, n_maps=n_maps, K=np.asarray(emp).shape[0], **plan_kw)
    try:
        return plan.run(CM, emp, G0, dG, sigma0, dsigma, streams, mapG, mapS, map_id, want_fc, node_params)
    finally:
        plan.close()
