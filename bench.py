#!/usr/bin/env python
"""Benchmark of the hot path: WC + BOLD + FC + GoF simulations per second on the reference's
homogeneous G x sigma x 50-seed sweep (BASELINE.json configs[1]; whole_sweep_both.py grid).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]

Workload = the 20 000-simulation sweep (50 seeds x 20 dG x 20 dsigma, AAL90, full length 1 + 400 + 600 s = 10.01 M
Euler steps per simulation) through the public host API (`nremmodfc_b200.sweep.SweepPlan`: NumPy in, NumPy out).
The time loop of a simulation is sequential, so the sweep is integrated ONCE and cut into SLICES = 20 equal time
slices with the state carried in the plan; one bench "step" = one slice = every simulation of the sweep advanced by
1/20 of its horizon (100 integrator launches of 5000 Euler steps per tile group + the BOLD/filter launches of the
recording phase).  Slice 0 of a sweep uploads the inputs (pinned H2D), slice 19 also runs the backward filter / FC /
GoF tail, downloads the result table (D2H) and, on several GPUs, gathers it.  `--steps 20` is therefore exactly one
whole sweep; other K cover K/20 sweeps.  Inside every timed step
  * `e2e`   = wall clock of the API calls (H2D of SC/targets/parameters, D2H of the GoF table included),
  * `value` = the same slices timed on the device with CUDA events (inputs resident in HBM -> GoF in HBM).
Warm-up: one whole-pipeline pass on 1 % of the horizon (loads every kernel), then W slices of the real sweep.
Multi-GPU (torchrun, one rank per GPU): weak scaling — every rank runs its own 20 000-simulation sweep
(seeds 50*rank ... 50*rank+49), no data-path collective, one final all-gather of the GoF table.  `--scaling strong`
instead shards ONE 60 000-simulation job (the paper's three modalities: homogeneous, NA/ACh maps, shuffled maps)
over the ranks with the reference's rule sim % world == rank.
`--impl reference` times the reference's CPU path on the host cores (the unmodified numba module from baseline/_ref
when it imports, else the oracle's C port), one process per core — the reference's own parallel model
(whole_sweep_both.py:23-24) — on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL = dict(n1=10_000, n2=4_000_000, n3=6_000_000)           # whole_sweep_both.py:43-50
STEPS_PER_SIM = sum(FULL.values())
FLOP_PER_STEP = 2 * 90 * 90 + 39 * 90                        # dense dgemv + elementwise
NODE_SECONDS_PER_SIM = 1001 * 90
METRIC = "WC+BOLD sims/sec (AAL90, 50-seed G x sigma sweep)"
SLICES = 20                                                  # bench steps per whole sweep
DATA = "reference inputs (AAL90 SC_opti_25julio + empirical mean_mat_{W,N1,N2,N3} + NA/ACh maps, data/aal90_inputs.npz); " \
       "parameter grid and noise generated (no dataset is read at run time)"


def load_inputs():
    d = np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"))
    return d, d["SC"], np.stack([d[s] for s in ("W", "N1", "N2", "N3")])


def sweep_grid(rank, n_sims):
    """whole_sweep_both.py:57-61 with the committed table's grid (whole_sweep_both_maps.py:92-93)."""
    from nremmodfc_b200 import sweep
    seeds = np.arange(50) + 50 * rank
    dG = np.linspace(-0.1, 0.3, 20, endpoint=False)
    dS = np.linspace(-0.2, 0.2, 20, endpoint=False)
    s, g, sg = sweep.product_grid(seeds, dG, dS)
    return s[:n_sims], g[:n_sims], sg[:n_sims]


def chunk_schedule(p, chunk_samples):
    """Euler steps of every integrator launch of a run, in order (mirror of csrc/nremfc_api.cu:integrate)."""
    cs = chunk_samples * p.downsamp
    out = []
    for n in (p.n1, p.n2, p.n3):
        out += [min(cs, n - i) for i in range(0, n, cs)]
    return out


# ---- CPU baselines ---------------------------------------------------------------------------------
def _port_worker(args):
    seed, frac = args
    from oracle import bold_oracle, cwrap, wc_oracle
    _, SC, emp = load_inputs()
    p = wc_oracle.params(P=0.4, rhoE=0.18)
    n1, n2, n3 = int(FULL["n1"] * frac), int(FULL["n2"] * frac), int(FULL["n3"] * frac)
    E = cwrap.wc_run(SC, 0.16, 7.68, n1, n2, n3, seed=1, stream=seed, p=p, want="E", fast_rng=True)
    bold = cwrap.bold_sim(E, 0.04)
    neq = min(2000, max(0, E.shape[0] - 1100))
    y = bold_oracle.filt_decimate(bold, 1000 if E.shape[0] > 12000 else 100, neq, 0.04)
    FC = bold_oracle.fc(y)
    g = [bold_oracle.get_all_metrics(FC, emp[k]) for k in range(4)]
    return float(g[0][0])


_REF = {}


def _ref_init():
    """Worker start-up for the numba arm: import the UNMODIFIED reference module from baseline/_ref (with the oracle's shims
    for its two absent dependencies, BOLDModel and skimage) and JIT-compile it once on a tiny horizon."""
    from oracle import refshim
    wc, utils = refshim.import_reference()
    d, SC, emp = load_inputs()
    wc.P, wc.rhoE, wc.CM = 0.4, 0.18, SC                      # whole_sweep_both.py:39-41
    wc.G, wc.sigmaE = 0.16, 7.68
    _REF.update(wc=wc, utils=utils, emp=emp)
    _ref_sim(0.0002)                                          # first call: numba compiles run()/wilsonCowan()


def _ref_sim(frac):
    """One simulation of the reference's driver loop body (whole_sweep_both.py:66-96) on `frac` of the horizon.
    Returns (recompile seconds, run seconds, simBOLD + FC + GoF seconds)."""
    wc, utils, emp = _REF["wc"], _REF["utils"], _REF["emp"]
    t1, t2, ts = 1.0 * frac, 400.0 * frac, 600.0 * frac
    wc.tTrans1, wc.tTrans2, wc.tstop = t1, t2, ts
    wc.timeTrans1, wc.timeTrans2 = np.arange(0, t1, wc.dtSim), np.arange(0, t2, wc.dtSim)
    wc.timeSim, wc.time = np.arange(0, ts, wc.dtSim), np.arange(0, ts, wc.dt)
    a = time.perf_counter()
    wc.run.recompile()                                        # whole_sweep_both.py:75 (every simulation)
    b = time.perf_counter()
    tray = wc.run()
    c = time.perf_counter()
    E_t = tray[:, 0, :]
    if E_t.shape[0] > 3200:
        BOLD = wc.simBOLD(E_t, nnodes=90, BOLD_downsamp=1000 if E_t.shape[0] > 12000 else 100)
        sFC = np.corrcoef(BOLD.T)
        for k in range(4):
            utils.get_all_metrics(sFC, emp[k], data_range=1)
    return b - a, c - b, time.perf_counter() - c


def _ref_worker(frac):
    return _ref_sim(frac)


def cpu_baseline(frac=0.05, kind="auto", pool_holder=None):
    """Times the reference's pipeline on all host cores, one process per core; returns sims/s of FULL-length
    simulations (the cost is linear in the number of Euler steps).  kind: "reference" = the unmodified numba module
    from baseline/_ref, "port" = the oracle's C restatement, "auto" = reference when it imports, else port."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    if kind in ("auto", "reference"):
        try:
            from oracle import refshim
            refshim.check_available()
            kind = "reference"
        except Exception as e:  # noqa: BLE001
            if kind == "reference":
                raise
            kind, why = "port", f"{type(e).__name__}: {e}"
    if kind == "reference":
        pool = pool_holder.get("pool") if pool_holder is not None else None
        if pool is None:
            pool = ctx.Pool(cores, initializer=_ref_init)
            pool.map(_ref_worker, [0.0002] * cores, chunksize=1)          # every worker is up and compiled
            if pool_holder is not None:
                pool_holder["pool"] = pool
        t0 = time.perf_counter()
        parts = pool.map(_ref_worker, [frac] * cores, chunksize=1)
        wall = time.perf_counter() - t0
        if pool_holder is None:
            pool.close()
        rec, run, tail = (float(np.mean([x[k] for x in parts])) for k in range(3))
        # cost model of a FULL simulation: integration and BOLD/filter scale with the horizon, the per-simulation
        # run.recompile() (whole_sweep_both.py:75) does not
        full_no_jit = (run + tail) / frac
        return {"value": cores / (full_no_jit + rec), "unit": "sims/s", "cores": cores, "kind": "reference",
                "value_without_recompile": cores / full_no_jit, "recompile_s_per_sim": rec,
                "sample": f"{cores} sims x {frac:g} of the 10.01M-step horizon through the unmodified netwWilsonCowanPlastic.run / simBOLD "
                          f"/ utils.get_all_metrics (numba, one process per core, JIT warm), {wall:.1f} s wall; integration + BOLD scaled "
                          f"linearly in steps, plus the driver's per-simulation run.recompile() ({rec:.2f} s)",
                "wall_s": wall}
    from oracle import cwrap
    cwrap.build()
    tasks = [(i, frac) for i in range(cores)]
    with ctx.Pool(cores) as pool:
        pool.map(_port_worker, [(0, 0.0005)] * cores)          # warm: page in scipy, build tables
        t0 = time.perf_counter()
        pool.map(_port_worker, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    return {"value": len(tasks) * frac / wall, "unit": "sims/s", "cores": cores, "kind": "port",
            "sample": f"{len(tasks)} sims x {frac:g} of the 10.01M-step horizon (+BOLD/filter/FC/GoF) in the oracle's C port, one process "
                      f"per core, {wall:.1f} s wall; scaled linearly in steps",
            "wall_s": wall}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, holder = [], {}
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(frac=args.cpu_frac, kind=args.cpu_kind, pool_holder=holder)
        if i >= args.warmup:
            vals.append(r)
    if holder.get("pool") is not None:
        holder["pool"].close()
    v = float(np.mean([r["value"] for r in vals]))
    wall = float(np.mean([r["wall_s"] for r in vals]))
    last = vals[-1]
    cb = {k: last[k] for k in last if k != "wall_s"} | {"value": v}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "sims/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": DATA,
            "config": {"workload": "configs[1]: homogeneous G x sigma sweep x 50 seeds (AAL90, 1+400+600 s)", "sample": last["sample"],
                       "step": "one bounded sample on all host cores"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "node_seconds_per_s": v * NODE_SECONDS_PER_SIM}
    print(json.dumps(line))


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median([float(r[0]) for r in rows])), "sm_max_mhz": float(rows[0][1]),
                "power_w_median": float(np.median(pw)) if pw else None, "samples": len(rows), "reasons": reasons}


# ---- extra legs (N = 1 only, outside the headline's timed region) -----------------------------------
def config5_leg(N=1000, B=4096, steps=40000):
    """BASELINE configs[4] (scaled synthetic connectome, 1000-node random SC, 4096 instances) on the per-step tcgen05 kernel
    (csrc/wc_big.cuh): Euler steps/s of the whole batch and the coupling rate, CUDA events around the step launches.
    Roofline: tensor (three BF16 passes of the 3xBF16 split)."""
    from nremmodfc_b200 import ops
    rng = np.random.default_rng(5)
    SC = rng.uniform(size=(N, N))                   # netwWilsonCowanPlastic.py:64 placeholder distribution
    np.fill_diagonal(SC, 0.0)
    SC *= 2.5 / SC.sum(axis=1).mean()               # mean row sum of AAL90
    dGv, dSv = np.linspace(-0.1, 0.3, 20, endpoint=False), np.linspace(-0.2, 0.2, 20, endpoint=False)
    dG, dS = dGv[rng.integers(0, 20, B)], dSv[rng.integers(0, 20, B)]
    pw = ops.make_params(N, 200, 1800, 0, P=0.4, rhoE=0.18, seed=1)
    ops.big_integrate_f32(pw, SC, np.full(B, 0.16), dG, np.full(B, 7.68), dS, kernel="auto", record=False)      # warm-up
    p = ops.make_params(N, steps // 10, steps - steps // 10, 0, P=0.4, rhoE=0.18, seed=1)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    _, fin = ops.big_integrate_f32(p, SC, np.full(B, 0.16), dG, np.full(B, 7.68), dS, kernel="auto", record=False)
    clocks = sampler.stop()
    us = ops.last_integrate_ms() * 1e3 / steps
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            bf16, src = float(json.load(fh)["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (long run under the power cap; TF32 = half)"
    except (OSError, KeyError, ValueError):
        bf16, src = 1400.0, "fallback sustained bf16 (B200_PROFILING.md; TF32 = half)"
    alg = 2.0 * B * N * N                            # flop per step, as the reference's dgemv
    ideal_us = 3 * alg / (bf16 * 1e12) * 1e6        # bf3: three BF16 passes (hi.hi + lo.hi + hi.lo)
    return {"workload": f"configs[4]: {N}-node random SC, {B} instances, {steps} Euler steps, kernel bf3 (3xBF16 split, CTA pairs)", "us_per_step": us,
            "steps_per_s": 1e6 / us, "node_updates_per_s": B * N / us * 1e6, "results_finite": bool(np.isfinite(fin).all()),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": alg / us / 1e6, "unit": "TFLOP/s (algorithmic 2*B*N^2 per step)",
                         "peak": alg / ideal_us / 1e6, "frac": ideal_us / us, "traffic": None,
                         "peak_note": "three BF16 passes at the tensor peak; " + src,
                         "executed_tflops": 3 * alg / us / 1e6,
                         "frac_round1_definition": (alg / (bf16 / 2 * 1e12) * 1e6 + 2 * alg / (bf16 * 1e12) * 1e6) / us,
                         "frac_round1_definition_note": "ideal time of the PREVIOUS kernel's pass mix (1 TF32 + 2 BF16 passes, tcb) over this run's time: "
                                                        "comparable with the 0.53 / 0.58 of round 1 / the first half of round 2; frac itself charges only the "
                                                        "three BF16 passes this kernel executes",
                         "note": "round 2: the 3xBF16 split needs 1.5 TF32-pass equivalents (tcb: 2, tc3: 3), so the step got 18 % faster while the "
                                 "executed-flop fraction stayed; with the FP32 plane of E in place and a_base as bf16 (working set 124 -> 93 MB) another 10 %.  "
                                 "The step is bound by the operand ring and phase 1 alike (a third of the state / operand reads still miss the L2), "
                                 "see profiles/r02_big_connectome.md",
                         "kernel": "wc_big_step_kernel<5, ., ., 0, 1>", "profile": "profiles/r02_big_connectome.md"}}


def k1_only_leg(d, steps=20000):
    """The integrator kernel alone on one full wave (one 128-simulation tile per SM, `steps` Euler steps in one launch chain, CUDA events):
    SM cycles per Euler step of a tile without the BOLD / filter launches and the tile-group scheduling of the sweep -- the denominator of the
    kernel's own composite-bound fraction."""
    import torch
    from nremmodfc_b200 import ops
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B = sms * 128
    rng = np.random.default_rng(0)
    dG, ds = rng.uniform(-0.1, 0.3, B), rng.uniform(-0.2, 0.2, B)
    p = ops.make_params(90, 0, steps, 0, P=0.4, rhoE=0.18, seed=1)
    ms = []
    for _ in range(3):          # first call warms up
        ops.integrate_f32(p, d["SC"], np.full(B, 0.16), dG, np.full(B, 7.68), ds, kernel="tc3", record=False)
        ms.append(ops.last_integrate_ms())
    best = min(ms[1:])
    return {"tiles": sms, "euler_steps": steps, "ms": best, "us_per_euler_step": best * 1e3 / steps,
            "sims_per_s_equivalent": B * steps / (best * 1e-3) / 1.001e7}


def config1_leg(d, SC, emp):
    """BASELINE configs[0]: ONE full-length run (AAL90, G = 0.16, sigma = 7.68, one seed) through the drop-in module surface, exactly the
    reference's call sequence (cortex_run.py:103-116 / whole_sweep_both.py:66-96): attributes, run.recompile(), run() -> Y_t float64
    [300000, 3, 90], simBOLD(), np.corrcoef, utils.get_all_metrics vs mean_mat_W.  Float64 one-CTA kernel: the time loop of a single simulation
    is sequential, so this is a latency figure (us per Euler step), not a throughput one."""
    from nremmodfc_b200 import netwWilsonCowanPlastic as wc, utils
    wc.P, wc.rhoE, wc.CM = 0.4, 0.18, SC
    wc.tTrans1, wc.tTrans2, wc.tstop = 1, 400, 600
    wc.timeTrans1, wc.timeTrans2 = np.arange(0, 1, wc.dtSim), np.arange(0, 400, wc.dtSim)
    wc.timeSim, wc.time = np.arange(0, 600, wc.dtSim), np.arange(0, 600, wc.dt)
    wc.G, wc.sigmaE, wc.sid = 0.16, 7.68, 0
    t0 = time.perf_counter()
    wc.run.recompile()
    tray = wc.run()
    t1 = time.perf_counter()
    BOLD = wc.simBOLD(tray[:, 0, :], nnodes=90)
    sFC = np.corrcoef(BOLD.T)
    m = utils.get_all_metrics(sFC, emp[0], data_range=1)
    t2 = time.perf_counter()
    return {"workload": "configs[0]: one Wilson-Cowan + BOLD run, AAL90, G=0.16, sigma=7.68, 1+400+600 s, FC + GoF vs mean_mat_W, through "
                        "netwWilsonCowanPlastic.run() / simBOLD() (float64, Y_t of 648 MB returned to the host)",
            "run_s": t1 - t0, "us_per_euler_step": (t1 - t0) / STEPS_PER_SIM * 1e6, "bold_fc_gof_s": t2 - t1, "total_s": t2 - t0,
            "corrW": float(m[0]), "eW": float(m[1]), "ssimW": float(m[2]), "results_finite": bool(np.isfinite(sFC).all())}


def modality_leg(args, d, SC, emp, which):
    """BASELINE configs[2] / configs[3]: one whole 20 000-simulation sweep with heterogeneous NA/ACh maps (whole_sweep_both_maps.py:
    104-108), full length, through the same host API; sims/s from the wall clock of the call (H2D and D2H included)."""
    import torch
    from nremmodfc_b200 import ops, sweep
    norm = lambda m: m / m.mean()                                                   # whole_sweep_both_maps.py:54,62
    seeds, dG, dS = sweep_grid(0, 20000)
    B0 = len(seeds)
    if which == "map":
        mapG, mapS = norm(d["map_ACh"])[None], norm(d["map_NA"])[None]
        mid = np.zeros(B0, np.int32)
        label = "configs[2]: map modality (DIST_VAChT -> G_i, DIST_LC_proj -> sigma_i), 50 seeds x 20 x 20"
    else:
        K = 4                                                                       # "many shuffled maps": K fresh hemisphere-mirrored shuffles
        mapG, _ = sweep.shuffled_symmetric_maps(norm(d["map_ACh"]), K, seed=1)
        mapS, _ = sweep.shuffled_symmetric_maps(norm(d["map_NA"]), K, seed=2)
        mid = (np.arange(B0) % K).astype(np.int32)
        label = f"configs[3]: shuffled-map control, {K} hemisphere-mirrored shuffles (empirical/retrieve_AALmaps.py:62-75) x 50 seeds x 20 x 20"
    src, valid = sweep.pad_by_map(mid)
    B = len(src)
    streams = (np.uint64(7 if which == "map" else 9) << np.uint64(48)) | src.astype(np.uint64)
    p = ops.make_params(90, FULL["n1"], FULL["n2"], FULL["n3"], P=0.4, rhoE=0.18, seed=2024)
    plan = sweep.SweepPlan(p, B, n_maps=mapG.shape[0], kernel=args.kernel, bold_f32=not args.bold_f64, chunk_samples=args.chunk_samples)
    plan.set_profiling(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = plan.run(SC, emp, 0.16, dG[src], 7.68, dS[src], streams, mapG=mapG, mapS=mapS, map_id=mid[src])
    wall = time.perf_counter() - t0
    dev_ms = plan.profile()["total_ms"]
    plan.close()
    return {"workload": label, "sims": int(B0), "padded_sims": int(B), "e2e_sims_per_s": B0 / wall, "device_sims_per_s": B0 / (dev_ms * 1e-3),
            "seconds": wall, "results_finite": bool(np.isfinite(out["gof"][valid]).all())}


def k1_limits():
    """Limiter figures of the dominant kernel from the committed ncu capture of the CURRENT kernel (profiles/k1_limits.json,
    written by tools/ncu_limits.py from the raw ncu page), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_limits.json")) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return None


# ---- our arm ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nremmodfc_b200 import ops, sweep

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, SC, emp = load_inputs()
    strong = args.scaling == "strong"
    mapG = mapS = map_id = None
    n_maps = 1
    if strong:
        # ONE job for all ranks: the paper's three modalities x 20 000 simulations, sharded sim % world == rank (whole_sweep_both.py:64)
        norm = lambda m: m / m.mean()
        mapG = np.stack([np.ones(90), norm(d["map_ACh"]), norm(d["map_ACh_shuf"])])
        mapS = np.stack([np.ones(90), norm(d["map_NA"]), norm(d["map_NA_shuf"])])
        s1, g1, sg1 = sweep_grid(0, args.sims)
        n1 = len(s1)
        ids = sweep.shard_ids(3 * n1, rank, world)
        mid_all = ids // n1
        src, valid = sweep.pad_by_map(mid_all)                 # every 128-tile holds one map id
        gid = ids[src]
        seeds, dG, dS, map_id = s1[gid % n1], g1[gid % n1], sg1[gid % n1], (gid // n1).astype(np.int32)
        streams = (gid // n1).astype(np.uint64) << np.uint64(48) | (seeds.astype(np.uint64) << np.uint64(32)) | (gid % 400).astype(np.uint64)
        B, B_real, total_real, n_maps = len(src), int(valid.sum()), 3 * n1, 3
    else:
        B = B_real = args.sims
        seeds, dG, dS = sweep_grid(rank, B)
        # replicate id: unique per (seed, cell) over all ranks -> results do not depend on the sharding
        streams = (seeds.astype(np.uint64) << np.uint64(32)) | np.arange(B, dtype=np.uint64) % np.uint64(400)
        total_real = B * world
        valid = np.ones(B, bool)
    G0, s0 = np.full(B, 0.16), np.full(B, 7.68)                 # whole_sweep_both.py:30

    scale = args.horizon_scale
    pf = ops.make_params(90, int(FULL["n1"] * scale), int(FULL["n2"] * scale), int(FULL["n3"] * scale), P=0.4, rhoE=0.18, seed=2024)
    pw = ops.make_params(90, 200, 40_000, 60_000, P=0.4, rhoE=0.18, seed=2024)     # pre-warm pass: 1 % of the horizon
    kw = dict(kernel=args.kernel, bold_f32=not args.bold_f64, chunk_samples=args.chunk_samples, n_maps=n_maps)
    plan = sweep.SweepPlan(pf, B, peakfreq=args.peakfreq, **kw)
    sched = chunk_schedule(pf, plan.opts.chunk_samples or 250)
    ntot = len(sched)
    assert ntot == plan.chunks_total
    bounds = [ntot * c // SLICES for c in range(SLICES + 1)]
    fma_peak, _ = ops.measure_fma_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_args = dict(mapG=mapG, mapS=mapS, map_id=map_id)
    warm = sweep.SweepPlan(pw, B, bold_downsamp=100, **kw)
    warm.run(SC, emp, G0, dG, s0, dS, streams, **run_args)
    warm.close()
    if args.warmup > 0:                                          # W warm-up steps = W slices of the real sweep (then rewound)
        plan.begin(SC, G0, dG, s0, dS, streams, **run_args)
        for c in range(args.warmup):
            c %= SLICES
            plan.advance(bounds[c + 1] - bounds[c])
            torch.cuda.synchronize()
    barrier()
    ops.launch_count(reset=True)
    plan.set_profiling(True)
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, k1_ms, k1_launches, table, groups = 0.0, 0.0, 0, None, 1
    euler_steps, h2d, d2h, sweeps_done = 0, 0, 0, 0
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        c = k % SLICES
        if c == 0:
            plan.begin(SC, G0, dG, s0, dS, streams, **run_args)                 # H2D of this sweep's inputs
            h2d += plan.h2d_bytes
        plan.advance(bounds[c + 1] - bounds[c])
        euler_steps += sum(sched[bounds[c]:bounds[c + 1]])
        if c == SLICES - 1:
            out = plan.finish(emp)                                               # tail + D2H of the result table
            h2d += plan.h2d_bytes
            d2h += plan.d2h_bytes
            rows = np.concatenate([out["gof"].reshape(B, 16), out["mean"][:, None]], axis=1)
            if world > 1:
                gl = (ids[src] if strong else np.arange(B) + rank * B)[valid]
                table = sweep.gather_rows(gl, rows[valid], total_real)
            else:
                table = rows[valid]
            sweeps_done += 1
        pr = plan.profile()                                                      # blocks until the slice has finished
        dev_ms += pr["total_ms"]
        k1_ms += pr["integrator_ms"]
        k1_launches += pr["integrator_launches"]
        groups = pr["tile_groups"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    launches = ops.launch_count()
    if world > 1:
        t = torch.tensor([wall_ms, dev_ms, k1_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_ms, dev_ms, k1_ms = t.tolist()
    frac_done = euler_steps / float(sum(sched))                                  # sweeps' worth of integration in the timed region
    total_sims = total_real * frac_done
    value = total_sims / (dev_ms * 1e-3)
    e2e = total_sims / (wall_ms * 1e-3)
    k1_flops = float(B) * euler_steps * FLOP_PER_STEP                            # algorithmic flop of one rank's integrator launches
    tiles = (B + 127) // 128
    if groups == 1:      # one grid per launch covers the whole batch: event time of the launches is the kernel time
        achieved, timing, share = k1_flops / (k1_ms * 1e-3) / 1e12, "CUDA events around every integrator launch", k1_ms / dev_ms
        flop_per_launch, avg_ms = k1_flops / max(k1_launches, 1), k1_ms / max(k1_launches, 1)
    else:                # tile groups overlap on the device: charge the integrator with the WHOLE step (lower bound)
        # share: fraction of tile group 0's stream time spent inside its integrator launches (CUDA events around each of them;
        # the rest is that group's BOLD/filter launches and the FC/GoF tail) -- all groups run the same sequence
        achieved, timing, share = k1_flops / (dev_ms * 1e-3) / 1e12, f"{groups} tile-group streams overlap; whole-step device time charged", k1_ms / dev_ms
        flop_per_launch, avg_ms = k1_flops / groups / max(k1_launches, 1), k1_ms / max(k1_launches, 1)
    # HBM side of the roofline: the integrator's only algorithmic HBM traffic is the E samples it records
    tiles_per_launch = tiles / groups
    rows_per_launch = (plan.opts.chunk_samples or 250)
    alg_bytes = rows_per_launch * 90 * tiles_per_launch * 128 * 4
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6650.0, "of fallback (B200_PROFILING.md)"
    lim = k1_limits()
    if rank == 0:
        ok = bool(table is None or np.isfinite(table).all())
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        # SM cycles one tile spends per Euler step, all SMs counted as busy for the whole device time
        clk_per_tile_step = dev_ms * 1e-3 * sm_mhz * 1e6 * min(148, tiles) / (tiles * max(euler_steps, 1)) if groups > 1 else \
            k1_ms * 1e-3 * sm_mhz * 1e6 / max(euler_steps, 1)
        roof = {"bound": "fp32_fma", "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
                "kernel": "wc_batch_tc_kernel" if args.kernel != "fma" else "wc_batch_v0_kernel",
                "algorithmic_flop_per_euler_step": FLOP_PER_STEP, "kernel_share_of_step": share, "timing": timing,
                "launches_timed": k1_launches, "avg_launch_ms": avg_ms, "algorithmic_flop_per_launch": flop_per_launch,
                "peak_source": "measured in this run: register-only FFMA chains on all SMs (nrem_measure_fma_peak)",
                "note": "the SC.E contraction (82 % of the algorithmic flop) runs on tcgen05 tensor cores, so the FP32-FMA roof "
                        "can be exceeded; `composite` is the utilisation figure",
                "hbm": {"algorithmic_bytes_per_recording_launch": alg_bytes, "achieved": alg_bytes / (avg_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (avg_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src}}
        if lim:
            # composite bound: the kernel cannot finish an Euler step of a tile faster than its busiest resource allows
            floors = {k: lim["floors_clk_per_tile_step"][k] for k in ("xu_mufu", "issue", "tensor", "fma")}
            binding = max(floors, key=floors.get)
            roof["traffic"] = lim["dram_bytes_per_launch"] / lim["tiles_in_capture"] * tiles_per_launch
            roof["traffic_note"] = "dram read+write bytes of one recording launch (ncu --set full), scaled to the tiles of one launch; " + lim["source"]
            roof["composite"] = {"floors_clk_per_tile_step": floors, "binding": binding, "measured_clk_per_tile_step": clk_per_tile_step,
                                 "frac": floors[binding] / clk_per_tile_step, "pipe_busy_pct_ncu": lim["pipe_busy_pct"],
                                 "definition": "floor = SM cycles one Euler step of a 128-simulation tile would need if that resource alone were the "
                                               "limit = its busy fraction in the committed ncu capture of this kernel x the capture's cycles per step "
                                               "(cross-check from the SASS: XU = MUFU lane-ops / 16 per clk, issue = warp instructions / 4 schedulers, "
                                               "tensor = 36 MMAs x 48 clk); measured = device time of the live run x SM clock x busy SMs / tile-steps, "
                                               "i.e. it also carries the BOLD/filter launches and the tile-group scheduling of the sweep",
                                 "source": lim["source"]}
        else:
            roof["traffic"] = None
        line = {
            "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_ms / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32 (integrator; coupling 3xTF32 on tcgen05) + f64 (filter/FC/GoF)" if args.kernel in ("auto", "tc3") else "f32",
            "data": DATA,
            "config": {"workload": ("configs[1]: homogeneous G x sigma sweep x 50 seeds = 20000 sims per GPU (AAL90 SC, "
                                    "G0=0.16, sigma0=7.68, 1+400+600 s, GoF vs W/N1/N2/N3)") if not strong else
                                   ("strong scaling: ONE job = configs[1]+[2]+[3] (homogeneous + map + shuffled, 3 x 20000 sims) "
                                    "sharded sim % world == rank"),
                       "step": f"one of {SLICES} equal time slices of the whole sweep (every simulation advanced by 1/{SLICES} of its "
                               f"horizon, state carried in the plan); slice 0 uploads the inputs, slice {SLICES - 1} adds the backward "
                               "filter + FC + GoF tail, the D2H of the table and the gather",
                       "sweeps_in_timed_region": frac_done, "whole_sweeps_finished": sweeps_done,
                       "sims_per_gpu": B_real, "padded_sims_per_gpu": B, "euler_steps_per_sim": int(sum(sched)), "kernel": args.kernel,
                       "bold_state": "f64" if args.bold_f64 else "f32", "horizon_scale": scale, "peakfreq_column": bool(args.peakfreq),
                       "warmup_pass": "whole pipeline on 1% of the horizon, then W slices of the real sweep",
                       "l2": "inputs are register/SMEM resident; each step streams its own E samples (> L2) through HBM",
                       "results_finite": ok},
            "e2e": {"value": e2e, "unit": "sims/s", "h2d_bytes_per_step": h2d / max(args.steps, 1), "d2h_bytes_per_step": d2h / max(args.steps, 1),
                    "note": "H2D happens in the first slice of a sweep and D2H in the last; bytes are averaged over the steps"},
            "gpu_launches": int(launches),
            "node_seconds_per_s": value * NODE_SECONDS_PER_SIM,
            "roofline": roof,
            "clocks": clocks,
        }
        if world == 1 and not strong:
            if not args.no_modalities:
                line["modalities"] = [modality_leg(args, d, SC, emp, w) for w in ("map", "shuffled")]
            if lim and args.kernel in ("auto", "tc3"):
                k1 = k1_only_leg(d)
                k1["clk_per_tile_step"] = k1["us_per_euler_step"] * sm_mhz
                k1["composite_frac"] = floors[binding] / k1["clk_per_tile_step"]
                if lim.get("sm_mhz_in_capture"):
                    # the cycle counter of the ncu capture of this kernel ran at sm_mhz_in_capture while nvidia-smi showed sm_mhz: cycles are
                    # wall time x the ACTUAL clock, so this is the figure comparable with frac_of_composite_bound_under_ncu
                    k1["sm_mhz_in_ncu_capture"] = lim["sm_mhz_in_capture"]
                    k1["clk_per_tile_step_at_capture_clock"] = k1["us_per_euler_step"] * lim["sm_mhz_in_capture"]
                    k1["composite_frac_at_capture_clock"] = floors[binding] / k1["clk_per_tile_step_at_capture_clock"]
                k1["note"] = ("the integrator alone on one full wave: the kernel's own fraction of its composite (" + binding + ") bound; "
                              "`composite.frac` above is the same floor over the whole sweep's device time")
                roof["composite"]["kernel_only"] = k1
            if not args.no_config5:
                line["config5"] = config5_leg()
            if not args.no_config1:
                line["config1"] = config1_leg(d, SC, emp)
            if not args.no_cpu:
                cb = cpu_baseline(frac=args.cpu_frac, kind=args.cpu_kind)
                cb.pop("wall_s", None)
                line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=SLICES, help=f"timed steps; {SLICES} steps = one whole sweep")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--sims", type=int, default=20000, help="simulations per GPU (default: the full 50 x 20 x 20 sweep)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "fma", "tc", "tc3"])
    ap.add_argument("--bold-f64", action="store_true")
    ap.add_argument("--peakfreq", action="store_true", help="also compute the Welch peak frequency column (whole_sweep_both.py:90-95)")
    ap.add_argument("--chunk-samples", type=int, default=0)
    ap.add_argument("--horizon-scale", type=float, default=1.0, help="DEBUG ONLY: shorten every phase (numbers are then not bench values)")
    ap.add_argument("--cpu-frac", type=float, default=None, help="fraction of the horizon per CPU sample")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the short large-connectome leg (configs[4])")
    ap.add_argument("--no-config1", action="store_true", help="skip the single full-length run through the drop-in surface (configs[0])")
    ap.add_argument("--no-modalities", action="store_true", help="skip the map / shuffled-map sweeps (configs[2], configs[3])")
    args = ap.parse_args()
    if args.cpu_frac is None:
        args.cpu_frac = 0.15          # >= 10 s of wall per sample on a 16-core box
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on communicator
    # creation), so everything but our own print() goes to stderr: fd 1 is pointed at fd 2 and sys.stdout keeps the real one.
    sys.stdout.flush()
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_out
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        real_out.flush()


if __name__ == "__main__":
    main()
