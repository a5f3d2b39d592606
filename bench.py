#!/usr/bin/env python
"""Benchmark of the hot path: WC + BOLD + FC + GoF simulations per second on the reference's
homogeneous G x sigma x 50-seed sweep (BASELINE.json configs[1]; whole_sweep_both.py grid).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the whole pipeline over the 20 000-simulation sweep (50 seeds x 20 dG x 20 dsigma,
AAL90, full length 1 + 400 + 600 s = 10.01 M Euler steps each) through the public host API
(`nremmodfc_b200.sweep.SweepPlan.run`: NumPy in, NumPy out).  Inside every timed step
  * `e2e`   = wall clock of the API call (pinned H2D of SC/targets/parameters, D2H of the GoF table),
  * `value` = the same pass timed on the device with CUDA events (inputs resident in HBM -> GoF in HBM).
Multi-GPU (torchrun, one rank per GPU): weak scaling — every rank runs its own 20 000-simulation sweep
(seeds 50*rank ... 50*rank+49), no data-path collective, one final all-gather of the GoF table.
`--impl reference` times the reference's CPU path (oracle port in C, one process per host core, the
reference's own parallel model whole_sweep_both.py:23-24) on a bounded sample of the same workload.
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


def load_inputs():
    d = np.load(os.path.join(ROOT, "tests", "golden", "aal90_inputs.npz"))
    return d["SC"], np.stack([d[s] for s in ("W", "N1", "N2", "N3")])


def sweep_grid(rank, n_sims):
    """whole_sweep_both.py:57-61 with the committed table's grid (whole_sweep_both_maps.py:92-93)."""
    from nremmodfc_b200 import sweep
    seeds = np.arange(50) + 50 * rank
    dG = np.linspace(-0.1, 0.3, 20, endpoint=False)
    dS = np.linspace(-0.2, 0.2, 20, endpoint=False)
    s, g, sg = sweep.product_grid(seeds, dG, dS)
    return s[:n_sims], g[:n_sims], sg[:n_sims]


# ---- CPU baseline (oracle port) ---------------------------------------------------------------
def _cpu_worker(args):
    seed, frac = args
    from oracle import bold_oracle, cwrap, wc_oracle
    SC, emp = load_inputs()
    p = wc_oracle.params(P=0.4, rhoE=0.18)
    n1, n2, n3 = int(FULL["n1"] * frac), int(FULL["n2"] * frac), int(FULL["n3"] * frac)
    E = cwrap.wc_run(SC, 0.16, 7.68, n1, n2, n3, seed=1, stream=seed, p=p, want="E", fast_rng=True)
    bold = cwrap.bold_sim(E, 0.04)
    neq = min(2000, max(0, E.shape[0] - 1100))
    y = bold_oracle.filt_decimate(bold, 1000 if E.shape[0] > 12000 else 100, neq, 0.04)
    FC = bold_oracle.fc(y)
    g = [bold_oracle.get_all_metrics(FC, emp[k]) for k in range(4)]
    return float(g[0][0])


def config5_leg(N=1000, B=4096, steps=40000):
    """BASELINE configs[4] (scaled synthetic connectome, 1000-node random SC, 4096 instances) on the per-step tcgen05 kernel
    (csrc/wc_big.cuh): Euler steps/s of the whole batch and the coupling rate, CUDA events around the step launches, outside
    the timed region of the headline metric.  Roofline: tensor (TF32 main pass + two BF16 correction passes)."""
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
        bf16, src = 2250.0, "nominal dense bf16 (B200_PROFILING.md fallback; TF32 = half)"
    alg = 2.0 * B * N * N                            # flop per step, as the reference's dgemv
    ideal_us = alg / (bf16 / 2 * 1e12) * 1e6 + 2 * alg / (bf16 * 1e12) * 1e6
    return {"workload": f"configs[4]: {N}-node random SC, {B} instances, {steps} Euler steps, kernel tcb", "us_per_step": us,
            "steps_per_s": 1e6 / us, "node_updates_per_s": B * N / us * 1e6, "results_finite": bool(np.isfinite(fin).all()),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": alg / us / 1e6, "unit": "TFLOP/s (algorithmic 2*B*N^2 per step)",
                         "peak": alg / ideal_us / 1e6, "frac": ideal_us / us, "traffic": None,
                         "peak_note": "1 TF32 pass + 2 BF16 passes at their tensor peaks; " + src,
                         "kernel": "wc_big_step_kernel<4>", "profile": "profiles/r01_big_connectome.md"}}


def cpu_baseline(frac=0.05, reps=1):
    """Times the oracle's C restatement of the reference pipeline on all host cores; returns sims/s of
    FULL-length simulations (the cost is linear in the number of Euler steps)."""
    import multiprocessing as mp
    from oracle import cwrap
    cwrap.build()
    cores = os.cpu_count() or 1
    tasks = [(i, frac) for i in range(cores * reps)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, 0.0005)] * cores)          # warm: page in scipy, build tables
        t0 = time.perf_counter()
        pool.map(_cpu_worker, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    sims_per_s = len(tasks) * frac / wall
    return {"value": sims_per_s, "unit": "sims/s", "cores": cores, "kind": "port",
            "sample": f"{len(tasks)} sims x {frac:g} of the 10.01M-step horizon (+BOLD/filter/FC/GoF), one process per core, "
                      f"{wall:.1f} s wall; scaled linearly in steps",
            "wall_s": wall}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(frac=args.cpu_frac)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    wall = float(np.mean([r["wall_s"] for r in vals]))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "sims/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: homogeneous G x sigma sweep x 50 seeds (AAL90, 1+400+600 s)", "sample": vals[-1]["sample"]},
            "cpu_baseline": {k: vals[-1][k] for k in ("unit", "cores", "kind", "sample")} | {"value": v},
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
    SC, emp = load_inputs()
    B = args.sims
    seeds, dG, dS = sweep_grid(rank, B)
    # replicate id: unique per (seed, cell) over all ranks -> results do not depend on the sharding
    streams = (seeds.astype(np.uint64) << np.uint64(32)) | np.arange(B, dtype=np.uint64) % np.uint64(400)
    G0, s0 = np.full(B, 0.16), np.full(B, 7.68)                 # whole_sweep_both.py:30

    scale = args.horizon_scale
    pf = ops.make_params(90, int(FULL["n1"] * scale), int(FULL["n2"] * scale), int(FULL["n3"] * scale), P=0.4, rhoE=0.18, seed=2024)
    pw = ops.make_params(90, 200, 40_000, 60_000, P=0.4, rhoE=0.18, seed=2024)     # warm-up pass: 1 % of the horizon
    plan = sweep.SweepPlan(pf, B, kernel=args.kernel, bold_f32=not args.bold_f64, chunk_samples=args.chunk_samples, peakfreq=args.peakfreq)
    warm = sweep.SweepPlan(pw, B, kernel=args.kernel, bold_f32=not args.bold_f64, chunk_samples=args.chunk_samples, bold_downsamp=100)
    fma_peak, _ = ops.measure_fma_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        warm.run(SC, emp, G0, dG, s0, dS, streams)
    warm.close()
    barrier()
    ops.launch_count(reset=True)
    plan.set_profiling(True)
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, k1_ms, k1_launches, table, groups = 0.0, 0.0, 0, None, 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = plan.run(SC, emp, G0, dG, s0, dS, streams)
        pr = plan.profile()
        dev_ms += pr["total_ms"]
        k1_ms += pr["integrator_ms"]
        k1_launches += pr["integrator_launches"]
        groups = pr["tile_groups"]
        rows = np.concatenate([out["gof"].reshape(B, 16), out["mean"][:, None]], axis=1)
        table = sweep.gather_rows(np.arange(B) + rank * B, rows, B * world) if world > 1 else rows
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    launches = ops.launch_count()
    if world > 1:
        t = torch.tensor([wall_ms, dev_ms, k1_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_ms, dev_ms, k1_ms = t.tolist()
    total_sims = B * world * args.steps * scale
    value = total_sims / (dev_ms * 1e-3)
    e2e = total_sims / (wall_ms * 1e-3)
    k1_flops = B * args.steps * scale * STEPS_PER_SIM * FLOP_PER_STEP          # algorithmic flop of one rank's launches
    if groups == 1:      # one grid per launch covers the whole batch: event time of the launches is the kernel time
        achieved, timing, share = k1_flops / (k1_ms * 1e-3) / 1e12, "CUDA events around every integrator launch", k1_ms / dev_ms
        flop_per_launch, avg_ms = k1_flops / max(k1_launches, 1), k1_ms / max(k1_launches, 1)
    else:                # tile groups overlap on the device: charge the integrator with the WHOLE step (lower bound)
        # share: fraction of tile group 0's stream time spent inside its integrator launches (CUDA events around each of them;
        # the rest is that group's BOLD/filter launches and the FC/GoF tail) -- all groups run the same sequence
        achieved, timing, share = k1_flops / (dev_ms * 1e-3) / 1e12, f"{groups} tile-group streams overlap; whole-step device time charged", k1_ms / dev_ms
        flop_per_launch, avg_ms = k1_flops / groups / max(k1_launches, 1), k1_ms / max(k1_launches, 1)
    # HBM side of the roofline: the integrator's only algorithmic HBM traffic is the E samples it records
    # (rows x 90 x sims x 4 B per recording launch); ncu (profiles/r01_final_ncu_integrator.md) measured
    # dram read+write = 1.713e9 B for a 148-tile, 250-row launch whose algorithmic bytes are 1.705e9.
    tiles_per_launch = (B + 127) // 128 / groups
    rows_per_launch = (args.chunk_samples or 250)
    alg_bytes = rows_per_launch * 90 * tiles_per_launch * 128 * 4
    ncu_traffic = 1.713e9 / 148 * tiles_per_launch
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6650.0, "of fallback (B200_PROFILING.md)"
    if rank == 0:
        ok = bool(np.isfinite(table).all())
        line = {
            "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (integrator; coupling 3xTF32 on tcgen05) + f64 (filter/FC/GoF)" if args.kernel in ("auto", "tc3") else "f32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: homogeneous G x sigma sweep x 50 seeds = 20000 sims per GPU (AAL90 SC, "
                                   "G0=0.16, sigma0=7.68, 1+400+600 s, GoF vs W/N1/N2/N3)",
                       "sims_per_gpu": B, "euler_steps_per_sim": int(STEPS_PER_SIM * scale), "kernel": args.kernel,
                       "bold_state": "f64" if args.bold_f64 else "f32", "horizon_scale": scale, "peakfreq_column": bool(args.peakfreq),
                       "warmup_pass": "same batch, 1% of the horizon", "l2": "inputs are register/SMEM resident; each step streams "
                       "its own E samples (> L2) through HBM", "results_finite": ok},
            "e2e": {"value": e2e, "unit": "sims/s", "h2d_bytes_per_step": plan.h2d_bytes, "d2h_bytes_per_step": plan.d2h_bytes},
            "gpu_launches": int(launches),
            "node_seconds_per_s": value * NODE_SECONDS_PER_SIM,
            "roofline": {"bound": "fp32_fma", "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
                         "traffic": ncu_traffic, "traffic_note": "dram read+write bytes of one recording launch, ncu --set full, "
                         "profiles/r01_final_ncu_integrator.md, scaled to the tiles of one launch", "kernel": "wc_batch_tc_kernel" if args.kernel != "fma" else "wc_batch_v0_kernel",
                         "algorithmic_flop_per_euler_step": FLOP_PER_STEP, "kernel_share_of_step": share, "timing": timing,
                         "launches_timed": k1_launches, "avg_launch_ms": avg_ms, "algorithmic_flop_per_launch": flop_per_launch,
                         "peak_source": "measured in this run: register-only FFMA chains on all SMs (nrem_measure_fma_peak)",
                         "note": "the SC.E contraction (82 % of the algorithmic flop) runs on tcgen05 tensor cores, so the FP32-FMA roof can be exceeded",
                         "limiters_ncu": {"issue_slots_pct": 62.9, "xu_mufu_pipe_pct": 64.9, "fma_pipe_pct": 42.7, "tensor_pipe_pct": 26.0,
                                          "source": "profiles/r01_final_ncu_integrator.md (static, from the ncu capture of this kernel)"},
                         "hbm": {"algorithmic_bytes_per_recording_launch": alg_bytes, "achieved": alg_bytes / (avg_ms * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (avg_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src}},
            "clocks": clocks,
        }
        if world == 1 and not args.no_config5:
            line["config5"] = config5_leg()
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline(frac=args.cpu_frac)
            cb.pop("wall_s", None)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sims", type=int, default=20000, help="simulations per GPU (default: the full 50 x 20 x 20 sweep)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "fma", "tc", "tc3"])
    ap.add_argument("--bold-f64", action="store_true")
    ap.add_argument("--peakfreq", action="store_true", help="also compute the Welch peak frequency column (whole_sweep_both.py:90-95)")
    ap.add_argument("--chunk-samples", type=int, default=0)
    ap.add_argument("--horizon-scale", type=float, default=1.0, help="DEBUG ONLY: shorten every phase (numbers are then not bench values)")
    ap.add_argument("--cpu-frac", type=float, default=0.05)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the short large-connectome leg (configs[4])")
    args = ap.parse_args()
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
