"""The reference's three 50-seed G x sigma sweeps (homogeneous, NA/ACh maps, shuffled maps) end to end on the GPU path.

    python tools/run_full_sweeps.py --out gpurun_out/sweeps                       # one GPU
    torchrun --nnodes=1 --nproc-per-node 8 tools/run_full_sweeps.py --out ...     # one rank per GPU

Replaces the SLURM array of whole_sweep_both.py / whole_sweep_both_maps.py (64 ranks x ~35 h on the authors' cluster):
for every modality the 50 x 20 x 20 product (whole_sweep_both.py:57-61 with the committed grid of
whole_sweep_both_maps.py:92-93) is sharded over the ranks with the reference's own rule (sim % threads == rank,
whole_sweep_both.py:64), every rank runs ONE batched call, the rows are gathered and rank 0 writes
  * output/sweep_<modality>.txt   the collapsed 20-column table that heatmaps.py / figures read (pd.read_csv),
  * report.md                     per-cell statistics against the reference's committed tables
                                  (tests/golden/sweep_cell_stats.npz) and the e/|corr| optima (heatmaps.py:28-58)
                                  against the optima the authors list in run_many_seeds.py:34-47.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COLS = ["ssimW", "ssimN1", "ssimN2", "ssimN3", "corrW", "corrN1", "corrN2", "corrN3", "eW", "eN1", "eN2", "eN3", "sync", "meta",
        "mean", "peakfreq"]
# run_many_seeds.py:34-47 (delta_G, delta_sigma) per state
PUBLISHED = {"homo": {"W": (0.0, 0.0), "N1": (0.04, 0.0), "N2": (0.0, 0.0), "N3": (-0.04, 0.04)},
             "map": {"W": (-0.02, -0.02), "N1": (0.18, -0.02), "N2": (0.02, -0.04), "N3": (0.02, -0.12)},
             "shuf": {"W": (0.0, 0.0), "N1": (0.0, 0.04), "N2": (0.0, 0.0), "N3": (0.0, -0.04)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweeps")
    ap.add_argument("--modalities", default="homo,map,shuf")
    ap.add_argument("--seeds", type=int, default=50)
    ap.add_argument("--no-peakfreq", action="store_true")
    ap.add_argument("--horizon-scale", type=float, default=1.0, help="DEBUG: shorten every phase")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from nremmodfc_b200 import ops, sweep, table

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    d = np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"))
    stats = np.load(os.path.join(ROOT, "tests", "golden", "sweep_cell_stats.npz"))
    emp = np.stack([d[s] for s in ("W", "N1", "N2", "N3")])
    norm = lambda m: m / m.mean()                                                   # whole_sweep_both_maps.py:54,62
    maps = {"homo": (np.ones(90), np.ones(90)), "map": (norm(d["map_ACh"]), norm(d["map_NA"])),
            "shuf": (norm(d["map_ACh_shuf"]), norm(d["map_NA_shuf"]))}
    dGv = np.linspace(-0.1, 0.3, 20, endpoint=False)                                # whole_sweep_both_maps.py:92-93
    dSv = np.linspace(-0.2, 0.2, 20, endpoint=False)
    sc = args.horizon_scale
    p = ops.make_params(90, int(10_000 * sc), int(4_000_000 * sc), int(6_000_000 * sc), P=0.4, rhoE=0.18, seed=20241209)
    os.makedirs(args.out, exist_ok=True)
    report = ["# Full sweeps on the GPU path vs the reference's committed tables", "",
              f"{world} GPU(s), {args.seeds} seeds x 20 x 20 cells per modality, horizon scale {sc:g}, peakfreq {'off' if args.no_peakfreq else 'on'}", ""]
    for mi, mod in enumerate(args.modalities.split(",")):
        seeds, dG, dS = sweep.product_grid(np.arange(args.seeds), dGv, dSv)
        n = len(seeds)
        mine = sweep.shard_ids(n, rank, world)                                      # sim % threads == rank
        plan = sweep.SweepPlan(p, len(mine), n_maps=1, K=4, peakfreq=not args.no_peakfreq)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        streams = (np.uint64(mi) << np.uint64(48)) | mine.astype(np.uint64)         # replicate id: unique per (modality, seed, cell)
        out = plan.run(d["SC"], emp, 0.16, dG[mine], 7.68, dS[mine], streams, mapG=maps[mod][0][None], mapS=maps[mod][1][None])
        rows = table.rows_from_sweep(out, rank, seeds[mine], dG[mine], dS[mine], out["peakfreq"])
        full = sweep.gather_rows(mine, rows, n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        plan.close()
        if rank != 0:
            continue
        path = os.path.join(args.out, f"sweep_{mod}.txt")
        with open(path, "w") as f:
            f.write(",".join(table.COLUMNS) + "\n")
            for r in full:
                f.write(",".join([str(int(r[0])), str(int(r[1]))] + [f"{v:.4f}" for v in r[2:]]) + "\n")
        tab = table.read_table(path)
        # per-cell statistics vs the committed tables
        zs, worst = [], []
        for i, g in enumerate(np.round(dGv, 4)):
            for j, s in enumerate(np.round(dSv, 4)):
                sel = (np.round(tab["delta_G"], 4) == g) & (np.round(tab["delta_sigma"], 4) == s)
                for k, c in enumerate(COLS):
                    if c == "peakfreq" and args.no_peakfreq:
                        continue
                    mine_v = tab[c][sel]
                    ref_m, ref_sd, ref_n = stats[f"{mod}_mean"][i, j, k], stats[f"{mod}_sd"][i, j, k], stats[f"{mod}_n"][i, j]
                    se = np.sqrt(ref_sd ** 2 / ref_n + mine_v.std(ddof=1) ** 2 / len(mine_v)) + (0.07 if c == "peakfreq" else 1e-3)
                    z = (mine_v.mean() - ref_m) / se
                    zs.append(z)
                    worst.append((abs(z), c, g, s, mine_v.mean(), ref_m))
        zs = np.asarray(zs)
        worst.sort(reverse=True)
        opt = table.euccorr_optima(tab)
        report += [f"## {mod}: {n} simulations in {dt:.1f} s ({n / dt:.0f} sims/s incl. H2D/D2H and the gather)", "",
                   f"cell statistics vs committed table ({len(zs)} comparisons = 400 cells x {len(zs) // 400} columns): "
                   f"mean z {zs.mean():+.3f}, sd of z {zs.std():.3f}, |z| < 3 in {np.mean(np.abs(zs) < 3) * 100:.2f} %, max |z| {np.abs(zs).max():.2f}", "",
                   "largest deviations: " + "; ".join(f"{c} at ({g:+.2f},{s:+.2f}): {a:.4f} vs {b:.4f} (z {z:.1f})" for z, c, g, s, a, b in worst[:4]), "",
                   "| state | optimum here (dG, dsigma, e/|corr|) | listed in run_many_seeds.py:34-47 |", "|---|---|---|"]
        for st in ("W", "N1", "N2", "N3"):
            report.append(f"| {st} | ({opt[st][0]:+.2f}, {opt[st][1]:+.2f}, {opt[st][2]:.2f}) | ({PUBLISHED[mod][st][0]:+.2f}, {PUBLISHED[mod][st][1]:+.2f}) |")
        report.append("")
        print(report[-9], flush=True)
    if rank == 0:
        with open(os.path.join(args.out, "report.md"), "w") as f:
            f.write("\n".join(report) + "\n")
        print("\n".join(report))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
