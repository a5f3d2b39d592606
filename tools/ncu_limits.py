"""Limiter figures of the dominant kernel from an `ncu --set full` capture -> profiles/k1_limits.json (read by bench.py) and a
markdown summary for profiles/.

    python tools/ncu_limits.py gpurun_out/r02c_k1.ncu-rep [steps_in_launch=5000] > profiles/r02_ncu_integrator.md

A floor is the number of SM cycles one Euler step of a 128-simulation tile would need if that resource alone were the limit:
busy fraction of the resource (ncu, % of peak sustained while active) x the measured cycles per step.  The composite bound of
the kernel is the largest floor; bench.py reports floor / measured cycles of the live run next to the FP32-FMA figure.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5000

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, val)}


def num(name, default=None):
    if name not in M:
        if default is not None:
            return default
        raise KeyError(name)
    return float(M[name][0].replace(",", ""))


def scaled(name):
    """value in base units (bytes) whatever prefix ncu chose"""
    v, u = num(name), M[name][1].lower()
    for pre, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(pre):
            return v * f
    return v


tiles = int(num("launch__grid_size"))
cyc = num("sm__cycles_elapsed.max")
clk_step = cyc / steps
pct = {
    "issue": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "xu_mufu": num("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active"),
    "fma": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "tensor": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
    "alu": num("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
    "lsu": num("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", 0.0),
}
floors = {k: pct[k] / 100.0 * clk_step for k in ("xu_mufu", "issue", "tensor", "fma")}
dram = scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")
out = {
    "source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '')} (ncu --set full --clock-control none, one {steps}-step launch of "
              f"{M['Kernel Name'][0] if 'Kernel Name' in M else 'the integrator'}, {tiles} tiles); written by tools/ncu_limits.py",
    "kernel": M.get("Kernel Name", ("?",))[0],
    "tiles_in_capture": tiles, "steps_in_capture": steps,
    "gpu_time_ms": num("gpu__time_duration.sum") * (1e-3 if M["gpu__time_duration.sum"][1].lower().startswith("us") else
                                                    1e-6 if M["gpu__time_duration.sum"][1].lower().startswith("ns") else 1.0),
    "clk_per_tile_step_ncu": clk_step,
    # SM clock the kernel actually ran at in the capture (cycle counter over wall time; --clock-control none): lower than the
    # clocks.sm figure nvidia-smi reports while this kernel runs
    "sm_mhz_in_capture": None,
    "warp_instructions_per_tile_step": num("smsp__inst_executed.sum") / (tiles * steps),
    "registers_per_thread": num("launch__registers_per_thread"),
    "pipe_busy_pct": pct,
    "floors_clk_per_tile_step": floors,
    "binding": max(floors, key=floors.get),
    "frac_of_composite_bound_under_ncu": max(floors.values()) / clk_step,
    "dram_bytes_per_launch": dram,
}
out["sm_mhz_in_capture"] = cyc / (out["gpu_time_ms"] * 1e-3) / 1e6
with open(os.path.join(ROOT, "profiles", "k1_limits.json"), "w") as fh:
    json.dump(out, fh, indent=1)

want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
want += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
               and num(h, 0.0) >= 0.05)
print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (one launch)\n")
print("| metric | value |\n|---|---|")
print(f"| Kernel Name | {out['kernel']} |")
for h in want:
    if h in M:
        print(f"| {h} [{M[h][1]}] | {M[h][0]} |")
print(f"\nPer Euler step of one tile: {clk_step:.0f} SM cycles, {out['warp_instructions_per_tile_step']:.0f} warp instructions.")
print("Floors (busy % x cycles per step): " + ", ".join(f"{k} {v:.0f}" for k, v in floors.items()) +
      f" -> binding {out['binding']}, {100 * out['frac_of_composite_bound_under_ncu']:.1f} % of the composite bound under ncu.")
print(f"DRAM read+write per launch: {dram / 1e6:.1f} MB.")
