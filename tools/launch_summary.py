"""Summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`):
launches, total and average duration and share per (kernel, grid, block).

    python tools/launch_summary.py gpurun_out/X.csv > profiles/rNN_launch_list_summary.txt
"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append(((r["Kernel Name"][:70], r["Grid Size"], r["Block Size"]), ms))
agg = collections.OrderedDict()
for k, ms in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print("kernel, grid, block, launches, total_ms, avg_ms, share")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]}, {k[1]}, {k[2]}, {n}, {ms:.3f}, {ms / n:.4f}, {100 * ms / tot:.2f}%")
print(f"# {len(rows)} launches, {tot:.1f} ms serialised under ncu")
