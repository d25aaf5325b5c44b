"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: total device time per kernel, share, launches."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0.0, 0])
for k, v in rows:
    name = k.split("(")[0]
    agg[name][0] += v
    agg[name][1] += 1
print("kernels profiled: %d   total device time: %.2f ms" % (len(rows), tot / 1e3))
print("%-60s %10s %7s %8s %10s" % ("kernel", "total ms", "share", "launches", "avg us"))
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-60s %10.3f %6.1f%% %8d %10.2f" % (name[:60], t / 1e3, 100 * t / tot, n, t / n))
