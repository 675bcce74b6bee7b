"""Per-kernel durations from an `ncu --metrics gpu__time_duration.sum --csv` log: python scripts/ncu_kernel_times.py file.csv"""
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = {}
for r in csv.DictReader(rows):
    n = r["Kernel Name"].split("(")[0].split("::")[-1]
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r["Metric Unit"], 1)
    agg.setdefault(n, []).append(round(v, 1))
for k, v in agg.items():
    print(f"{k:32s} n={len(v):3d} total={sum(v):9.1f} us  first: {v[:8]}")
