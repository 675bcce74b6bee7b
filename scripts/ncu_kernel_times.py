"""Per-kernel durations from an `ncu --metrics gpu__time_duration.sum --csv` log: python scripts/ncu_kernel_times.py file.csv
[--last-step]  (--last-step keeps only the launches after the LAST posenc_rows_kernel / first kernel of a training step)"""
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
recs = []
for r in csv.DictReader(rows):
    n = r["Kernel Name"].split("(")[0].split("::")[-1]
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r["Metric Unit"], 1)
    recs.append((n, round(v, 1)))
if "--last-step" in sys.argv:
    starts = [i for i, (n, _) in enumerate(recs) if n.startswith("posenc_rows_kernel")]
    if starts:
        recs = recs[starts[-1]:]
agg = {}
for n, v in recs:
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:36s} n={len(v):3d} total={sum(v):9.1f} us  {100 * sum(v) / tot:5.1f}%  first: {v[:6]}")
print(f"TOTAL {tot:.1f} us over {len(recs)} launches")
