"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("hd::<", "")
    rows.append((int(r["ID"]), name, float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
rows = rows[skip:]
agg = OrderedDict()
for _, n, ns, *_ in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values())
print(f"source: {path} (first {skip} launches skipped); {len(rows)} launches, {tot / 1e6:.3f} ms total device time\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |")
