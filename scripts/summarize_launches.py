"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
per_id = OrderedDict()
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("unnamed>::", "").replace("hd::<", "")
    e = per_id.setdefault(int(r["ID"]), {"name": name, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    try:
        val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
    except ValueError:
        continue
    m = r.get("Metric Name")
    if m == "gpu__time_duration.sum":
        e["ns"] = val
    elif m == "dram__bytes_read.sum":
        e["rd"] = val
    elif m == "dram__bytes_write.sum":
        e["wr"] = val
rows = list(per_id.values())[skip:]
agg = OrderedDict()
for e in rows:
    a = agg.setdefault(e["name"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e["ns"]
    a[2] += e["rd"] + e["wr"]
tot = sum(v[1] for v in agg.values())
print(f"source: {path} (first {skip} launches skipped); {len(rows)} launches, {tot / 1e6:.3f} ms total device time\n")
print("| kernel | launches | total ms | share | avg us | DRAM MB (rd+wr) total |")
print("|---|---:|---:|---:|---:|---:|")
for n, (c, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} | {by / 1e6:.1f} |")
