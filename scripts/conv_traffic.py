"""DRAM traffic of one sampling step per kernel family from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list.  usage: conv_traffic.py launches.csv <workload> <batch> > profiles/conv_traffic.json
One step = from one stem_conv launch (first kernel of the eps-net) up to the next."""
import csv
import json
import re
import sys

path, workload, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
per = {}
for r in csv.DictReader(l for l in open(path) if l.startswith('"')):
    e = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "ns": 0.0, "bytes": 0.0})
    try:
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
    except ValueError:
        continue
    if r["Metric Name"] == "gpu__time_duration.sum":
        e["ns"] = v
    elif r["Metric Name"].startswith("dram__bytes"):
        e["bytes"] += v
rows = [per[k] for k in sorted(per)]
stems = [i for i, e in enumerate(rows) if "stem_conv" in e["name"]]
a, b = stems[-2], stems[-1]          # the last complete step in the capture
fam = {}
for e in rows[a:b]:
    n = e["name"]
    key = ("conv_gemm" if "conv_gemm" in n else "groupnorm" if "groupnorm" in n else "linattn_fused" if re.search(r"linattn_(kv|mix|out)2?_kernel", n)
           else "other")
    f = fam.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "ncu_ms": 0.0})
    f["launches"] += 1
    f["dram_bytes"] += e["bytes"]
    f["ncu_ms"] += e["ns"] / 1e6
print(json.dumps({workload: {"batch": batch, "launches_in_step": b - a, "families": fam,
                             "source": f"{path}: ncu dram__bytes_read.sum + dram__bytes_write.sum, one step, cold-cache serialised replay"}}, indent=1))
