"""Condense an `ncu --set full` report into a markdown table of the metrics the roofline discussion uses.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [max_rows] > profiles/rNN_x_ncu.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
max_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [
    ("ID", "id"), ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM rd"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch values, cold-cache replay)\n")
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:2 + max_rows]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("void hd::<unnamed>::", "").split("(")[0][:40]
        else:
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
