#!/bin/bash
# Round-2 pass Y: ncu --set full of the fused linear-attention kernels of one step (final binary), raw-metric CSV.
TAG=${1:-r02y}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
timeout 200 ncu --set full --clock-control none -k "regex:linattn_(kv2|kv|mix|out2|out)_kernel" -s 15 -c 15 -f -o gpurun_out/${TAG}_linattn $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-100
ncu -i gpurun_out/${TAG}_linattn.ncu-rep --page raw --csv > gpurun_out/${TAG}_linattn_raw.csv 2>/dev/null; ls -la gpurun_out/${TAG}_linattn*
