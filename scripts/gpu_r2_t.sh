#!/bin/bash
# Round-2 GPU pass T (profiler, final binary): ncu launch list with DRAM bytes (unconditional Unet, B = 256) and one --set full
# pass over one step's 63 conv_gemm launches, reduced to its raw-metric CSV on the box (the report is ~100 MB).
TAG=${1:-r02t}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 420 --csv --log-file gpurun_out/${TAG}_launches_unet_uncond.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/ncu_launch.log | cut -c1-120
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none -k "regex:conv_gemm_kernel" -s 63 -c 63 -f -o /tmp/${TAG}_conv_gemm $CMD > gpurun_out/ncu_full_conv.log 2>&1
echo "full rc=$?"; tail -1 gpurun_out/ncu_full_conv.log | cut -c1-120
ncu -i /tmp/${TAG}_conv_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_gemm_raw.csv 2>/dev/null; ls -la gpurun_out/${TAG}_conv_gemm_raw.csv
du -sh gpurun_out
