#!/bin/bash
# Round-2 GPU pass D (profiler): ncu launch lists with DRAM bytes for the unconditional and the conditional Unet (B = 256),
# then ncu --set full captures of one step's conv_gemm launches and of the fused linear-attention kernels (with source).
# Each ncu command follows a plain run of the same command line that exited 0 (B200_PROFILING.md).  gpurun_out/ must stay
# under 64 MiB: the conv report (63 launches, ~100 MB) is reduced to its raw-metric CSV on the box and deleted.
TAG=${1:-r02}
mkdir -p gpurun_out
for w in unet_uncond unet_cond; do
  CMD="python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
  echo "=== ncu launch list $w"
  $CMD > gpurun_out/ncu_plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 420 --csv --log-file gpurun_out/${TAG}_launches_$w.csv $CMD > gpurun_out/ncu_launch_$w.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/ncu_launch_$w.log | cut -c1-120
done
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
echo "=== ncu --set full conv_gemm (one step)"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none -k "regex:conv_gemm_kernel" -s 63 -c 63 -f -o /tmp/${TAG}_conv_gemm $CMD > gpurun_out/ncu_full_conv.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/ncu_full_conv.log | cut -c1-120
ncu -i /tmp/${TAG}_conv_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_gemm_raw.csv 2>/dev/null; ls -la gpurun_out/${TAG}_conv_gemm_raw.csv
echo "=== ncu --set full linattn (one step)"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:linattn_(kv|mix|out)_kernel" -s 15 -c 6 -f -o gpurun_out/${TAG}_linattn $CMD > gpurun_out/ncu_full_linattn.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/ncu_full_linattn.log | cut -c1-120
ncu -i gpurun_out/${TAG}_linattn.ncu-rep --page raw --csv > gpurun_out/${TAG}_linattn_raw.csv 2>/dev/null
du -sh gpurun_out
