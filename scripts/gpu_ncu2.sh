#!/bin/bash
# launch list + one --set full capture of kernels matching $KREGEX on the short bench
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profile-reps 1 ${BENCH_ARGS}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c ${LCOUNT:-700} --csv --log-file gpurun_out/${TAG:-x}_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s ${SKIP:-0} -c ${COUNT:-12} -f -o gpurun_out/${TAG:-x}_${OUT:-prof} $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
