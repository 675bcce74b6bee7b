#!/bin/bash
# one ncu --set full capture (few launches of the kernels named in $KREGEX) on a short bench run
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profile-reps 1 ${BENCH_ARGS}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-groupnorm_kernel}" -s ${SKIP:-20} -c ${COUNT:-4} -f -o gpurun_out/${OUT:-prof} $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
