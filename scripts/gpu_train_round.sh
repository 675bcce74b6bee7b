#!/bin/bash
# training-path evidence: parity tests, benchmark with per-family profile, ncu launch list + --set full capture of wgrad_kernel
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_train_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/train_test.log
timeout 300 python scripts/bench_train.py --batch 64 --steps 10 --warmup 3 --profile > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
tail -3 gpurun_out/train_bench.err; cat gpurun_out/train_bench.json
if [ -n "$NCU" ]; then
CMD="python scripts/bench_train.py --batch 64 --blocks 4 --steps 1 --warmup 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-r01t}_train_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wgrad_kernel|film_silu" -s 6 -c 6 -f -o gpurun_out/${TAG:-r01t}_wgrad $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
