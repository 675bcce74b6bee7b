#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --profile-out gpurun_out/profile_unet_uncond_b256.json > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profile-reps 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
