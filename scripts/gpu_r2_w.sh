#!/bin/bash
# Round-2 pass W (final binary): the default bench line (full 1000-step chain, secondary results: conditional Unet B = 256 and B = 16),
# the conditional / SR3 lines, the Unet training step.
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 600 python bench.py --profile-out gpurun_out/${TAG}_step_profile.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench.json
for w in unet_cond unet_sr3; do
  timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  cut -c1-200 gpurun_out/${TAG}_bench_$w.json
done
timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 40 --warmup 8 --profile > gpurun_out/${TAG}_train_unet.json 2> gpurun_out/${TAG}_train_unet.err
cut -c1-200 gpurun_out/${TAG}_train_unet.json
