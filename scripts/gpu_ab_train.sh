#!/bin/bash
# A/B of two training-step changes on one box (conditional Unet, B = 64, fused Adam): batched weight prep x finisher slices
mkdir -p gpurun_out
for rep in 1 2; do
for pb in 0 1; do for sl in 8 32; do
  HD_PREP_BATCH=$pb HD_SUM_SLICES=$sl timeout 200 python scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 > gpurun_out/ab_${pb}_${sl}.json 2> gpurun_out/ab.err
  python -c "import json;d=json.load(open('gpurun_out/ab_${pb}_${sl}.json'));print('prep_batch=$pb slices=$sl', round(d['ms_per_step'],3),'ms')" || tail -3 gpurun_out/ab.err
done; done; done
