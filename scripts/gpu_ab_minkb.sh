#!/bin/bash
# A/B: minimum K blocks per wgrad CTA (the K-split cap on the low-resolution levels)
mkdir -p gpurun_out
for r in 32 16 8 32 16 8; do
  HD_WGRAD_MINKB=$r timeout 100 python scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 > gpurun_out/mk_$r.json 2> gpurun_out/mk.err
  python -c "import json;d=json.load(open('gpurun_out/mk_$r.json'));print('minkb=$r', round(d['ms_per_step'],3),'ms')" || tail -3 gpurun_out/mk.err
done
