#!/bin/bash
# quick training check: parity suite, then both training benches (OPTIM=torch|fused, default fused)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/train_test.log
for m in unet hicedrn; do
  timeout 300 python scripts/bench_train.py --model $m --optim ${OPTIM:-fused} --steps 20 --warmup 5 > gpurun_out/train_${m}.json 2> gpurun_out/train_${m}.err
  python -c "import json;d=json.load(open('gpurun_out/train_${m}.json'));print('$m', round(d['ms_per_step'],3),'ms', round(d['value'],1),'tiles/s')" || tail -3 gpurun_out/train_${m}.err
done
