#!/bin/bash
# fused-Adam check: parity tests (set BENCH=1 to also run the two training benches with each optimiser)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -s -k "fused_adam" -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/adam_test.log
[ -z "$BENCH" ] && exit 0
for m in unet hicedrn; do for o in torch fused; do
  timeout 300 python scripts/bench_train.py --model $m --optim $o --steps 20 --warmup 5 > gpurun_out/train_${m}_${o}.json 2> gpurun_out/train_${m}_${o}.err
  python -c "import json;d=json.load(open('gpurun_out/train_${m}_${o}.json'));print('$m $o', round(d['ms_per_step'],3),'ms', round(d['value'],1),'tiles/s')" || tail -3 gpurun_out/train_${m}_${o}.err
done; done
