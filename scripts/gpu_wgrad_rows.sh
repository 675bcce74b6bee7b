#!/bin/bash
# filter-row wgrad: operator parity first, then the whole training suite and the Unet bench with the form on and off
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py -x -q -k "general_conv_wgrad" -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/wgrad_rows_test.log
grep -q "failed\|error" gpurun_out/wgrad_rows_test.log && exit 1
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
for r in 64 32 64 32; do
  HD_WGRAD_ROWS=$r timeout 200 python scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 > gpurun_out/wr_$r.json 2> gpurun_out/wr.err
  python -c "import json;d=json.load(open('gpurun_out/wr_$r.json'));print('rows=$r', round(d['ms_per_step'],3),'ms', round(d['value'],1))" || tail -3 gpurun_out/wr.err
done
