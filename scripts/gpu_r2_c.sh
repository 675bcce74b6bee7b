#!/bin/bash
# Round-2 GPU pass C: training path -- parity suite, then the Unet / hicedrn_Diff training benches (fused Adam), with the
# tensor-core linear-attention backward (default) and the CUDA-core form it replaces (HD_LA_BWD_LEGACY=1) on the same box.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/${TAG}_train_test.log
for leg in 0 1; do
  HD_LA_BWD_LEGACY=$leg timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 --profile > gpurun_out/${TAG}_train_unet_legacy${leg}.json 2> gpurun_out/${TAG}_train_unet_legacy${leg}.err
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_train_unet_legacy${leg}.json'));print('unet legacy=$leg', round(d['ms_per_step'],3),'ms', round(d['value'],1),'tiles/s', d.get('families'))" || tail -3 gpurun_out/${TAG}_train_unet_legacy${leg}.err
done
timeout 300 python scripts/bench_train.py --model hicedrn --optim fused --steps 20 --warmup 5 > gpurun_out/${TAG}_train_hicedrn.json 2> gpurun_out/${TAG}_train_hicedrn.err
python -c "import json;d=json.load(open('gpurun_out/${TAG}_train_hicedrn.json'));print('hicedrn', round(d['ms_per_step'],3),'ms', round(d['value'],1),'tiles/s')" || tail -3 gpurun_out/${TAG}_train_hicedrn.err
