#!/bin/bash
# Round-2 GPU pass N: dx-stacked form of the 3x3, Cout = 64 conv (N = 192 MMAs, pixel shifts in the epilogue).
# Parity of the new kind, same-box A/B against the one-MMA-per-tap form, per-launch profile, full suite.
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv" 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_conv.log
for mode in 0 1 0 1; do
HD_CONV_DX3=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_dx$mode.json > gpurun_out/${TAG}_bench_dx$mode.json 2> gpurun_out/${TAG}_bench_dx$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_dx$mode.json') if l.startswith('{')][-1]);print('HD_CONV_DX3=$mode sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})" || tail -3 gpurun_out/${TAG}_bench_dx$mode.err
done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for mode in 0 1; do
HD_CONV_DX3=$mode timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 40 --warmup 8 > gpurun_out/${TAG}_train_unet_dx$mode.json 2> gpurun_out/${TAG}_train_unet_dx$mode.err
cut -c1-200 gpurun_out/${TAG}_train_unet_dx$mode.json
done
