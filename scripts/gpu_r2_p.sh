#!/bin/bash
# Round-2 GPU pass P: dx-stacked conv with two epilogue groups on alternating tiles (HD_CONV_DX3=2, the default) against the
# one-group form (=1) and one MMA per tap (=0): parity, same-box A/B with per-launch profiles, full suite, training A/B.
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv" 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_conv.log
for mode in 0 1 2 0 1 2; do
HD_CONV_DX3=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_dx$mode.json > gpurun_out/${TAG}_bench_dx$mode.json 2> gpurun_out/${TAG}_bench_dx$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_dx$mode.json') if l.startswith('{')][-1]);print('HD_CONV_DX3=$mode sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})
L=json.load(open('gpurun_out/${TAG}_step_profile_dx$mode.json'))
g=lambda t:[round(l['ms']*1e3,1) for l in L if l['tag']==t][0]
print('   downs.0.0.block1', g('downs.0.0.block1.proj.weight'), 'ups.3.0.block1', g('ups.3.0.block1.proj.weight'), 'ups.3.3', g('ups.3.3.weight'), 'us')" || tail -3 gpurun_out/${TAG}_bench_dx$mode.err
done 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for mode in 0 2; do
HD_CONV_DX3=$mode timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 40 --warmup 8 > gpurun_out/${TAG}_train_unet_dx$mode.json 2> gpurun_out/${TAG}_train_unet_dx$mode.err
cut -c1-200 gpurun_out/${TAG}_train_unet_dx$mode.json
done
