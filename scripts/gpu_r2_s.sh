#!/bin/bash
# Round-2 GPU pass S: ResnetBlock tail as one launch (res_conv epilogue adds SiLU(GroupNorm(block2 conv))) against the two-launch form:
# full parity suite with the fusion on (default), same-box A/B with per-launch profiles, all three Unets.
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
for mode in 0 1 0 1; do
HD_RESBLOCK_TAIL=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_tail$mode.json > gpurun_out/${TAG}_bench_tail$mode.json 2> gpurun_out/${TAG}_bench_tail$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_tail$mode.json') if l.startswith('{')][-1]);print('HD_RESBLOCK_TAIL=$mode sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})
L=json.load(open('gpurun_out/${TAG}_step_profile_tail$mode.json'))
for l in L:
    if l['tag'].startswith('ups.3.0') or l['tag'].startswith('ups.1.0'): print('   ', l['tag'], l['kernel'], round(l['ms']*1e3,1))" || tail -3 gpurun_out/${TAG}_bench_tail$mode.err
done 2>&1 | tee gpurun_out/${TAG}_ab.log
for w in unet_cond unet_sr3; do
  timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  cut -c1-200 gpurun_out/${TAG}_bench_$w.json
done
