#!/bin/bash
# Round-2 pass X: 2-stream GroupNorm apply with 8 instead of 4 chunks in flight per thread (HD_GN_CH8=0/1), full suite, bench line.
TAG=${1:-r02x}
mkdir -p gpurun_out
for mode in 0 1 0 1; do
HD_GN_CH8=$mode timeout 300 python bench.py --steps 150 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_ch8_$mode.json > gpurun_out/${TAG}_bench_ch8_$mode.json 2> gpurun_out/${TAG}_bench_ch8_$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_ch8_$mode.json') if l.startswith('{')][-1]);print('HD_GN_CH8=$mode', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', 'groupnorm', d['roofline']['families']['groupnorm_film_silu']['ms'])
L=json.load(open('gpurun_out/${TAG}_step_profile_ch8_$mode.json'))
for l in L:
    if l['tag'] in ('downs.0.0.block1.norm','ups.2.0.block1.norm','downs.2.0.block1.norm','mid_block1.block1.norm'): print('   ', l['tag'], round(l['ms']*1e3,1), 'us', round(l['bytes']/l['ms']/1e6), 'GB/s')" || tail -3 gpurun_out/${TAG}_bench_ch8_$mode.err
done 2>&1 | tee gpurun_out/${TAG}_ab.log
echo "=== pytest -m gpu" | tee gpurun_out/${TAG}_pytest.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee -a gpurun_out/${TAG}_pytest.log
echo "=== smoke" | tee gpurun_out/${TAG}_smoke.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_k20.json 2> gpurun_out/${TAG}_bench_k20.err
cut -c1-200 gpurun_out/${TAG}_bench_k20.json
