#!/bin/bash
# Round-2 GPU pass J: pipelined linattn_kv2_kernel -- its parity cases, then a same-box A/B against the first form (HD_LA_KV=1).
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "linattn or linear_attention" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_linattn.log
for mode in 0 1 0 1; do
HD_LA_KV=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_kv$mode.json > gpurun_out/${TAG}_bench_kv$mode.json 2> gpurun_out/${TAG}_bench_kv$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_kv$mode.json') if l.startswith('{')][-1]);print('HD_LA_KV=$mode sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})" || tail -3 gpurun_out/${TAG}_bench_kv$mode.err
done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
