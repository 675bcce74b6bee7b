#!/bin/bash
# quick iteration: operator + eps parity, then the sampling bench with per-family times
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_eps_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile.json 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items() if v['ms']>0.05})"
done
