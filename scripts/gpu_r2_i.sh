#!/bin/bash
# Round-2 GPU pass I: mbarrier try_wait with a suspend-time hint -- full parity suite, sampling and training benches, ubench.
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile.json > gpurun_out/${TAG}_bench_$rep.json 2> gpurun_out/${TAG}_bench_$rep.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$rep.json') if l.startswith('{')][-1]);print('sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})" || tail -3 gpurun_out/${TAG}_bench_$rep.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('K=20', round(d['ms_per_step'],4), round(d['value'],2), 'e2e', d['e2e']['value'])"
timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 40 --warmup 8 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('train unet', round(d['ms_per_step'],4))"
timeout 300 python scripts/bench_train.py --model hicedrn --optim fused --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('train hicedrn', round(d['ms_per_step'],4))"
timeout 300 python bench.py --workload hicedrn_cond --steps 50 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('hicedrn sampling', round(d['ms_per_step'],4), round(d['value'],3), d['roofline']['achieved'])"
