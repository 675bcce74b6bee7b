#!/bin/bash
# Round-2 GPU pass G: programmatic dependent launch (PDL) of the sampling-path kernels -- full parity suite with it on, then
# A/B of the step time on the same box (HD_PDL=0 = plain launches).
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
for pdl in 1 0 1 0; do
  HD_PDL=$pdl timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_pdl${pdl}.json 2> gpurun_out/${TAG}_bench_pdl${pdl}.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_pdl${pdl}.json') if l.startswith('{')][-1]);print('HD_PDL=$pdl', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s')" || tail -3 gpurun_out/${TAG}_bench_pdl${pdl}.err
done
for pdl in 1 0; do
  HD_PDL=$pdl timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 > gpurun_out/${TAG}_train_pdl${pdl}.json 2> gpurun_out/${TAG}_train_pdl${pdl}.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/${TAG}_train_pdl${pdl}.json') if l.startswith('{')][-1]);print('train HD_PDL=$pdl', round(d['ms_per_step'],4),'ms')" || tail -3 gpurun_out/${TAG}_train_pdl${pdl}.err
done
HD_PDL=1 timeout 300 python bench.py --workload unet_cond --batch 16 --steps 300 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('B=16 PDL=1', round(d['ms_per_step'],4))"
HD_PDL=0 timeout 300 python bench.py --workload unet_cond --batch 16 --steps 300 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('B=16 PDL=0', round(d['ms_per_step'],4))"
