#!/bin/bash
# Round-2 GPU pass M: kv kernels with the LayerNorm applied in place before the MMAs (no per-pixel constants in the epilogue).
# Parity for both kv forms, same-box A/B, per-kernel durations, one --set full capture each, full suite.
TAG=${1:-r02m}
mkdir -p gpurun_out
for mode in 0 1; do
HD_LA_KV=$mode timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "linattn or linear_attention" 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_linattn_kv$mode.log
done
for mode in 0 1 0 1; do
HD_LA_KV=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_kv$mode.json > gpurun_out/${TAG}_bench_kv$mode.json 2> gpurun_out/${TAG}_bench_kv$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_kv$mode.json') if l.startswith('{')][-1]);print('HD_LA_KV=$mode sampling', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})" || tail -3 gpurun_out/${TAG}_bench_kv$mode.err
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
for mode in 0 1; do
HD_LA_KV=$mode timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:linattn" -c 60 --csv --log-file gpurun_out/${TAG}_linattn_launches_kv$mode.csv $CMD > gpurun_out/${TAG}_ncu_launch_kv$mode.log 2>&1
done
HD_LA_KV=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:linattn_kv2" -c 1 -f -o gpurun_out/${TAG}_kv2 $CMD > gpurun_out/${TAG}_ncu_kv2.log 2>&1

timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
