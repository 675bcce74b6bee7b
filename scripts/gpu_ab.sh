#!/bin/bash
# A/B a conv-path switch on the short bench: usage VAR=HD_CONV_PAD VALS="0 1 2" bash scripts/gpu_ab.sh
mkdir -p gpurun_out
for v in $VALS; do
  echo "=== $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline --profile-out gpurun_out/profile_ab_$v.json > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  tail -2 gpurun_out/bench_ab_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_$v.json").read().strip().splitlines()[-1])
print("tiles/s", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "conv TF/s", round(d["roofline"]["achieved"],1))
for k,v in d["roofline"]["families"].items(): print("   ",k,{a:round(b,3) for a,b in v.items()})
PY
done
