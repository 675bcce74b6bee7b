#!/bin/bash
# iteration helper: selected op tests (each -k group in its own process), eps/golden tests, short bench with per-launch profile
mkdir -p gpurun_out
rm -f gpurun_out/iter.log
for t in ${TESTS}; do
  echo "=== $t" | tee -a gpurun_out/iter.log
  timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "$t" -p no:cacheprovider 2>&1 | tail -${TAILN:-25} | tee -a gpurun_out/iter.log
done
if [ -z "$NO_EPS" ]; then
echo "=== eps" | tee -a gpurun_out/iter.log
timeout 900 python -m pytest tests/test_eps_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -${TAILN:-25} | tee -a gpurun_out/iter.log
fi
if [ -z "$NO_BENCH" ]; then
timeout 600 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline --profile-out gpurun_out/profile_iter.json > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
tail -3 gpurun_out/bench_iter.err
python - <<'PY' | tee -a gpurun_out/iter.log
import json
d=json.loads(open("gpurun_out/bench_iter.json").read().strip().splitlines()[-1])
print("tiles/s", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "conv TF/s", round(d["roofline"]["achieved"],1), "step frac", round(d["roofline"]["whole_step"]["frac"],3))
for k,v in d["roofline"]["families"].items(): print("   ",k,{a:round(b,3) for a,b in v.items()})
PY
fi
