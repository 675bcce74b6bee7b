#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_k200.json 2> gpurun_out/bench_k200.err; tail -3 gpurun_out/bench_k200.err
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --workload hicedrn_cond > gpurun_out/bench_hicedrn.json 2> gpurun_out/bench_hicedrn.err; tail -3 gpurun_out/bench_hicedrn.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_k200.json","gpurun_out/bench_hicedrn.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "tiles/s", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "conv TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "step frac", round(d["roofline"]["whole_step"]["frac"],3))
        for k,v in d["roofline"]["families"].items(): print("   ",k,v)
        print("   e2e", d.get("e2e"), "clocks", d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
