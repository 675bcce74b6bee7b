#!/bin/bash
# Round-2 GPU pass A (no profiler): tcgen05 issue-rate microbenchmark, GPU parity suite (incl. the T = 1000 chains), smoke,
# default bench (all legs + secondary results), reference arm, batch-size sweep.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r2_a.sh r02a'
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== ubench tcgen05"
timeout 120 scripts/bin/ubench_tcgen05 4096 > gpurun_out/${TAG}_ubench_tcgen05.jsonl 2> gpurun_out/${TAG}_ubench.err
echo "rc=$? lines=$(wc -l < gpurun_out/${TAG}_ubench_tcgen05.jsonl)"; tail -2 gpurun_out/${TAG}_ubench.err
echo "=== pytest -m gpu" | tee gpurun_out/${TAG}_pytest.log
rm -f gpurun_out/parity_metrics.jsonl
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee -a gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_metrics.jsonl gpurun_out/${TAG}_parity_metrics.jsonl 2>/dev/null
echo "=== smoke" | tee gpurun_out/${TAG}_smoke.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_smoke.log
echo "=== bench (default)"
timeout 900 python bench.py --profile-out gpurun_out/${TAG}_step_profile.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json
echo "=== bench --steps 20 --warmup 5 (the driver's form)"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_k20.json 2> gpurun_out/${TAG}_bench_k20.err
tail -2 gpurun_out/${TAG}_bench_k20.err; cut -c1-300 gpurun_out/${TAG}_bench_k20.json
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-400 gpurun_out/${TAG}_bench_reference.json
for b in 128 64; do
  echo "=== bench --batch $b"
  timeout 300 python bench.py --batch $b --steps 200 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_b${b}.json 2> gpurun_out/${TAG}_bench_b${b}.err
  cut -c1-260 gpurun_out/${TAG}_bench_b${b}.json
done
echo "=== bench unet_cond"
timeout 300 python bench.py --workload unet_cond --steps 200 --no-e2e --no-cpu-baseline --profile-out gpurun_out/${TAG}_step_profile_unet_cond.json > gpurun_out/${TAG}_bench_unet_cond.json 2> gpurun_out/${TAG}_bench_unet_cond.err
cut -c1-260 gpurun_out/${TAG}_bench_unet_cond.json
echo "=== strong scaling form at N=1 (SR3, 4096 tiles)"
timeout 600 python bench.py --workload unet_sr3 --scaling strong --total-tiles 4096 --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_sr3_strong_n1.json 2> gpurun_out/${TAG}_bench_sr3_strong_n1.err
tail -2 gpurun_out/${TAG}_bench_sr3_strong_n1.err; cut -c1-260 gpurun_out/${TAG}_bench_sr3_strong_n1.json
