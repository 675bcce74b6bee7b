#!/bin/bash
# Round-2 final evidence pass (1 GPU): full GPU parity suite, smoke, the default bench (all legs + secondary), the driver's form,
# the reference arm, the conditional-Unet and bf16w2 lines, the training benches.
TAG=${1:-r02f}
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "=== pytest -m gpu" | tee gpurun_out/${TAG}_pytest.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee -a gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_metrics.jsonl gpurun_out/${TAG}_parity_metrics.jsonl 2>/dev/null
echo "=== smoke" | tee gpurun_out/${TAG}_smoke.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_smoke.log
echo "=== bench (default)"
timeout 900 python bench.py --profile-out gpurun_out/${TAG}_step_profile.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench.json
echo "=== bench --steps 20 --warmup 5"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_k20.json 2> gpurun_out/${TAG}_bench_k20.err
cut -c1-200 gpurun_out/${TAG}_bench_k20.json
echo "=== bench --impl reference --steps 20 --warmup 5"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
for w in unet_cond unet_sr3 hicedrn_cond; do
  timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  cut -c1-200 gpurun_out/${TAG}_bench_$w.json
done
HICDIFF_B200_PRECISION=bf16w2 timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/${TAG}_bench_unet_uncond_bf16w2.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_unet_uncond_bf16w2.json
for m in unet hicedrn; do
  timeout 300 python scripts/bench_train.py --model $m --optim fused --steps 40 --warmup 8 --profile > gpurun_out/${TAG}_train_$m.json 2> gpurun_out/${TAG}_train_$m.err
  cut -c1-200 gpurun_out/${TAG}_train_$m.json
done
