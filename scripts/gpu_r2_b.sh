#!/bin/bash
# Round-2 GPU pass B: precision modes (bf16 / bf16w2) -- eps + T = 1000 parity, and the cost of the split-weight mode.
TAG=${1:-r02b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 900 python -m pytest tests/test_t1000_gpu.py tests/test_eps_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity_metrics.jsonl gpurun_out/${TAG}_parity_metrics.jsonl 2>/dev/null
for w in unet_uncond unet_cond; do
HICDIFF_B200_PRECISION=bf16w2 timeout 300 python bench.py --workload $w --steps 100 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_${w}_bf16w2.json 2> gpurun_out/${TAG}_bench_${w}_bf16w2.err
tail -2 gpurun_out/${TAG}_bench_${w}_bf16w2.err; cut -c1-260 gpurun_out/${TAG}_bench_${w}_bf16w2.json
done
