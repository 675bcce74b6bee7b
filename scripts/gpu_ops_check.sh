#!/bin/bash
# Runs each per-kernel parity group in its own process (a trap in one kernel must not poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_conv_gemm test_downsample test_groupnorm test_channel_layernorm test_linear_attention test_full_attention test_stem_conv test_philox; do
  echo "=== $t" | tee -a gpurun_out/ops_check.log
  timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "$t" -p no:cacheprovider 2>&1 | tail -40 | tee -a gpurun_out/ops_check.log
done
