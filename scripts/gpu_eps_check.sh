#!/bin/bash
mkdir -p gpurun_out
for v in unet_cond hicedrn_cond; do
  echo "=== layer debug $v" | tee -a gpurun_out/eps_check.log
  timeout 900 python scripts/gpu_layer_debug.py $v 2 2>&1 | tail -70 | tee -a gpurun_out/eps_check.log
done
echo "=== pytest eps" | tee -a gpurun_out/eps_check.log
timeout 1500 python -m pytest tests/test_eps_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -60 | tee -a gpurun_out/eps_check.log
