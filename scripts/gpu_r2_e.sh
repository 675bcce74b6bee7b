#!/bin/bash
# Round-2 GPU pass E: ONE compute-sanitizer tool per call (B200_PROFILING.md) over the operator-level parity tests at their
# small shapes, then the B = 16 (BASELINE config 1) per-launch profile.   usage: ... 'bash scripts/gpu_r2_e.sh memcheck|racecheck r02'
TOOL=${1:-memcheck}
TAG=${2:-r02}
mkdir -p gpurun_out
SEL="test_conv_gemm or test_upsample_conv or test_downsample_unshuffle or test_groupnorm_film_silu or test_channel_layernorm or test_linear_attention or test_linattn_block_fused or test_full_attention or test_stem_conv"
echo "=== plain run of the selection"
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/${TAG}_sanitizer_plain.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_sanitizer_plain.log
echo "=== compute-sanitizer --tool $TOOL"
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/${TAG}_sanitizer_${TOOL}.log python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/${TAG}_sanitizer_${TOOL}_pytest.log 2>&1
echo "sanitizer rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_${TOOL}_pytest.log; tail -5 gpurun_out/${TAG}_sanitizer_${TOOL}.log
if [ "$TOOL" = "memcheck" ]; then
echo "=== B = 16 profile"
timeout 300 python bench.py --workload unet_cond --batch 16 --steps 300 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_unet_cond_b16.json > gpurun_out/${TAG}_bench_unet_cond_b16.json 2> gpurun_out/${TAG}_bench_unet_cond_b16.err
cut -c1-200 gpurun_out/${TAG}_bench_unet_cond_b16.json
fi
