#!/bin/bash
# Full evidence pass for one round: GPU parity suite, smoke, default bench (all legs), reference arm, ncu launch list,
# one ncu --set full capture of the conv GEMM.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh rNN'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu" | tee gpurun_out/${TAG}_pytest.log
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 | tee -a gpurun_out/${TAG}_pytest.log
echo "=== smoke" | tee gpurun_out/${TAG}_smoke.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_smoke.log
echo "=== bench (default)"
timeout 900 python bench.py --profile-out gpurun_out/${TAG}_step_profile.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json | cut -c1-600
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cat gpurun_out/${TAG}_bench_reference.json | cut -c1-400
if [ -z "$NO_NCU" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profile-reps 1"
echo "=== ncu launch list"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c ${LCOUNT:-450} --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
echo "=== ncu --set full (conv_gemm)"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-conv_gemm_kernel}" -s ${SKIP:-0} -c ${COUNT:-40} -f -o gpurun_out/${TAG}_conv_gemm $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
