#!/bin/bash
# Round-2 multi-GPU pass on N GPUs of one box: config 4 (sharded whole-genome pipeline == single rank, bit-identical),
# config 3 (SR3, 4096 tiles, STRONG scaling), config 2 (weak scaling, the driver's form), config 5 (training replicas).
# Usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_r2_multi.sh N r02'
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "=== multi_gpu_check (config 4) N=$N"
timeout 600 $RUN scripts/multi_gpu_check.py > gpurun_out/${TAG}_multi_gpu_check_n${N}.log 2>&1; echo "rc=$?"; grep multi_gpu_check gpurun_out/${TAG}_multi_gpu_check_n${N}.log
echo "=== strong scaling, SR3 4096 tiles (config 3) N=$N"
E2E=""; [ "$N" -lt 4 ] && E2E="--no-e2e"
timeout 900 $RUN bench.py --gpus $N --workload unet_sr3 --scaling strong --total-tiles 4096 --steps 20 --warmup 5 --no-cpu-baseline $E2E > gpurun_out/${TAG}_bench_sr3_strong_n${N}.json 2> gpurun_out/${TAG}_bench_sr3_strong_n${N}.err
tail -2 gpurun_out/${TAG}_bench_sr3_strong_n${N}.err; cut -c1-240 gpurun_out/${TAG}_bench_sr3_strong_n${N}.json
echo "=== weak scaling, unconditional Unet 256 tiles / GPU (config 2) N=$N"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_uncond_weak_n${N}.json 2> gpurun_out/${TAG}_bench_uncond_weak_n${N}.err
tail -2 gpurun_out/${TAG}_bench_uncond_weak_n${N}.err; cut -c1-240 gpurun_out/${TAG}_bench_uncond_weak_n${N}.json
echo "=== training replicas, conditional Unet B=64 / GPU (config 5) N=$N"
timeout 600 $RUN scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 > gpurun_out/${TAG}_train_unet_n${N}.json 2> gpurun_out/${TAG}_train_unet_n${N}.err
tail -2 gpurun_out/${TAG}_train_unet_n${N}.err; cut -c1-300 gpurun_out/${TAG}_train_unet_n${N}.json
