#!/bin/bash
# Round-2 pass V (2 GPUs, final binary): sharded == single rank (config 4), weak-scaling bench line (config 2), training replicas (config 5).
N=${1:-2}
TAG=${2:-r02v}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $RUN scripts/multi_gpu_check.py > gpurun_out/${TAG}_multi_gpu_check_n${N}.log 2>&1; echo "rc=$?"; grep multi_gpu_check gpurun_out/${TAG}_multi_gpu_check_n${N}.log
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_uncond_weak_n${N}.json 2> gpurun_out/${TAG}_bench_uncond_weak_n${N}.err
tail -2 gpurun_out/${TAG}_bench_uncond_weak_n${N}.err; cut -c1-240 gpurun_out/${TAG}_bench_uncond_weak_n${N}.json
timeout 300 $RUN scripts/bench_train.py --model unet --optim fused --steps 30 --warmup 5 > gpurun_out/${TAG}_train_unet_n${N}.json 2> gpurun_out/${TAG}_train_unet_n${N}.err
tail -2 gpurun_out/${TAG}_train_unet_n${N}.err; cut -c1-300 gpurun_out/${TAG}_train_unet_n${N}.json
