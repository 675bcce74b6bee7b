#!/bin/bash
# Round-2 GPU pass F (N GPUs): bucketed / overlapped gradient all-reduce -- equality with the flat all-reduce, then the Unet
# training bench in both modes.   usage: gpurun --gpus N -- 'bash scripts/gpu_r2_f.sh N r02'
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 300 python -m pytest tests/test_train_gpu.py -q -x -p no:cacheprovider -k "unet" 2>&1 | tail -3
timeout 600 $RUN scripts/multi_gpu_train_check.py > gpurun_out/${TAG}_multi_gpu_train_check_n${N}.log 2>&1; echo "rc=$?"; grep -h "multi_gpu_train_check\|Error\|error" gpurun_out/${TAG}_multi_gpu_train_check_n${N}.log | tail -5
for mode in flat overlap; do
  timeout 600 $RUN scripts/bench_train.py --model unet --optim fused --steps 40 --warmup 5 --allreduce $mode > gpurun_out/${TAG}_train_unet_n${N}_${mode}.json 2> gpurun_out/${TAG}_train_unet_n${N}_${mode}.err
  grep "^{" gpurun_out/${TAG}_train_unet_n${N}_${mode}.json | cut -c1-200; tail -2 gpurun_out/${TAG}_train_unet_n${N}_${mode}.err | cut -c1-200
done
