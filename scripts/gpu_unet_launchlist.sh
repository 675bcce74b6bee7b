#!/bin/bash
# whole-step ncu launch list of one eager Unet training step (the graph's contents): warm-up step + one step, last step summarised
mkdir -p gpurun_out
export HD_TRAIN_GRAPH=0
CMD="python scripts/bench_train.py --model unet --optim fused --steps 1 --warmup 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/unet_train_launches2.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
python scripts/ncu_kernel_times.py gpurun_out/unet_train_launches2.csv --last-step > gpurun_out/unet_train_kernels2.txt; tail -5 gpurun_out/unet_train_kernels2.txt
