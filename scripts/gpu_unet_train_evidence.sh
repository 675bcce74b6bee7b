#!/bin/bash
# evidence for the Unet training step as it stands: whole-step ncu launch list (eager op order) + ncu --set full of the wgrad
# kernels and the GroupNorm-backward passes.  The plain run goes first (a number printed under ncu is never a bench value).
mkdir -p gpurun_out
export HD_TRAIN_GRAPH=0
CMD="python scripts/bench_train.py --model unet --optim fused --steps 1 --warmup 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/unet_train_launches3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python scripts/ncu_kernel_times.py gpurun_out/unet_train_launches3.csv --last-step > gpurun_out/unet_train_kernels3.txt; tail -2 gpurun_out/unet_train_kernels3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wgrad_rows_kernel|wgrad_general_kernel|gn_bwd_sums_kernel|gn_bwd_dx_kernel" -s 100 -c 16 -f -o gpurun_out/r01w_unet_train $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
