#!/bin/bash
# Round-2 GPU pass Q: why the two-group dx-stacked epilogue is no faster than one group -- ablations + one --set full capture.
# NOTE: HD_DX3_DIAG was a measurement-only switch (results wrong by construction); it exists in commit 63b567c (one-group form) and was
# removed afterwards -- re-running this script on a later tree measures the unablated kernel seven times.
TAG=${1:-r02q}
mkdir -p gpurun_out
for diag in 0 1 2 8 16 27; do
HD_DX3_DIAG=$diag timeout 200 python bench.py --steps 60 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_diag$diag.json > gpurun_out/${TAG}_bench_diag$diag.json 2> gpurun_out/${TAG}_bench_diag$diag.err
python -c "
import json
L=json.load(open('gpurun_out/${TAG}_step_profile_diag$diag.json'))
g=lambda t:[round(l['ms']*1e3,1) for l in L if l['tag']==t][0]
print('diag $diag', 'downs.0.0.block1', g('downs.0.0.block1.proj.weight'), 'ups.3.0.block1', g('ups.3.0.block1.proj.weight'), 'ups.3.3', g('ups.3.3.weight'), 'us')" || tail -3 gpurun_out/${TAG}_bench_diag$diag.err
done 2>&1 | tee gpurun_out/${TAG}_diag.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_gemm_kernel" -s 126 -c 2 -f -o gpurun_out/${TAG}_dx3g_k576 $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_dx3g_k576.ncu-rep --page raw --csv > gpurun_out/${TAG}_dx3g_k576_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_dx3g_k576.ncu-rep --page source --csv > gpurun_out/${TAG}_dx3g_k576_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_dx3g_k576*
