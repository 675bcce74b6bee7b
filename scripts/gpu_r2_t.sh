#!/bin/bash
# Round-2 GPU pass T (profiler, final binary): ncu launch list with DRAM bytes (unconditional Unet, B = 256) and one --set full
# pass over one step's 63 conv_gemm launches, reduced to its raw-metric CSV on the box (the report is ~100 MB).
TAG=${1:-r02t}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 420 --csv --log-file gpurun_out/${TAG}_launches_unet_uncond.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/ncu_launch.log | cut -c1-120
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none -k "regex:conv_gemm_kernel" -s 63 -c 63 -f -o /tmp/${TAG}_conv_gemm $CMD > gpurun_out/ncu_full_conv.log 2>&1
echo "full rc=$?"; tail -1 gpurun_out/ncu_full_conv.log | cut -c1-120
ncu -i /tmp/${TAG}_conv_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_gemm_raw.csv 2>/dev/null; ls -la gpurun_out/${TAG}_conv_gemm_raw.csv
du -sh gpurun_out
# A/B: the 2-stream GroupNorm pass writing a second buffer instead of rewriting its input
for mode in 0 1 0 1; do
HD_GN_OUT_OF_PLACE=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_gnoop$mode.json > gpurun_out/${TAG}_bench_gnoop$mode.json 2> gpurun_out/${TAG}_bench_gnoop$mode.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_gnoop$mode.json') if l.startswith('{')][-1]);print('HD_GN_OUT_OF_PLACE=$mode', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})
L=json.load(open('gpurun_out/${TAG}_step_profile_gnoop$mode.json'))
for l in L:
    if l['tag'] in ('downs.0.0.block1.norm','ups.3.0.block1.norm','ups.2.0.block1.norm'): print('   ', l['tag'], round(l['ms']*1e3,1), 'us', round(l['bytes']/l['ms']/1e6), 'GB/s')" || tail -3 gpurun_out/${TAG}_bench_gnoop$mode.err
done 2>&1 | tee gpurun_out/${TAG}_gnoop_ab.log
# one --set full capture WITH source of a two-group dx-stacked launch (K = 576, first conv of a step)
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_gemm_kernel" -s 126 -c 1 -f -o gpurun_out/${TAG}_dx3g_k576 $CMD > gpurun_out/${TAG}_ncu_g.log 2>&1
ncu -i gpurun_out/${TAG}_dx3g_k576.ncu-rep --page source --csv > gpurun_out/${TAG}_dx3g_k576_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_dx3g_k576.ncu-rep --page raw --csv > gpurun_out/${TAG}_dx3g_k576_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_dx3g_k576*
