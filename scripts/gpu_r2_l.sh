#!/bin/bash
# Round-2 GPU pass L: what the mbarrier try_wait suspend-time hint costs / gains now that the attention epilogues are no longer issue-bound
# (library variants built by scripts/build_variant.sh), and one --set full capture (source view) of each fused attention kernel.
TAG=${1:-r02l}
mkdir -p gpurun_out
for lib in default; do
  for mode in 0 1; do
    if [ $lib = default ]; then unset HICDIFF_B200_LIB; else export HICDIFF_B200_LIB=$PWD/hicdiff_b200/lib/variants/lib$lib.so; fi
    HD_LA_KV=$mode timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_${lib}_kv$mode.json 2> gpurun_out/${TAG}_bench_${lib}_kv$mode.err
    python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_${lib}_kv$mode.json') if l.startswith('{')][-1]);print('$lib HD_LA_KV=$mode', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', {k:v['ms'] for k,v in d['roofline']['families'].items()})" || tail -3 gpurun_out/${TAG}_bench_${lib}_kv$mode.err
  done
done
unset HICDIFF_B200_LIB
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --profile-reps 1"
HD_LA_KV=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:linattn_kv2" -c 1 -f -o gpurun_out/${TAG}_kv2 $CMD > gpurun_out/${TAG}_ncu_kv2.log 2>&1
HD_LA_KV=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:linattn_(kv|out|mix)_kernel" -c 3 -f -o gpurun_out/${TAG}_kv1_mix_out $CMD > gpurun_out/${TAG}_ncu_kv1.log 2>&1
ls -la gpurun_out/${TAG}*.ncu-rep
