#!/bin/bash
# Round-2 GPU pass U: two-group dx-stacked epilogue with ONE TMA store per tile (main library) against two per-half stores
# (hicdiff_b200/lib/variants/libtwostore.so = the previous commit), same box; then the full suite + smoke + driver-form bench.
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_conv.log
for lib in twostore main twostore main; do
if [ $lib = main ]; then unset HICDIFF_B200_LIB; else export HICDIFF_B200_LIB=$PWD/hicdiff_b200/lib/variants/lib$lib.so; fi
timeout 300 python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-secondary --profile-out gpurun_out/${TAG}_step_profile_$lib.json > gpurun_out/${TAG}_bench_$lib.json 2> gpurun_out/${TAG}_bench_$lib.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$lib.json') if l.startswith('{')][-1]);print('$lib', round(d['ms_per_step'],4),'ms', round(d['value'],2),'tiles/s', 'conv', d['roofline']['families']['conv_gemm']['ms'])
L=json.load(open('gpurun_out/${TAG}_step_profile_$lib.json'))
g=lambda t:[round(l['ms']*1e3,1) for l in L if l['tag']==t][0]
print('   downs.0.0.block1', g('downs.0.0.block1.proj.weight'), 'downs.1.0.block1', g('downs.1.0.block1.proj.weight'), 'ups.3.3', g('ups.3.3.weight'), 'us')" || tail -3 gpurun_out/${TAG}_bench_$lib.err
done 2>&1 | tee gpurun_out/${TAG}_ab.log
unset HICDIFF_B200_LIB
echo "=== pytest -m gpu" | tee gpurun_out/${TAG}_pytest.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee -a gpurun_out/${TAG}_pytest.log
echo "=== smoke" | tee gpurun_out/${TAG}_smoke.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_k20.json 2> gpurun_out/${TAG}_bench_k20.err
cut -c1-200 gpurun_out/${TAG}_bench_k20.json
