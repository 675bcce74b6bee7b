#!/bin/bash
# Round-2 GPU pass H: PDL inside the training step -- same-box A/B/C with repeats: HD_PDL=0 (plain launches), fwd (the forward-path
# kernels only), all (every kernel of the step).
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
for rep in 1 2 3; do
for pdl in 0 fwd all; do
  HD_PDL=$pdl timeout 300 python scripts/bench_train.py --model unet --optim fused --steps 40 --warmup 8 > gpurun_out/${TAG}_train_unet_pdl${pdl}_$rep.json 2> gpurun_out/${TAG}_train_unet_pdl${pdl}_$rep.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/${TAG}_train_unet_pdl${pdl}_$rep.json') if l.startswith('{')][-1]);print('unet HD_PDL=$pdl', round(d['ms_per_step'],4),'ms')" || tail -3 gpurun_out/${TAG}_train_unet_pdl${pdl}_$rep.err
done
done
for pdl in 0 fwd all; do
  HD_PDL=$pdl timeout 300 python scripts/bench_train.py --model hicedrn --optim fused --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('hicedrn HD_PDL=$pdl', round(d['ms_per_step'],4),'ms')"
done
