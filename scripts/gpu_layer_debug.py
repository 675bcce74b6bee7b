"""Layer-by-layer comparison of the CUDA eps-net against the CPU oracle (debug aid; run on the GPU box).
usage: python scripts/gpu_layer_debug.py [variant] [B]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch

import helpers
from oracle import hicdiff_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "unet_cond"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net, v = helpers.build_net(name)
okw = v["oracle"]
sd = {k: t.clone() for k, t in net.state_dict().items()}
print("state_dict sha ok:", helpers.sd_checksum(sd) == v["state_dict_sha256"])
clean, noisy = O.synthetic_tiles(B, seed=1234)
x_t = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(77))
t = 37
if okw["sr3"]:
    lv = O.sr3_noise_levels(v["schedule"], 1000)
    time = torch.FloatTensor([lv[t + 1]]).repeat(B, 1)
else:
    time = torch.full((B,), t, dtype=torch.long)
cond = noisy if okw["self_condition"] else None
taps = {}
with torch.no_grad():
    ref = helpers.oracle_eps_fn(sd, okw, taps)(x_t, time, cond)
net = net.cuda()
net.eps_plan.debug_keep = True
eps = net(x_t.cuda(), time.cuda(), cond.cuda() if cond is not None else None)
torch.cuda.synchronize()
print(f"eps rel-rms {helpers.rel_rms(eps, ref):.4e}  max abs {float((eps.cpu() - ref).abs().max()):.4e}  ref rms {float(ref.pow(2).mean().sqrt()):.4f}")
names = net.eps_plan.debug_names(B)
for n in names:
    if n in taps:
        got = net.eps_plan.debug_read(B, n)
        if got.shape != taps[n].shape:   # ups.k.2 is stored already 2x-upsampled (fused into the closing LayerNorm)
            got = got[:, :, ::2, ::2]
        print(f"{n:28s} rel-rms {helpers.rel_rms(got, taps[n]):.4e}  shape {tuple(got.shape)}")
