"""Pins the oracle's training-step restatement (p_losses_and_grads) against the UNMODIFIED reference and writes
tests/golden/hicedrn_train.json.

Run in the container that has /root/reference:   python oracle/make_golden_train.py
  reference: loss = GaussianDiffusion(hicedrn_Diff(...))([noisy, clean]); loss.backward()   (train.py:84-129;
  src/hicdiff_condition.py:715-750, src/hicdiff.py:711-755, src/model/hicedrn_Diff.py:267-289)
torch.randint / torch.randn_like are patched so both sides see the same t and noise.  Full gradients of even a 2-block net are
~14 MB, so the fixture keeps, per parameter, (sum, L2 norm, 32 entries at fixed strided indices) -- enough to pin the oracle
on any box; the GPU tests compare the device gradients with the oracle's FULL gradients computed live."""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from oracle import hicdiff_oracle as O  # noqa: E402

CASES = [
    dict(name="cond_l2", flavour="cond", self_condition=True, loss_type="l2", schedule="linear", B=2, blocks=2, T=1000, t=[17, 803]),
    dict(name="uncond_l1", flavour="uncond", self_condition=False, loss_type="l1", schedule="linear", B=2, blocks=2, T=1000, t=[0, 999]),
]


def summary(g: torch.Tensor):
    flat = g.reshape(-1).double()
    n = flat.numel()
    idx = torch.linspace(0, n - 1, min(32, n)).long()
    return {"sum": float(flat.sum()), "norm": float(flat.norm()), "idx": idx.tolist(), "val": [float(v) for v in g.reshape(-1)[idx]]}


def main():
    from src import hicdiff as R_u
    from src import hicdiff_condition as R_c
    from src.model.hicedrn_Diff import hicedrn_Diff

    torch.set_num_threads(os.cpu_count() or 8)
    out = {"weight_seed": 0, "tile_seed": 1234, "noise_seed": 99, "cases": {}}
    for c in CASES:
        torch.manual_seed(0)
        net = hicedrn_Diff(number_resnet=c["blocks"], self_condition=c["self_condition"])
        G = R_c.GaussianDiffusion if c["flavour"] == "cond" else R_u.GaussianDiffusion
        diff = G(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"], auto_normalize=False)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        clean, noisy = O.synthetic_tiles(c["B"], seed=1234)
        t = torch.tensor(c["t"], dtype=torch.long)
        noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(99))
        o_randint, o_randn_like = torch.randint, torch.randn_like
        torch.randint = lambda *a, **k: t.clone()
        torch.randn_like = lambda *a, **k: noise.clone()
        try:
            loss = diff([noisy, clean]) if c["flavour"] == "cond" else diff(clean)
            loss.backward()
        finally:
            torch.randint, torch.randn_like = o_randint, o_randn_like
        ref_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        buf = O.diffusion_buffers(c["schedule"], c["T"])
        o_loss, o_grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type=c["loss_type"],
                                               self_condition=c["self_condition"], num_blocks=c["blocks"])
        assert torch.equal(o_loss, loss.detach()), (float(o_loss), float(loss))
        assert ref_grads.keys() == o_grads.keys()
        for k in ref_grads:
            assert torch.equal(ref_grads[k], o_grads[k]), f"{c['name']}: oracle grad of {k} differs from the reference"
        out["cases"][c["name"]] = {**{k: v for k, v in c.items() if k != "name"}, "loss": float(loss.detach()),
                                   "grads": {k: summary(g) for k, g in ref_grads.items()}}
        print(f"{c['name']}: oracle == reference bit-for-bit (loss {float(loss):.6f}, {len(ref_grads)} gradients)")
    # ---- SR3 flavour over hicedrn_sr3_Diff (pretrain/train_hicedrn_Diff_sr3.py): numpy's global RNG picks t and the level
    from src import hicdiff_sr3 as R_s
    from src.model.hicedrn_sr3_Diff import hicedrn_Diff as hicedrn_sr3

    c = dict(flavour="sr3", self_condition=True, loss_type="l2", schedule="linear", B=2, blocks=2, T=1000, np_seed=5)
    torch.manual_seed(0)
    net = hicedrn_sr3(number_resnet=c["blocks"], self_condition=True)
    diff = R_s.GaussianDiffusion(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"], auto_normalize=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    clean, noisy = O.synthetic_tiles(c["B"], seed=1234)
    noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(99))
    import numpy as np
    np.random.seed(c["np_seed"])
    loss = diff.p_losses([noisy, clean], noise=noise.clone())
    loss.backward()
    np.random.seed(c["np_seed"])                                   # replay the two draws of p_losses :754-762
    t = np.random.randint(1, c["T"] + 1)
    lv_table = O.sr3_noise_levels(c["schedule"], c["T"])
    level = torch.FloatTensor(np.random.uniform(lv_table[t - 1], lv_table[t], size=c["B"]))
    ref_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    o_loss, o_grads = O.sr3_p_losses_and_grads(sd, noisy, clean, level, noise, loss_type=c["loss_type"], self_condition=True, num_blocks=c["blocks"])
    assert torch.equal(o_loss, loss.detach()), (float(o_loss), float(loss))
    for k in ref_grads:
        assert torch.equal(ref_grads[k], o_grads[k]), f"sr3: oracle grad of {k} differs from the reference"
    out["cases"]["sr3_l2"] = {**c, "t": int(t), "level": [float(v) for v in level], "loss": float(loss.detach()),
                              "grads": {k: summary(g) for k, g in ref_grads.items()}}
    print(f"sr3_l2: oracle == reference bit-for-bit (t = {t}, loss {float(loss.detach()):.6f}, {len(ref_grads)} gradients)")
    # ---- SR3 Unet (pretrain/train_unet_Diff_sr3.py)
    c = dict(flavour="sr3", self_condition=True, loss_type="l2", schedule="linear", B=2, T=1000, np_seed=6)
    torch.manual_seed(0)
    net = R_s.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True, noise_level_emb=True)
    diff = R_s.GaussianDiffusion(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"], auto_normalize=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    np.random.seed(c["np_seed"])
    loss = diff.p_losses([noisy, clean], noise=noise.clone())
    loss.backward()
    np.random.seed(c["np_seed"])
    t = np.random.randint(1, c["T"] + 1)
    level = torch.FloatTensor(np.random.uniform(lv_table[t - 1], lv_table[t], size=c["B"]))
    ref_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    o_loss, o_grads = O.sr3_p_losses_and_grads(sd, noisy, clean, level, noise, loss_type=c["loss_type"], self_condition=True, net="unet")
    assert torch.equal(o_loss, loss.detach()), (float(o_loss), float(loss))
    for k in ref_grads:
        assert torch.equal(ref_grads[k], o_grads[k]), f"unet sr3: oracle grad of {k} differs from the reference"
    out["cases"]["unet_sr3_l2"] = {**c, "t": int(t), "level": [float(v) for v in level], "loss": float(loss.detach()),
                                   "grads": {k: summary(g) for k, g in ref_grads.items()}}
    print(f"unet_sr3_l2: oracle == reference bit-for-bit (t = {t}, loss {float(loss.detach()):.6f}, {len(ref_grads)} gradients)")

    # ---- the Unet eps-net (pretrain/train_unet_Diff_cond*.py, train_unet_uncond.py)
    for c in (dict(name="unet_cond_l2", flavour="cond", self_condition=True, loss_type="l2", schedule="sigmoid", B=2, T=1000, t=[17, 803]),
              dict(name="unet_uncond_l1", flavour="uncond", self_condition=False, loss_type="l1", schedule="linear", B=2, T=1000, t=[17, 803])):
        R = R_c if c["flavour"] == "cond" else R_u
        torch.manual_seed(0)
        net = R.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=c["self_condition"])
        diff = R.GaussianDiffusion(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"])
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        clean, noisy = O.synthetic_tiles(c["B"], seed=1234)
        t = torch.tensor(c["t"], dtype=torch.long)
        noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(99))
        o_randint, o_randn_like = torch.randint, torch.randn_like
        torch.randint = lambda *a, **k: t.clone()
        torch.randn_like = lambda *a, **k: noise.clone()
        try:
            loss = diff([noisy, clean]) if c["flavour"] == "cond" else diff(clean)
            loss.backward()
        finally:
            torch.randint, torch.randn_like = o_randint, o_randn_like
        ref_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        o_loss, o_grads = O.p_losses_and_grads(sd, O.diffusion_buffers(c["schedule"], c["T"]), noisy, clean, t, noise,
                                               loss_type=c["loss_type"], self_condition=c["self_condition"], net="unet")
        assert torch.equal(o_loss, loss.detach()), (float(o_loss), float(loss.detach()))
        assert ref_grads.keys() == o_grads.keys()
        for k in ref_grads:
            assert torch.equal(ref_grads[k], o_grads[k]), f"{c['name']}: oracle grad of {k} differs from the reference"
        out["cases"][c["name"]] = {**{k: v for k, v in c.items() if k != "name"}, "loss": float(loss.detach()),
                                   "grads": {k: summary(g) for k, g in ref_grads.items()}}
        print(f"{c['name']}: oracle == reference bit-for-bit (loss {float(loss.detach()):.6f}, {len(ref_grads)} gradients)")
    path = ROOT / "tests" / "golden" / "hicedrn_train.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
