"""Pins the oracle for objective = 'pred_x0' / 'pred_v' (model_predictions hicdiff_condition.py:559-579, p_losses :731-740) against
the UNMODIFIED reference and writes tests/golden/objectives.pt: a 40-step conditional chain and one training loss per objective.

Run in the container that has /root/reference:   python oracle/make_golden_objectives.py"""
from __future__ import annotations

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


def main():
    from src import hicdiff_condition as R_c

    torch.set_num_threads(os.cpu_count() or 8)
    B, T = 2, 40
    clean, noisy = O.synthetic_tiles(B, seed=1234)
    chain_noise = torch.randn(T, B, 1, 64, 64, generator=torch.Generator().manual_seed(2024))
    t = torch.tensor([3, 31], dtype=torch.long)
    noise = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(99))
    out = {"B": B, "T": T, "schedule": "linear", "t": t, "cases": {}}
    for objective in ("pred_x0", "pred_v"):
        torch.manual_seed(0)
        net = R_c.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True).eval()
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        diff = R_c.GaussianDiffusion(net, image_size=64, timesteps=T, loss_type="l2", beta_schedule="linear", objective=objective)
        buf = O.diffusion_buffers("linear", T)
        it = iter(chain_noise)
        o_randn, o_randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: next(it).clone()
        torch.randn_like = lambda *a, **k: next(it).clone()
        try:
            with torch.no_grad():
                ref = diff.super_resolution(noisy)
        finally:
            torch.randn, torch.randn_like = o_randn, o_randn_like
        eps_fn = lambda x, tt, c: O.unet_forward(sd, x, tt, c, self_condition=True)  # noqa: E731
        with torch.no_grad():
            ora = O.p_sample_loop(eps_fn, buf, noisy, chain_noise, timesteps=T, objective=objective)
        assert torch.equal(ref, ora), f"{objective}: oracle chain differs from the reference (max {float((ref - ora).abs().max()):.3e})"
        o_randint = torch.randint
        torch.randint = lambda *a, **k: t.clone()
        torch.randn_like = lambda *a, **k: noise.clone()
        try:
            net.train()
            loss = diff([noisy, clean])
            loss.backward()
        finally:
            torch.randint, torch.randn_like = o_randint, o_randn_like
        o_loss, o_grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type="l2", self_condition=True, net="unet", objective=objective)
        assert torch.equal(o_loss, loss.detach())
        for k, p in net.named_parameters():
            assert torch.equal(p.grad, o_grads[k]), f"{objective}: grad of {k}"
        out["cases"][objective] = {"final": ref.clone(), "loss": float(loss.detach())}
        print(f"{objective}: oracle == reference bit-for-bit ({T}-step chain, loss {float(loss.detach()):.6f} and all gradients)")
    torch.save(out, ROOT / "tests" / "golden" / "objectives.pt")
    print("wrote tests/golden/objectives.pt")


if __name__ == "__main__":
    main()
