"""Pins the oracle's DDIM restatement (ddim_sample) against the UNMODIFIED reference and writes tests/golden/ddim_uncond.pt.

Run in the container that has /root/reference:   python oracle/make_golden_ddim.py
  reference: GaussianDiffusion(Unet(self_condition=False), sampling_timesteps=10, ddim_sampling_eta in {0, 0.5}).sample(x)
  (src/hicdiff.py:623-673).  torch.randn / torch.randn_like are patched to pop injected draws."""
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
    from src import hicdiff as R_u

    torch.set_num_threads(os.cpu_count() or 8)
    torch.manual_seed(0)
    net = R_u.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=False).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    B, T, S = 2, 1000, 10
    out = {"B": B, "T": T, "S": S, "schedule": "linear", "noise_seed": 31, "cases": {}}
    noise = torch.randn(S, B, 1, 64, 64, generator=torch.Generator().manual_seed(31))       # x_T + one z per step but the last
    for eta in (0.0, 0.5):
        diff = R_u.GaussianDiffusion(net, image_size=64, timesteps=T, sampling_timesteps=S, loss_type="l2", beta_schedule="linear",
                                     ddim_sampling_eta=eta)
        it = iter(noise)
        o_randn, o_randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: next(it).clone()
        torch.randn_like = lambda *a, **k: next(it).clone()
        try:
            with torch.no_grad():
                ref = diff.sample(torch.zeros(B, 1, 64, 64), return_all_timesteps=True)
        finally:
            torch.randn, torch.randn_like = o_randn, o_randn_like
        eps_fn = lambda x, t, c: O.unet_forward(sd, x, t, None, self_condition=False)  # noqa: E731
        with torch.no_grad():
            ora = O.ddim_sample(eps_fn, O.diffusion_buffers("linear", T), noise, timesteps=T, sampling_timesteps=S, eta=eta, return_all=True)
        assert ref.shape == ora.shape == (B, S + 1, 1, 64, 64), (ref.shape, ora.shape)
        assert torch.equal(ref, ora), f"oracle DDIM (eta = {eta}) differs from the reference (max {float((ref - ora).abs().max()):.3e})"
        out["cases"][str(eta)] = {"final": ref[:, -1].clone(), "trace": ref.clone()}
        print(f"eta = {eta}: oracle == reference bit-for-bit over {S} DDIM steps")
    torch.save(out, ROOT / "tests" / "golden" / "ddim_uncond.pt")
    print("wrote tests/golden/ddim_uncond.pt")


if __name__ == "__main__":
    main()
