"""Pins the oracle's DDRM restatement against the UNMODIFIED reference sampler and writes tests/golden/ddrm_uncond.pt.

Run in the container that has /root/reference:   python oracle/make_golden_ddrm.py
  reference: src/functions/denoising.py::efficient_generalized_steps with src/functions/svd_replacement.py::Denoising,
  driven as src/Utils/metrics_diff.py:36-81,215-224 does (linear betas 1e-4..2e-2, T = 1000, 20 strided steps),
  eps-net = the reference's unconditional Unet under torch seed 0.
torch.randn_like is patched to a seeded generator and every draw is recorded; the fixture stores the draw each step USES."""
from __future__ import annotations

import os
import sys
from pathlib import Path

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import hicdiff_oracle as O  # noqa: E402


def main():
    from src import hicdiff as R_u
    from src.functions.denoising import efficient_generalized_steps as ref_steps
    from src.functions.svd_replacement import Denoising

    torch.set_num_threads(os.cpu_count() or 8)
    torch.manual_seed(0)
    net = R_u.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=False).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    B, T, steps = 2, 1000, 20
    betas = torch.from_numpy(np.linspace(0.0001, 0.02, T, dtype=np.float64)).float()      # metrics_diff.py:36-81 ('linear')
    seq = list(range(0, T, T // steps))
    sigma_0, etaB, etaA, etaC = 0.2, 1.0, 0.85, 0.85                                         # metrics_diff.py:220
    _, noisy = O.synthetic_tiles(B, seed=1234)
    x_init = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(77))
    H = Denoising(1, 64, torch.device("cpu"))

    g = torch.Generator().manual_seed(2025)
    draws = []
    orig = torch.randn_like

    def rec(t, *a, **k):
        z = torch.randn(t.shape, generator=g, dtype=t.dtype)
        draws.append(z)
        return z

    torch.randn_like = rec
    try:
        with torch.no_grad():
            xs_ref, x0_ref = ref_steps(x_init.clone(), seq, net, betas, H, noisy.clone(), sigma_0, etaB=etaB, etaA=etaA, etaC=etaC)
    finally:
        torch.randn_like = orig
    assert len(draws) == 3 * steps, len(draws)
    used = []
    for k in range(steps):
        a, b, c = draws[3 * k:3 * k + 3]
        used.append((c if b.numel() == 0 else b).reshape(B, 1, 64, 64).clone())     # 'before' uses the third draw, 'after' the second
    modes = ["before" if draws[3 * k + 1].numel() == 0 else "after" for k in range(steps)]
    assert "before" in modes and "after" in modes, modes

    eps_fn = lambda x, t: O.unet_forward(sd, x, t, None, self_condition=False)  # noqa: E731
    with torch.no_grad():
        xs_o, x0_o = O.ddrm_denoising_steps(eps_fn, betas, x_init.clone(), seq, noisy.clone(), sigma_0, etaB, etaA, etaC, used)
    for k, (a, b) in enumerate(zip(xs_ref, xs_o)):
        assert torch.equal(a, b), f"oracle DDRM xs[{k}] differs from the reference (max {float((a - b).abs().max()):.3e})"
    for k, (a, b) in enumerate(zip(x0_ref, x0_o)):
        assert torch.equal(a, b), f"oracle DDRM x0_preds[{k}] differs from the reference"
    out = ROOT / "tests" / "golden" / "ddrm_uncond.pt"
    torch.save({"seq": seq, "sigma_0": sigma_0, "etaB": etaB, "etaA": etaA, "etaC": etaC, "tile_seed": 1234, "x_init": x_init,
                "noise": torch.stack(used), "modes": modes, "final": xs_ref[-1].clone(), "x0_last": x0_ref[-1].clone(),
                "xs": torch.stack(xs_ref), "betas": betas}, out)
    print(f"oracle == reference bit-for-bit over {steps} DDRM steps ({modes.count('before')} before / {modes.count('after')} after); wrote {out}")


if __name__ == "__main__":
    main()
