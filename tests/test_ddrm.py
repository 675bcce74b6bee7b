"""DDRM sampler: oracle restatement vs the reference fixture (CPU, bit-exact), host-side step
coefficients, and the CUDA path (hd_eps_forward + hd_ddrm_step through `efficient_generalized_steps`) vs the same fixture.

Fixture: tests/golden/ddrm_uncond.pt, written by oracle/make_golden_ddrm.py from the UNMODIFIED reference
(src/functions/denoising.py with the Denoising operator, unconditional Unet under seed 0, linear betas, 20 strided steps,
sigma_0 = 0.2 so that both the "noisier than y" and the "less noisy than y" branches occur)."""
import pytest
import torch

import helpers
from oracle import hicdiff_oracle as O

GOLD = helpers.GOLD / "ddrm_uncond.pt"


def test_oracle_ddrm_reproduces_reference_fixture_bit_exact():
    gold = torch.load(GOLD)
    net, v = helpers.build_net("unet_uncond")
    sd = net.state_dict()
    _, noisy = O.synthetic_tiles(gold["x_init"].shape[0], seed=gold["tile_seed"])
    eps_fn = lambda x, t: O.unet_forward(sd, x, t, None, self_condition=False)  # noqa: E731
    with torch.no_grad():
        xs, x0s = O.ddrm_denoising_steps(eps_fn, gold["betas"], gold["x_init"].clone(), gold["seq"], noisy, gold["sigma_0"],
                                         gold["etaB"], gold["etaA"], gold["etaC"], list(gold["noise"]))
    assert len(xs) == len(gold["seq"]) + 1
    assert torch.equal(torch.stack(xs), gold["xs"])
    assert torch.equal(x0s[-1], gold["x0_last"])


def test_ddrm_step_coefficients_follow_the_reference_branches():
    """The host picks ONE of the reference's three masked cases per step (all singular values are 1)."""
    from hicdiff_b200.functions.denoising import _step_scalars

    gold = torch.load(GOLD)
    seq = gold["seq"]
    seq_next = [-1] + seq[:-1]
    modes = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        mode, sq_at, sq_1m_at, sq_at_next, c0, c1, c2 = _step_scalars(gold["betas"], i, j, gold["sigma_0"], gold["etaB"],
                                                                      gold["etaA"], gold["etaC"])
        modes.append("before" if mode == 0 else "after")
        assert abs(sq_at ** 2 + sq_1m_at ** 2 - 1.0) < 1e-6
        if mode == 0:
            assert c0 == gold["etaB"] and c1 == 1 - gold["etaB"] and c2 >= 0
    assert modes == gold["modes"]
    assert _step_scalars(gold["betas"], seq[0], -1, gold["sigma_0"], 1.0, 0.85, 0.85)[3] == 1.0   # a_{-1} = 1: the last step lands on x0


def test_ddrm_rejects_other_operators_and_cpu_tensors():
    from hicdiff_b200.functions.denoising import efficient_generalized_steps
    from hicdiff_b200.functions.svd_replacement import Denoising

    gold = torch.load(GOLD)
    H = Denoising(1, 64, torch.device("cpu"))
    with pytest.raises(RuntimeError):          # no CPU fallback
        efficient_generalized_steps(gold["x_init"], gold["seq"], None, gold["betas"], H, gold["x_init"], 0.2, 1.0, 0.85, 0.85)


@pytest.mark.gpu
def test_ddrm_sampler_matches_reference_fixture():
    from hicdiff_b200.functions.denoising import efficient_generalized_steps
    from hicdiff_b200.functions.svd_replacement import Denoising

    gold = torch.load(GOLD)
    net, v = helpers.build_net("unet_uncond")
    net = net.cuda()
    B = gold["x_init"].shape[0]
    _, noisy = O.synthetic_tiles(B, seed=gold["tile_seed"])
    H = Denoising(1, 64, torch.device("cuda"))
    xs, x0s = efficient_generalized_steps(gold["x_init"].cuda(), gold["seq"], net, gold["betas"].cuda(), H, noisy.cuda(),
                                          gold["sigma_0"], gold["etaB"], gold["etaA"], gold["etaC"], noise=list(gold["noise"]))
    assert len(xs) == len(gold["seq"]) + 1 and len(x0s) == len(gold["seq"])
    out = xs[-1].cpu()
    assert torch.isfinite(out).all()
    # tolerances as for the DDPM chains (bf16 eps-net, fp32 update): final-tile RMS <= 1e-2
    rms = float((out - gold["final"]).pow(2).mean().sqrt())
    rms0 = float((x0s[-1].cpu() - gold["x0_last"]).pow(2).mean().sqrt())
    assert rms <= 1e-2, f"DDRM final RMS {rms:.3e}"
    assert rms0 <= 1e-2, f"DDRM x0 RMS {rms0:.3e}"
    # first step is teacher-forced (same x_T): only the eps-net error enters
    first = float((xs[1].cpu() - gold["xs"][1]).pow(2).mean().sqrt() / gold["xs"][1].pow(2).mean().sqrt())
    assert first <= 2e-2, f"DDRM first step rel-RMS {first:.3e}"
    # Philox mode: reproducible under torch.manual_seed, different across seeds
    torch.manual_seed(3)
    a = efficient_generalized_steps(gold["x_init"].cuda(), gold["seq"][::4], net, gold["betas"].cuda(), H, noisy.cuda(), 0.2, 1.0, 0.85, 0.85)[0][-1]
    torch.manual_seed(3)
    b = efficient_generalized_steps(gold["x_init"].cuda(), gold["seq"][::4], net, gold["betas"].cuda(), H, noisy.cuda(), 0.2, 1.0, 0.85, 0.85)[0][-1]
    assert torch.equal(a, b) and torch.isfinite(a).all()
