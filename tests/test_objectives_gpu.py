"""objective = 'pred_x0' / 'pred_v' (model_predictions hicdiff_condition.py:559-579, p_losses :731-740) on the GPU: the objective
only changes the two coefficient columns of the reverse step and the training target.  Golden: 40-step conditional chains and
training losses of the unmodified reference (oracle/make_golden_objectives.py), reference draws injected.
Tolerance: chain final tiles RMS <= 1e-2, loss 5e-3 relative (bf16 eps-net)."""
import pytest
import torch

import helpers
from oracle import hicdiff_oracle as O

pytestmark = pytest.mark.gpu
GOLD = torch.load(helpers.GOLD / "objectives.pt")


@pytest.mark.parametrize("objective", ["pred_x0", "pred_v"])
def test_objective_chain_and_training_loss(objective):
    from hicdiff_b200 import hicdiff_condition as H

    B, T = GOLD["B"], GOLD["T"]
    torch.manual_seed(helpers.MANIFEST["weight_seed"])
    net = H.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True)
    diff = H.GaussianDiffusion(net, image_size=64, timesteps=T, loss_type="l2", beta_schedule="linear", objective=objective).to("cuda")
    clean, noisy = O.synthetic_tiles(B, seed=1234)
    chain_noise = torch.randn(T, B, 1, 64, 64, generator=torch.Generator().manual_seed(2024))
    out = diff.super_resolution(noisy.cuda(), noise=chain_noise.cuda()).cpu()
    ref = GOLD["cases"][objective]["final"]
    rms = float((out - ref).pow(2).mean().sqrt())
    assert rms <= 1e-2, rms
    noise = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(99))
    diff.train()
    loss = diff.p_losses([noisy.cuda(), clean.cuda()], t=GOLD["t"].cuda(), noise=noise.cuda())
    loss.backward()
    want = GOLD["cases"][objective]["loss"]
    assert abs(float(loss.detach()) - want) <= 5e-3 * abs(want), (float(loss.detach()), want)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    # the plan keeps one schedule per objective: a second call must not re-finalize (stable coefficient pointers)
    key = diff.model.eps_plan._schedule_id
    diff.super_resolution(noisy.cuda(), noise=chain_noise.cuda())
    assert diff.model.eps_plan._schedule_id == key
