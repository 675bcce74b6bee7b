"""DDIM sampling (ddim_sample, src/hicdiff.py:623-673) on the GPU: hd_eps_forward + hd_ddim_step per strided step, against the
golden chain of the unmodified reference (oracle/make_golden_ddim.py) with the reference's exact draws injected.
Tolerances: the step given eps is bit-exact; TEACHER-FORCED (each step started from the reference's x_k) the next state is within
RMS 1e-2 (measured 1e-3 .. 4e-3: the bf16 eps-net); the free-running 10-step chain amplifies that per-step error -- with
random-init weights the deterministic eta = 0 map expands by ~1.5x per 100-timestep stride (measured 0.003 -> 0.09), while the
injected noise of eta = 0.5 damps it (0.010) -- so the chain is held to 2e-2 at eta = 0.5 and only to 0.15 at eta = 0."""
import pytest
import torch

import helpers
from oracle import hicdiff_oracle as O

GOLD = torch.load(helpers.GOLD / "ddim_uncond.pt")


def test_oracle_ddim_matches_golden_cpu():
    """(CPU) the oracle still reproduces the reference chain stored in the fixture -- first two steps, eta = 0.5."""
    from hicdiff_b200 import hicdiff

    torch.manual_seed(helpers.MANIFEST["weight_seed"])
    net = hicdiff.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    noise = torch.randn(GOLD["S"], GOLD["B"], 1, 64, 64, generator=torch.Generator().manual_seed(GOLD["noise_seed"]))
    buf = O.diffusion_buffers("linear", GOLD["T"])
    # two steps of the chain are enough to pin the update rule on the CPU (each eps-net call is ~1 s)
    x = noise[0].clone()
    t = torch.full((GOLD["B"],), GOLD["T"] - 1, dtype=torch.long)
    with torch.no_grad():
        eps = O.unet_forward(sd, x, t, None, self_condition=False)
    x0 = (buf["sqrt_recip_alphas_cumprod"][t].view(-1, 1, 1, 1) * x - buf["sqrt_recipm1_alphas_cumprod"][t].view(-1, 1, 1, 1) * eps).clamp(-1, 1)
    times = list(reversed(torch.linspace(-1, GOLD["T"] - 1, steps=GOLD["S"] + 1).int().tolist()))
    a, an = buf["alphas_cumprod"][times[0]], buf["alphas_cumprod"][times[1]]
    sg = 0.5 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    want = x0 * an.sqrt() + (1 - an - sg ** 2).sqrt() * eps + sg * noise[1]
    assert torch.equal(want, GOLD["cases"]["0.5"]["trace"][:, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("eta", [0.0, 0.5])
def test_ddim_chain_matches_reference_golden(eta):
    from hicdiff_b200 import hicdiff

    torch.manual_seed(helpers.MANIFEST["weight_seed"])
    net = hicdiff.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=False)
    diff = hicdiff.GaussianDiffusion(net, image_size=64, timesteps=GOLD["T"], sampling_timesteps=GOLD["S"], loss_type="l2",
                                     beta_schedule="linear", ddim_sampling_eta=eta).to("cuda")
    assert diff.is_ddim_sampling
    noise = torch.randn(GOLD["S"], GOLD["B"], 1, 64, 64, generator=torch.Generator().manual_seed(GOLD["noise_seed"]))
    out = diff.sample(torch.zeros(GOLD["B"], 1, 64, 64, device="cuda"), return_all_timesteps=True, noise=noise)
    ref = GOLD["cases"][str(eta)]["trace"]
    assert out.shape == ref.shape
    final = float((out[:, -1].cpu() - GOLD["cases"][str(eta)]["final"]).pow(2).mean().sqrt())
    assert final <= (2e-2 if eta > 0 else 0.15), final
    # teacher-forced: one step from each of the reference's states
    from hicdiff_b200 import _lib

    times = list(reversed(torch.linspace(-1, GOLD["T"] - 1, steps=GOLD["S"] + 1).int().tolist()))
    ac, sr, srm1 = diff.alphas_cumprod.cpu(), diff.sqrt_recip_alphas_cumprod.cpu(), diff.sqrt_recipm1_alphas_cumprod.cpu()
    for k in range(GOLD["S"]):
        x = ref[:, k].cuda().contiguous().clone()
        t, tn = times[k], times[k + 1]
        eps = net(x, torch.full((GOLD["B"],), t, device="cuda"), None)
        last = tn < 0
        san = c = sg = 0.0
        if not last:
            a, an = ac[t], ac[tn]
            s_ = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
            c, san, sg = float((1 - an - s_ ** 2).sqrt()), float(an.sqrt()), float(s_)
        z = noise[k + 1].cuda().contiguous() if not last else None
        _lib.check(_lib.load().hd_ddim_step(x.data_ptr(), eps.data_ptr(), _lib.ptr(z), None, float(sr[t]), float(srm1[t]), san, c, sg,
                                            1 if last else 0, x.numel(), 0, 0, k, _lib.stream_ptr()), "hd_ddim_step")
        step_rms = float((x.cpu() - ref[:, k + 1]).pow(2).mean().sqrt())
        assert step_rms <= 1e-2, (k, step_rms)
    assert float(out[:, -1].abs().max()) <= 1.0                                     # the last step returns the clipped x_start
    # Philox path: reproducible under torch.manual_seed, different across seeds
    torch.manual_seed(5)
    a = diff.sample(torch.zeros(2, 1, 64, 64, device="cuda"))
    torch.manual_seed(5)
    b = diff.sample(torch.zeros(2, 1, 64, 64, device="cuda"))
    assert torch.equal(a, b) and a.shape == (2, 1, 64, 64)


@pytest.mark.gpu
def test_ddim_step_given_eps_is_bit_exact():
    from hicdiff_b200 import _lib

    g = torch.Generator().manual_seed(2)
    x, eps, z = (torch.randn(3, 1, 64, 64, generator=g) for _ in range(3))
    sr, srm1, san, c, sg = 1.37, 0.93, 0.81, 0.55, 0.21
    x0 = (torch.tensor(sr) * x - torch.tensor(srm1) * eps).clamp(-1, 1)
    want = x0 * torch.tensor(san) + torch.tensor(c) * eps + torch.tensor(sg) * z
    xd, ed, zd = x.cuda().clone(), eps.cuda(), z.cuda()          # keep the device copies alive across the call
    _lib.check(_lib.load().hd_ddim_step(xd.data_ptr(), ed.data_ptr(), zd.data_ptr(), None, sr, srm1, san, c, sg, 0,
                                        xd.numel(), 0, 0, 0, _lib.stream_ptr()), "hd_ddim_step")
    assert torch.equal(xd.cpu(), want)
    xd = x.cuda().clone()
    _lib.check(_lib.load().hd_ddim_step(xd.data_ptr(), ed.data_ptr(), None, None, sr, srm1, 0.0, 0.0, 0.0, 1, xd.numel(), 0, 0, 0,
                                        _lib.stream_ptr()), "hd_ddim_step")
    assert torch.equal(xd.cpu(), x0)
