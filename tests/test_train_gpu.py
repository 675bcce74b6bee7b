"""Training step on the GPU (hd_trainer_*) against the oracle's torch.autograd restatement of what
train.py:120-129 computes (the oracle itself is pinned bit-for-bit to the unmodified reference by oracle/make_golden_train.py
-> tests/golden/hicedrn_train.json, re-checked on the CPU in test_oracle_cpu.py).

Tolerances (stated): activations and the GEMM operands of dgrad / wgrad are bf16 with fp32 accumulation, parameters and all
reductions fp32, so per-parameter gradients are compared by relative RMS:  rel-RMS(grad - oracle) <= 3e-2 (measured 3e-3 .. 1.5e-2),
the loss to 5e-3 relative.  The single-op conv gradients see only the bf16 output rounding / fp32 order: <= 4e-3."""
import json

import pytest
import torch
import torch.nn.functional as F

import helpers
from oracle import hicdiff_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = json.loads((helpers.GOLD / "hicedrn_train.json").read_text())


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def test_conv3x3_wgrad_and_dgrad_single_ops():
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(5)
    B = 3
    x = torch.randn(B, 64, 64, 256, generator=g).to(torch.bfloat16)
    dy = (torch.randn(B, 64, 64, 256, generator=g) * 0.1).to(torch.bfloat16)
    w = torch.randn(256, 256, 3, 3, generator=g) * 0.02
    xr = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    wr = w.to(torch.bfloat16).double().requires_grad_(True)
    y = F.conv2d(xr, wr, padding=1)
    gx, gw = torch.autograd.grad(y, [xr, wr], dy.permute(0, 3, 1, 2).double())
    dw = ops.conv3x3_wgrad_nhwc(x.to(DEV), dy.to(DEV))
    assert torch.isfinite(dw).all()
    assert _rel(dw, gw) <= 4e-3, _rel(dw, gw)
    # every tap individually (a wrong shift / flip would hide in the aggregate only if taps were symmetric)
    for t in range(9):
        assert _rel(dw[:, :, t // 3, t % 3], gw[:, :, t // 3, t % 3]) <= 4e-3, t
    dx = ops.conv3x3_dgrad_nhwc(dy.to(DEV), w.to(DEV))
    assert _rel(dx.permute(0, 3, 1, 2), gx) <= 4e-3, _rel(dx.permute(0, 3, 1, 2), gx)


def _case(name):
    c = GOLD["cases"][name]
    from hicdiff_b200 import hicdiff, hicdiff_condition
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    torch.manual_seed(GOLD["weight_seed"])
    net = hicedrn_Diff(number_resnet=c["blocks"], self_condition=c["self_condition"])
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    G = hicdiff_condition.GaussianDiffusion if c["flavour"] == "cond" else hicdiff.GaussianDiffusion
    diff = G(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"], auto_normalize=False)
    clean, noisy = O.synthetic_tiles(c["B"], seed=GOLD["tile_seed"])
    t = torch.tensor(c["t"], dtype=torch.long)
    noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(GOLD["noise_seed"]))
    return c, net, sd, diff, clean, noisy, t, noise


@pytest.mark.parametrize("name", ["cond_l2", "uncond_l1"])
def test_loss_backward_matches_oracle(name):
    c, net, sd, diff, clean, noisy, t, noise = _case(name)
    buf = O.diffusion_buffers(c["schedule"], c["T"])
    o_loss, o_grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type=c["loss_type"],
                                           self_condition=c["self_condition"], num_blocks=c["blocks"])
    assert abs(float(o_loss) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"]))       # the oracle still equals the pinned reference run
    diff = diff.to(DEV)
    diff.train()
    if c["flavour"] == "cond":
        loss = diff.p_losses([noisy.to(DEV), clean.to(DEV)], t=t.to(DEV), noise=noise.to(DEV))
    else:
        loss = diff.p_losses(clean.to(DEV), t.to(DEV), noise=noise.to(DEV))
    assert loss.requires_grad
    loss.backward()                                                                  # train.py:128
    assert abs(float(loss.detach()) - float(o_loss)) <= 5e-3 * abs(float(o_loss)), (float(loss.detach()), float(o_loss))
    worst = {}
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        assert torch.isfinite(p.grad).all(), k
        worst[k] = _rel(p.grad, o_grads[k])
    bad = {k: v for k, v in worst.items() if v > 3e-2}
    assert not bad, f"{name}: gradient rel-RMS above 3e-2: {bad}"
    print(f"{name}: loss {float(loss.detach()):.6f} (oracle {float(o_loss):.6f}); worst grad rel-RMS {max(worst.values()):.3e} "
          f"({max(worst, key=worst.get)})")


def test_sr3_loss_backward_matches_oracle():
    """pretrain/train_hicedrn_Diff_sr3.py: hicedrn_sr3_Diff (additive noise-level embedding) under the SR3 GaussianDiffusion."""
    from hicdiff_b200.hicdiff_sr3 import GaussianDiffusion
    from hicdiff_b200.model.hicedrn_sr3_Diff import hicedrn_Diff as hicedrn_sr3

    c = GOLD["cases"]["sr3_l2"]
    torch.manual_seed(GOLD["weight_seed"])
    net = hicedrn_sr3(number_resnet=c["blocks"], self_condition=True)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    clean, noisy = O.synthetic_tiles(c["B"], seed=GOLD["tile_seed"])
    noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(GOLD["noise_seed"]))
    level = torch.tensor(c["level"], dtype=torch.float32)
    o_loss, o_grads = O.sr3_p_losses_and_grads(sd, noisy, clean, level, noise, loss_type=c["loss_type"], self_condition=True,
                                               num_blocks=c["blocks"])
    assert abs(float(o_loss) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"]))
    diff = GaussianDiffusion(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"],
                             auto_normalize=False).to(DEV)
    diff.train()
    loss = diff.p_losses([noisy.to(DEV), clean.to(DEV)], noise=noise.to(DEV), level=level.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - float(o_loss)) <= 5e-3 * abs(float(o_loss)), (float(loss.detach()), float(o_loss))
    worst = {k: _rel(p.grad, o_grads[k]) for k, p in net.named_parameters()}
    bad = {k: v for k, v in worst.items() if v > 3e-2}
    assert not bad, f"sr3: gradient rel-RMS above 3e-2: {bad}"
    # the reference's own call: numpy's global RNG picks t and the level (hicdiff_sr3.py:754-762)
    import numpy as np

    np.random.seed(c["np_seed"])
    net.zero_grad()
    l2 = diff([noisy.to(DEV), clean.to(DEV)], noise=noise.to(DEV))
    l2.backward()
    assert abs(float(l2.detach()) - float(loss.detach())) <= 1e-6                  # same seed -> same t / level -> same step
    print(f"sr3_l2: loss {float(loss.detach()):.6f} (oracle {float(o_loss):.6f}); worst grad rel-RMS {max(worst.values()):.3e}")


def test_train_loop_runs_unchanged_and_is_deterministic():
    """train.py:109-136 verbatim: Adam over diffusion.parameters(), loss = diffusion(x); loss.backward(); step; zero_grad."""
    c, net, sd, diff, clean, noisy, t, noise = _case("cond_l2")
    diff = diff.to(DEV)
    diff.train()
    opt = torch.optim.Adam(diffusion_params := list(diff.parameters()), lr=2e-4)
    x = [noisy.to(DEV), clean.to(DEV)]
    losses = []
    for it in range(8):
        loss = diff.p_losses(x, t=t.to(DEV), noise=noise.to(DEV))
        loss.backward()
        if it == 0:
            g0 = {k: p.grad.clone() for k, p in net.named_parameters()}
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))), losses
    assert losses[-1] < 0.9 * losses[0], losses                                     # same batch, same t / noise: the loss must fall
    assert len(diffusion_params) == len(list(net.parameters()))
    # random t / noise path of forward() (what train.py calls), gradients present and finite
    loss = diff(x)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    # determinism: the same weights and inputs give bit-identical gradients (fixed-order reductions, no atomics)
    net2 = type(net)(number_resnet=c["blocks"], self_condition=c["self_condition"])
    net2.load_state_dict(sd)
    diff2 = type(diff)(net2, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"]).to(DEV)
    l2 = diff2.p_losses(x, t=t.to(DEV), noise=noise.to(DEV))
    l2.backward()
    for k, p in net2.named_parameters():
        assert torch.equal(p.grad, g0[k]), k
    # sampling still works with the trained weights (the plan re-reads the updated parameters)
    with torch.no_grad():
        eps = net(x[1], t.to(DEV), x[0])
    assert torch.isfinite(eps).all()


def test_eval_path_needs_no_trainer():

    c, net, sd, diff, clean, noisy, t, noise = _case("cond_l2")
    diff = diff.to(DEV)
    with torch.no_grad():                                                            # validation loop: value only, no trainer
        v = diff.p_losses([noisy.to(DEV), clean.to(DEV)], t=t.to(DEV), noise=noise.to(DEV))
    assert not v.requires_grad and abs(float(v) - c["loss"]) <= 2e-2 * c["loss"]


WGRAD_CASES = [
    # B, H, Cin, Cout, k      (the Unet's levels: 64x64 .. 8x8, 1x1 and 3x3, Cout = 64 rides on a zero-filled 128-row tile)
    (2, 64, 64, 64, 3), (2, 64, 128, 64, 3), (2, 32, 128, 128, 3), (3, 16, 256, 256, 3), (4, 8, 512, 512, 3),
    (5, 64, 64, 128, 3), (3, 64, 192, 64, 3),          # filter-row form (W = 64): Cout tile 128, three Cin tiles, uneven K splits
    (2, 8, 256, 512, 3), (2, 64, 64, 384, 1), (2, 32, 128, 64, 1), (3, 16, 256, 128, 1), (2, 8, 512, 256, 1),
]


@pytest.mark.parametrize("B,H,Cin,Cout,k", WGRAD_CASES)
def test_general_conv_wgrad(B, H, Cin, Cout, k):
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + k)
    x = torch.randn(B, H, H, Cin, generator=g).to(torch.bfloat16)
    dy = (torch.randn(B, H, H, Cout, generator=g) * 0.1).to(torch.bfloat16)
    wr = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), wr, padding=k // 2)
    (gw,) = torch.autograd.grad(y, [wr], dy.permute(0, 3, 1, 2).double())
    dw = ops.conv_wgrad_nhwc(x.to(DEV), dy.to(DEV), k)
    assert torch.isfinite(dw).all()
    assert _rel(dw, gw) <= 4e-3, _rel(dw, gw)
    for t in range(k * k):
        assert _rel(dw[:, :, t // k, t % k], gw[:, :, t // k, t % k]) <= 4e-3, t


def test_general_conv_wgrad_concat_slices():
    """torch.cat((x, skip), 1) feeding a conv: the weight gradient is filled one operand at a time."""
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(11)
    B, H, C0, C1, Cout = 2, 32, 128, 64, 128
    x0 = torch.randn(B, H, H, C0, generator=g).to(torch.bfloat16)
    x1 = torch.randn(B, H, H, C1, generator=g).to(torch.bfloat16)
    dy = (torch.randn(B, H, H, Cout, generator=g) * 0.1).to(torch.bfloat16)
    wr = torch.zeros(Cout, C0 + C1, 3, 3, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(torch.cat((x0, x1), -1).permute(0, 3, 1, 2).double(), wr, padding=1)
    (gw,) = torch.autograd.grad(y, [wr], dy.permute(0, 3, 1, 2).double())
    dw = ops.conv_wgrad_nhwc(x0.to(DEV), dy.to(DEV), 3, cin_total=C0 + C1, ci0=0)
    dw = ops.conv_wgrad_nhwc(x1.to(DEV), dy.to(DEV), 3, dw=dw, cin_total=C0 + C1, ci0=C0)
    assert _rel(dw, gw) <= 4e-3, _rel(dw, gw)


# ------------------------------------------------------------------------------------------------ Unet backward building blocks
@pytest.mark.parametrize("B,H,C,film", [(2, 64, 64, True), (2, 32, 128, True), (3, 16, 256, False), (2, 8, 512, True), (1, 64, 64, False)])
def test_groupnorm_film_silu_backward(B, H, C, film):
    """Block's tail (hicdiff_condition.py:159-171) against torch.autograd in float64 on the same bf16-rounded inputs."""
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(B + H + C)
    y = (torch.randn(B, H, H, C, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    ds = (torch.randn(B, H, H, C, generator=g) * 0.1).to(torch.bfloat16)
    gamma = torch.randn(C, generator=g) * 0.5 + 1.0
    beta = torch.randn(C, generator=g) * 0.2
    scale = torch.randn(B, C, generator=g) * 0.3 if film else None
    shift = torch.randn(B, C, generator=g) * 0.3 if film else None
    leaves = [y.permute(0, 3, 1, 2).double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)]
    n = F.group_norm(leaves[0], 8, leaves[1], leaves[2], eps=1e-5)
    if film:
        leaves += [scale.double().requires_grad_(True), shift.double().requires_grad_(True)]
        n = n * (leaves[3][:, :, None, None] + 1) + leaves[4][:, :, None, None]
    s = F.silu(n)
    grads = torch.autograd.grad(s, leaves, ds.permute(0, 3, 1, 2).double())
    dy, dgamma, dbeta, dscale, dshift, dcb = ops.groupnorm_silu_bwd_nhwc(
        y.to(DEV), ds.to(DEV), gamma.to(DEV), beta.to(DEV), scale.to(DEV) if film else None, shift.to(DEV) if film else None,
        want_conv_bias=True)
    # the bias gradient of the conv that produced y = sum over (b, pixels) of dy; values are O(1e-3) sums of cancelling terms,
    # so compare on the scale of |dy| summed in quadrature
    ref_cb = grads[0].sum(dim=(0, 2, 3))
    scale_cb = float(grads[0].pow(2).sum(dim=(0, 2, 3)).sqrt().mean())
    assert float((dcb.cpu().double() - ref_cb).abs().max()) <= 2e-3 * scale_cb + 1e-7, float((dcb.cpu().double() - ref_cb).abs().max())
    assert _rel(dy.permute(0, 3, 1, 2), grads[0]) <= 6e-3, _rel(dy.permute(0, 3, 1, 2), grads[0])     # bf16 output rounding
    assert _rel(dgamma, grads[1]) <= 1e-4 and _rel(dbeta, grads[2]) <= 1e-4
    if film:
        assert _rel(dscale, grads[3]) <= 1e-4 and _rel(dshift, grads[4]) <= 1e-4


@pytest.mark.parametrize("M,C", [(4096 * 2, 64), (1024 * 3, 128), (256 * 2, 256), (64 * 5, 512)])
def test_channel_layernorm_backward(M, C):
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(torch.bfloat16)
    dz = (torch.randn(M, C, generator=g) * 0.1).to(torch.bfloat16)
    gain = torch.randn(C, generator=g) * 0.3 + 1.0
    xr, gr = x.double().requires_grad_(True), gain.double().requires_grad_(True)
    var = xr.var(dim=1, unbiased=False, keepdim=True)
    z = (xr - xr.mean(dim=1, keepdim=True)) * (var + 1e-5).rsqrt() * gr          # LayerNorm.forward :104-108
    gx, gg = torch.autograd.grad(z, [xr, gr], dz.double())
    dx, dg = ops.channel_layernorm_bwd_nhwc(x.to(DEV), dz.to(DEV), gain.to(DEV))
    assert _rel(dx, gx) <= 6e-3, _rel(dx, gx)
    assert _rel(dg, gg) <= 1e-4, _rel(dg, gg)


@pytest.mark.parametrize("Cout,Cin,k", [(64, 64, 3), (128, 192, 3), (512, 768, 3)])
def test_weight_standardize_backward(Cout, Cin, k):
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(Cout + Cin)
    w = torch.randn(Cout, Cin, k, k, generator=g) * 0.05
    dwt = torch.randn(Cout, Cin, k, k, generator=g)
    wr = w.double().requires_grad_(True)
    mean = wr.mean(dim=(1, 2, 3), keepdim=True)
    var = wr.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    wt = (wr - mean) * (var + 1e-5).rsqrt()                                     # WeightStandardizedConv2d.forward :89-95
    (gw,) = torch.autograd.grad(wt, [wr], dwt.double())
    dw = ops.weight_standardize_bwd(w.to(DEV), dwt.to(DEV))
    assert _rel(dw, gw) <= 1e-4, _rel(dw, gw)


def _attn_ref(qkv, linear):
    """qkv [B, n, 384] float64 -> out [B, n, 128]: LinearAttention / Attention cores (hicdiff_condition.py:212-227, 239-251)."""
    B, n, _ = qkv.shape
    q, k, v = (t.reshape(B, n, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=2))      # 'b h d n'
    scale = 32 ** -0.5
    if linear:
        q = q.softmax(dim=-2) * scale
        k = k.softmax(dim=-1)
        v = v / n
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q)                                      # 'b h e n'
        return out.permute(0, 3, 1, 2).reshape(B, n, 128)
    sim = torch.einsum("bhdi,bhdj->bhij", q * scale, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)                                          # 'b h i d'
    return out.permute(0, 2, 1, 3).reshape(B, n, 128)


@pytest.mark.parametrize("B,n,linear", [(2, 4096, True), (3, 1024, True), (2, 256, True), (5, 64, True), (3, 64, False)])
def test_attention_core_backward(B, n, linear):
    from hicdiff_b200 import ops

    g = torch.Generator().manual_seed(B * 7 + n)
    qkv = (torch.randn(B, n, 384, generator=g) * 1.2).to(torch.bfloat16)
    dout = (torch.randn(B, n, 128, generator=g) * 0.1).to(torch.bfloat16)
    x = qkv.double().requires_grad_(True)
    (gq,) = torch.autograd.grad(_attn_ref(x, linear), [x], dout.double())
    dqkv = ops.attention_bwd(qkv.to(DEV), dout.to(DEV), linear=linear)
    for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
        r = _rel(dqkv[:, :, sl], gq[:, :, sl])
        assert r <= 6e-3, (name, r)


# ------------------------------------------------------------------------------------------------ the Unet training step
def _unet_case(self_condition, loss_type, schedule, seed=0):
    from hicdiff_b200 import hicdiff, hicdiff_condition

    mod = hicdiff_condition if self_condition else hicdiff
    torch.manual_seed(seed)
    net = mod.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=self_condition)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    diff = mod.GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type=loss_type, beta_schedule=schedule)
    B = 2
    clean, noisy = O.synthetic_tiles(B, seed=1234)
    t = torch.tensor([17, 803], dtype=torch.long)
    noise = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(99))
    return net, sd, diff, clean, noisy, t, noise


@pytest.mark.parametrize("self_condition,loss_type,schedule", [(True, "l2", "sigmoid"), (False, "l1", "linear")])
def test_unet_loss_backward_matches_oracle(self_condition, loss_type, schedule):
    """pretrain/train_unet_Diff_cond*.py / train_unet_uncond.py: loss = diffusion(x); loss.backward() on the Unet eps-net.
    Tolerance: per-parameter gradient rel-RMS <= 5e-2 (bf16 activations through ~75 convs, 38 GroupNorms and 9 attention blocks;
    measured worst case printed), loss 5e-3."""
    net, sd, diff, clean, noisy, t, noise = _unet_case(self_condition, loss_type, schedule)
    buf = O.diffusion_buffers(schedule, 1000)
    o_loss, o_grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type=loss_type, self_condition=self_condition, net="unet")
    gold = GOLD["cases"]["unet_cond_l2" if self_condition else "unet_uncond_l1"]          # the reference's own loss.backward()
    assert abs(float(o_loss) - gold["loss"]) <= 1e-6 * max(1.0, abs(gold["loss"]))
    diff = diff.to(DEV)
    diff.train()
    if self_condition:
        loss = diff.p_losses([noisy.to(DEV), clean.to(DEV)], t=t.to(DEV), noise=noise.to(DEV))
    else:
        loss = diff.p_losses(clean.to(DEV), t.to(DEV), noise=noise.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - float(o_loss)) <= 5e-3 * abs(float(o_loss)), (float(loss.detach()), float(o_loss))
    worst = {}
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        worst[k] = _rel(p.grad, o_grads[k])
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:6]
    print(f"unet(self_condition={self_condition}, {loss_type}): loss {float(loss.detach()):.6f} (oracle {float(o_loss):.6f}); worst grad rel-RMS: {top}")
    bad = {k: v for k, v in worst.items() if v > 5e-2}
    assert not bad, f"gradient rel-RMS above 5e-2: {bad}"


def test_unet_train_loop_runs_unchanged():
    """The loop of pretrain/train_unet_Diff_cond.py: Adam over diffusion.parameters(); same batch / t / noise -> the loss must fall;
    deterministic gradients; the sampling plan sees the updated weights."""
    net, sd, diff, clean, noisy, t, noise = _unet_case(True, "l2", "sigmoid")
    diff = diff.to(DEV)
    diff.train()
    opt = torch.optim.Adam(diff.parameters(), lr=1e-4)
    x = [noisy.to(DEV), clean.to(DEV)]
    losses = []
    for it in range(6):
        loss = diff.p_losses(x, t=t.to(DEV), noise=noise.to(DEV))
        loss.backward()
        if it == 0:
            g0 = {k: p.grad.clone() for k, p in net.named_parameters()}
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert all(l == l for l in losses) and losses[-1] < 0.9 * losses[0], losses
    loss = diff(x)                                                                   # random t / noise, what the scripts call
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    net2, _, diff2, *_ = _unet_case(True, "l2", "sigmoid")
    diff2 = diff2.to(DEV)
    l2 = diff2.p_losses(x, t=t.to(DEV), noise=noise.to(DEV))
    l2.backward()
    for k, p in net2.named_parameters():
        assert torch.equal(p.grad, g0[k]), k
    with torch.no_grad():
        eps = net(x[1], t.to(DEV), x[0])
    assert torch.isfinite(eps).all()


def test_unet_sr3_loss_backward_matches_oracle():
    """pretrain/train_unet_Diff_sr3.py: Unet(noise_level_emb=True) under the SR3 GaussianDiffusion (additive noise-level
    embedding after block1, numpy-RNG level sampling)."""
    from hicdiff_b200 import hicdiff_sr3

    c = GOLD["cases"]["unet_sr3_l2"]
    torch.manual_seed(GOLD["weight_seed"])
    net = hicdiff_sr3.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True, noise_level_emb=True)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    clean, noisy = O.synthetic_tiles(c["B"], seed=GOLD["tile_seed"])
    noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(GOLD["noise_seed"]))
    level = torch.tensor(c["level"], dtype=torch.float32)
    o_loss, o_grads = O.sr3_p_losses_and_grads(sd, noisy, clean, level, noise, loss_type=c["loss_type"], self_condition=True, net="unet")
    assert abs(float(o_loss) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"]))
    diff = hicdiff_sr3.GaussianDiffusion(net, image_size=64, timesteps=c["T"], loss_type=c["loss_type"], beta_schedule=c["schedule"],
                                         auto_normalize=False).to(DEV)
    diff.train()
    loss = diff.p_losses([noisy.to(DEV), clean.to(DEV)], noise=noise.to(DEV), level=level.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - float(o_loss)) <= 5e-3 * abs(float(o_loss)), (float(loss.detach()), float(o_loss))
    worst = {k: _rel(p.grad, o_grads[k]) for k, p in net.named_parameters()}
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:4]
    print(f"unet_sr3: loss {float(loss.detach()):.6f} (oracle {float(o_loss):.6f}); worst grad rel-RMS: {top}")
    bad = {k: v for k, v in worst.items() if v > 5e-2}
    assert not bad, f"gradient rel-RMS above 5e-2: {bad}"


@pytest.mark.parametrize("B", [1, 3])
def test_training_step_ragged_batches(B):
    """An epoch's last batch is ragged: B = 1 and B = 3 through both trainers (loss and a sample of gradients vs the oracle);
    trainers for two batch sizes stay resident side by side."""
    from hicdiff_b200 import hicdiff_condition
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    clean, noisy = O.synthetic_tiles(B, seed=77)
    t = torch.tensor([5, 400, 999][:B], dtype=torch.long)
    noise = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(3))
    buf = O.diffusion_buffers("linear", 1000)
    for kind in ("hicedrn", "unet"):
        torch.manual_seed(1)
        net = hicedrn_Diff(number_resnet=1, self_condition=True) if kind == "hicedrn" else hicdiff_condition.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        o_loss, o_grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type="l2", self_condition=True, num_blocks=1, net=kind)
        diff = hicdiff_condition.GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type="l2", beta_schedule="linear").to(DEV)
        loss = diff.p_losses([noisy.to(DEV), clean.to(DEV)], t=t.to(DEV), noise=noise.to(DEV))
        loss.backward()
        assert abs(float(loss.detach()) - float(o_loss)) <= 5e-3 * abs(float(o_loss)), (kind, B, float(loss.detach()), float(o_loss))
        worst = max(_rel(p.grad, o_grads[k]) for k, p in net.named_parameters())
        assert worst <= 6e-2, (kind, B, worst)
        # a second batch size keeps its own trainer; the first one is still there
        if B == 3:
            l1 = diff.p_losses([noisy[:1].to(DEV), clean[:1].to(DEV)], t=t[:1].to(DEV), noise=noise[:1].to(DEV))
            l1.backward()
            assert sorted(net._trainers) == [1, 3]


def test_two_forwards_before_backward_keep_their_own_gradients():
    """(loss1 + loss2).backward(): each loss node must deliver the gradients of ITS step, not the trainer's latest buffer."""
    c, net, sd, diff, clean, noisy, t, noise = _case("cond_l2")
    diff = diff.to(DEV)
    x = [noisy.to(DEV), clean.to(DEV)]
    t2 = torch.tensor([500, 2], device=DEV)
    l1 = diff.p_losses(x, t=t.to(DEV), noise=noise.to(DEV))
    l2 = diff.p_losses(x, t=t2, noise=noise.to(DEV))
    (l1 + l2).backward()
    both = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad()
    diff.p_losses(x, t=t.to(DEV), noise=noise.to(DEV)).backward()
    g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad()
    diff.p_losses(x, t=t2, noise=noise.to(DEV)).backward()
    for k, p in net.named_parameters():
        assert torch.equal(both[k], g1[k] + p.grad), k


def test_p2_loss_weights_reach_the_device_loss():
    """p2_loss_weight_gamma != 0: the per-sample weights (hicdiff_condition.py:519, :744) scale the loss and every gradient."""
    from hicdiff_b200 import hicdiff_condition
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    torch.manual_seed(2)
    net = hicedrn_Diff(number_resnet=1, self_condition=True)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    diff = hicdiff_condition.GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type="l2", beta_schedule="linear",
                                               p2_loss_weight_gamma=1.0, p2_loss_weight_k=1).to(DEV)
    clean, noisy = O.synthetic_tiles(3, seed=5)
    t = torch.tensor([2, 500, 990])
    noise = torch.randn(3, 1, 64, 64, generator=torch.Generator().manual_seed(8))
    buf = O.diffusion_buffers("linear", 1000, p2_gamma=1.0, p2_k=1)
    assert float(buf["p2_loss_weight"][t].max() / buf["p2_loss_weight"][t].min()) > 100           # the weights really differ
    o_loss, o_grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type="l2", self_condition=True, num_blocks=1)
    loss = diff.p_losses([noisy.to(DEV), clean.to(DEV)], t=t.to(DEV), noise=noise.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - float(o_loss)) <= 5e-3 * abs(float(o_loss))
    worst = max(_rel(p.grad, o_grads[k]) for k, p in net.named_parameters())
    assert worst <= 3e-2, worst


@pytest.mark.parametrize("kind", ["hicedrn", "unet"])
def test_graph_replay_equals_eager_step(kind):
    """Step 1 runs eagerly, step 2 is captured as a CUDA graph (with the side-stream schedule), step 3 replays it: with the same
    inputs and untouched weights all three must give bit-identical losses and gradients."""
    from hicdiff_b200 import hicdiff_condition
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    torch.manual_seed(4)
    net = hicedrn_Diff(number_resnet=3, self_condition=True) if kind == "hicedrn" else hicdiff_condition.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True)
    diff = hicdiff_condition.GaussianDiffusion(net, image_size=64, timesteps=1000, loss_type="l2", beta_schedule="linear").to(DEV)
    clean, noisy = O.synthetic_tiles(4, seed=21)
    t = torch.tensor([1, 250, 600, 999], device=DEV)
    noise = torch.randn(4, 1, 64, 64, generator=torch.Generator().manual_seed(6)).to(DEV)
    x = [noisy.to(DEV), clean.to(DEV)]
    runs = []
    for _ in range(3):
        net.zero_grad()
        loss = diff.p_losses(x, t=t, noise=noise)
        loss.backward()
        runs.append((loss.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    for later in runs[1:]:
        assert torch.equal(later[0], runs[0][0])
        for k in runs[0][1]:
            assert torch.equal(later[1][k], runs[0][1][k]), k


# ---- fused Adam (hd_adam_step) against torch.optim.Adam, the optimiser train.py:111 constructs ----
@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(weight_decay):
    """Same parameters, same gradients, 6 steps: parameters and both moments within 1 fp32 ulp-scale tolerance of torch's
    multi-tensor Adam (rtol 1e-6; measured: bit-identical).  Ragged sizes: 1 element, odd counts (scalar tail, unaligned
    neighbours), more than one 4096-element chunk, a 4-D conv weight."""
    from hicdiff_b200.optim import Adam as FusedAdam

    g = torch.Generator().manual_seed(11)
    shapes = [(1,), (7,), (64,), (3 * 4096 + 5,), (64, 32, 3, 3), (0,), (256, 256)]
    base = [torch.randn(s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    pb = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    ref = torch.optim.Adam(pa, foreach=True, **kw)
    ours = FusedAdam(pb, **kw)
    exact = True
    for it in range(6):
        grads = [torch.randn(s, generator=g).to(DEV) * (10.0 ** (it - 3)) for s in shapes]
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        ref.step()
        ours.step()
        for p, q in zip(pa, pb):
            assert torch.allclose(q, p, rtol=1e-6, atol=1e-9), (it, p.shape, float((p - q).abs().max()))
            exact = exact and torch.equal(p, q)
    for p, q in zip(pa, pb):
        if p.numel() == 0:
            continue
        m, v, step = ours.moments(q)
        st = ref.state[p]
        assert step == 6 and int(st["step"]) == 6
        assert torch.allclose(m, st["exp_avg"], rtol=1e-6, atol=1e-12) and torch.allclose(v, st["exp_avg_sq"], rtol=1e-6, atol=1e-12)
    print("fused Adam bit-identical to torch foreach Adam:", exact)


def test_fused_adam_drives_the_training_loop():
    """train.py:109-136 with the optimiser swapped: the same three steps under torch.optim.Adam and hicdiff_b200.optim.Adam leave
    the same parameters (the gradients are deterministic, the update arithmetic is torch's)."""
    from hicdiff_b200.hicdiff_condition import GaussianDiffusion
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff
    from hicdiff_b200.optim import Adam as FusedAdam

    outs = []
    for fused in (False, True):
        torch.manual_seed(3)
        net = hicedrn_Diff(number_resnet=2, self_condition=True)
        diff = GaussianDiffusion(net, image_size=64, timesteps=50, loss_type="l2", beta_schedule="linear", auto_normalize=False).to(DEV)
        diff.train()
        opt = (FusedAdam if fused else torch.optim.Adam)(diff.parameters(), lr=1e-4)
        clean, noisy = O.synthetic_tiles(4, seed=5)
        x = [noisy.to(DEV), clean.to(DEV)]
        g = torch.Generator().manual_seed(9)
        for _ in range(3):
            t = torch.randint(0, 50, (4,), generator=g)
            noise = torch.randn(4, 1, 64, 64, generator=g)
            loss = diff.p_losses(x, t=t.to(DEV), noise=noise.to(DEV))
            loss.backward()
            opt.step()
            opt.zero_grad()
        outs.append({k: p.detach().clone() for k, p in net.named_parameters()})
    for k in outs[0]:
        assert torch.allclose(outs[0][k], outs[1][k], rtol=1e-6, atol=1e-9), k


def test_fused_adam_rejects_what_it_cannot_do():
    from hicdiff_b200.optim import Adam as FusedAdam

    p = torch.nn.Parameter(torch.zeros(8, device=DEV))
    opt = FusedAdam([p], lr=1e-3)
    with pytest.raises(RuntimeError, match="no gradient"):
        opt.step()
    with pytest.raises(RuntimeError, match="closure"):
        opt.step(lambda: None)
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([p], betas=(1.0, 0.999))
    with pytest.raises(RuntimeError, match="CUDA fp32"):
        q = torch.nn.Parameter(torch.zeros(8))
        q.grad = torch.zeros(8)
        FusedAdam([q]).step()


def test_forward_after_fused_adam_step_sees_the_new_weights():
    """hd_adam_step updates the parameters through raw pointers; the sampling plan caches its own copies of
    the weights keyed on `Tensor._version`, so the optimiser must bump the versions.  After a fused step, `net(x, t)`, a
    no-grad validation `diffusion(x)` (train.py:150-170) and `super_resolution` must all run on the UPDATED weights, i.e.
    equal a fresh net loaded from the same state_dict."""
    from hicdiff_b200.hicdiff_condition import GaussianDiffusion
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff
    from hicdiff_b200.optim import Adam as FusedAdam

    torch.manual_seed(3)
    net = hicedrn_Diff(number_resnet=2, self_condition=True)
    diff = GaussianDiffusion(net, image_size=64, timesteps=20, loss_type="l2", beta_schedule="sigmoid", auto_normalize=False).to(DEV)
    diff.train()
    opt = FusedAdam(diff.parameters(), lr=5e-3)          # large enough that stale weights give a visibly different eps
    clean, noisy = O.synthetic_tiles(4, seed=5)
    x = [noisy.to(DEV), clean.to(DEV)]
    g = torch.Generator().manual_seed(9)
    xt = torch.randn(4, 1, 64, 64, generator=g).to(DEV)
    tt = torch.tensor([3, 7, 11, 19], device=DEV)
    before = net(xt, tt, noisy.to(DEV)).clone()            # builds the plan (and its weight copies) BEFORE the update
    versions = [p._version for p in net.parameters()]
    for _ in range(2):
        loss = diff.p_losses(x, t=torch.randint(0, 20, (4,), generator=g).to(DEV), noise=torch.randn(4, 1, 64, 64, generator=g).to(DEV))
        loss.backward()
        opt.step()
        opt.zero_grad()
    assert all(p._version > v for p, v in zip(net.parameters(), versions)), "fused Adam must bump Tensor._version"
    after = net(xt, tt, noisy.to(DEV))
    torch.manual_seed(4)
    fresh_net = hicedrn_Diff(number_resnet=2, self_condition=True)
    fresh = GaussianDiffusion(fresh_net, image_size=64, timesteps=20, loss_type="l2", beta_schedule="sigmoid", auto_normalize=False).to(DEV)
    fresh.load_state_dict(diff.state_dict(), strict=True)
    want = fresh_net(xt, tt, noisy.to(DEV))
    assert torch.equal(after, want), f"stale weights after a fused step: {float((after - want).abs().max()):.3e}"
    assert not torch.equal(after, before)
    noise = O.synthetic_noise(20, 4, seed=8).to(DEV)
    diff.eval()
    fresh.eval()
    assert torch.equal(diff.super_resolution(noisy.to(DEV), noise=noise), fresh.super_resolution(noisy.to(DEV), noise=noise))
    with torch.no_grad():
        nz = torch.randn(4, 1, 64, 64, generator=g).to(DEV)
        assert torch.equal(diff.p_losses(x, t=tt, noise=nz), fresh.p_losses(x, t=tt, noise=nz))


def test_fused_adam_state_dict_round_trip():
    """moments and the step count live in the C handle; state_dict / load_state_dict carry them in
    torch.optim.Adam's own layout, so a resumed run continues instead of restarting from zero moments, and a checkpoint
    written by torch.optim.Adam loads into the fused optimiser (and vice versa)."""
    from hicdiff_b200.optim import Adam as FusedAdam

    g = torch.Generator().manual_seed(21)
    shapes = [(5,), (3 * 4096 + 1,), (16, 8, 3, 3)]
    base = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(5)]

    def run(opt_cls, params, its):
        for it in its:
            for p, gr in zip(params, grads[it]):
                p.grad = gr.clone().to(DEV)
            opt_cls.step()

    kw = dict(lr=2e-3, weight_decay=0.01)
    # uninterrupted fused run of 5 steps
    pa = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    oa = FusedAdam(pa, **kw)
    run(oa, pa, range(5))
    # 3 steps, checkpoint, fresh optimiser (no step taken yet -> state is applied when the handle is created), 2 more
    pb = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    ob = FusedAdam(pb, **kw)
    run(ob, pb, range(3))
    ck = ob.state_dict()
    assert set(ck["state"]) == {0, 1, 2} and float(ck["state"][0]["step"]) == 3.0
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pb]
    oc = FusedAdam(pc, **kw)
    oc.load_state_dict(ck)
    run(oc, pc, range(3, 5))
    for p, q in zip(pa, pc):
        assert torch.equal(p, q), "resumed run differs from the uninterrupted one"
    assert oc.moments(pc[1])[2] == 5
    # the same checkpoint drives torch.optim.Adam to the same place (layout compatibility)
    pd = [torch.nn.Parameter(p.detach().clone()) for p in pb]
    od = torch.optim.Adam(pd, foreach=True, **kw)
    od.load_state_dict(ck)
    run(od, pd, range(3, 5))
    for p, q in zip(pa, pd):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-9)
    # and torch's checkpoint loads into a fused optimiser that has already stepped (handle exists)
    pe = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    oe = FusedAdam(pe, **kw)
    run(oe, pe, range(1))
    pt = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    ot = torch.optim.Adam(pt, foreach=True, **kw)
    run(ot, pt, range(3))
    oe.load_state_dict(ot.state_dict())
    with torch.no_grad():
        for p, q in zip(pe, pt):
            p.copy_(q)
    run(oe, pe, range(3, 5))
    run(ot, pt, range(3, 5))
    for p, q in zip(pe, pt):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-9)
