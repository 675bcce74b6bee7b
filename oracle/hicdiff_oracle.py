"""CPU oracle for the HiCDiff reverse-diffusion sampling path.  TEST INFRASTRUCTURE ONLY.

This is a functional, state_dict-driven restatement (plain torch fp32 on CPU + numpy for the tile indexing) of the
reference algorithm; it is the CHECKER for the CUDA path and must never be imported by the product package
(`hicdiff_b200/`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this path, so the
oracle is pinned against the reference itself: `oracle/make_golden.py` imports the unmodified reference from
/root/reference, checks this restatement against it bit-for-bit (same torch ops in the same order) and writes the
fixtures under tests/golden/ that travel to the GPU box (where /root/reference does not exist).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------------------------
# Schedules and the 13 fp32 [T] buffers      src/hicdiff_condition.py:393-427, 469-519 (sr3: 460-587)
# --------------------------------------------------------------------------------------------------------------


def beta_schedule(name: str, timesteps: int) -> torch.Tensor:
    if name == "linear":  # :393-400
        scale = 1000 / timesteps
        return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
    if name == "cosine":  # :402-412
        s = 0.008
        t = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
        ac = torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    if name == "sigmoid":  # :414-427
        start, end, tau = -3, 3, 1
        t = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
        v_start = torch.tensor(start / tau).sigmoid()
        v_end = torch.tensor(end / tau).sigmoid()
        ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    raise ValueError(f"unknown beta schedule {name}")  # :467


def diffusion_buffers(name: str, timesteps: int, p2_gamma: float = 0.0, p2_k: float = 1) -> SD:
    """The registered buffers of GaussianDiffusion.__init__ (float64 math, cast to fp32 on registration :489)."""
    betas = beta_schedule(name, timesteps)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    b64 = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
        "p2_loss_weight": (p2_k + ac / (1 - ac)) ** -p2_gamma,
    }
    return {k: v.to(torch.float32) for k, v in b64.items()}


def sr3_noise_levels(name: str, timesteps: int) -> torch.Tensor:
    """float64 [T+1] table `sqrt_alphas_cumprod_prev` of src/hicdiff_sr3.py:535-536 (two left pads with 1)."""
    betas = beta_schedule(name, timesteps)
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    return torch.sqrt(F.pad(ac_prev, (1, 0), value=1.0))


# --------------------------------------------------------------------------------------------------------------
# UNet building blocks                                            src/hicdiff_condition.py:64-251
# --------------------------------------------------------------------------------------------------------------


def _ws_conv3x3(x, w, b):  # WeightStandardizedConv2d.forward :89-97 (fp32 -> eps 1e-5)
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    wn = (w - mean) * (var + 1e-5).rsqrt()
    return F.conv2d(x, wn, b, padding=1)


def _block(sd: SD, p: str, x, scale_shift=None):  # Block.forward :162-171
    x = _ws_conv3x3(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    x = F.group_norm(x, 8, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def _resnet_block(sd: SD, p: str, x, t_emb, sr3: bool):
    """ResnetBlock.forward :185-197; SR3 flavour src/hicdiff_sr3.py:246-251 (+FeatureWiseAffine :175-183)."""
    if sr3:
        h = _block(sd, p + ".block1", x)
        nf = F.linear(t_emb, sd[p + ".noise_func.noise_func.0.weight"], sd[p + ".noise_func.noise_func.0.bias"])
        h = h + nf.view(x.shape[0], -1, 1, 1)
    else:
        te = F.linear(F.silu(t_emb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
        te = te[:, :, None, None]
        h = _block(sd, p + ".block1", x, te.chunk(2, dim=1))
    h = _block(sd, p + ".block2", h)
    if (p + ".res_conv.weight") in sd:
        return h + F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    return h + x


def _chan_layernorm(x, g):  # LayerNorm.forward :104-108
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def _linear_attention(sd: SD, p: str, x, heads=4):
    """Residual(PreNorm(LinearAttention)) :64-70, 110-118, 199-227.  `p` is the Residual's prefix."""
    b, c, h, w = x.shape
    xn = _chan_layernorm(x, sd[p + ".fn.norm.g"])
    qkv = F.conv2d(xn, sd[p + ".fn.fn.to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, -1, h * w) for t in qkv)
    q = q.softmax(dim=-2)
    k = k.softmax(dim=-1)
    q = q * (32 ** -0.5)
    v = v / (h * w)
    context = torch.einsum("b h d n, b h e n -> b h d e", k, v)
    out = torch.einsum("b h d e, b h d n -> b h e n", context, q)
    out = out.reshape(b, -1, h, w)
    out = F.conv2d(out, sd[p + ".fn.fn.to_out.0.weight"], sd[p + ".fn.fn.to_out.0.bias"])
    out = _chan_layernorm(out, sd[p + ".fn.fn.to_out.1.g"])
    return out + x


def _attention(sd: SD, p: str, x, heads=4):
    """Residual(PreNorm(Attention)) :229-251."""
    b, c, h, w = x.shape
    xn = _chan_layernorm(x, sd[p + ".fn.norm.g"])
    qkv = F.conv2d(xn, sd[p + ".fn.fn.to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, -1, h * w) for t in qkv)
    q = q * (32 ** -0.5)
    sim = torch.einsum("b h d i, b h d j -> b h i j", q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("b h i j, b h d j -> b h i d", attn, v)
    out = out.permute(0, 1, 3, 2).reshape(b, -1, h, w)  # 'b h (x y) d -> b (h d) x y'
    return F.conv2d(out, sd[p + ".fn.fn.to_out.weight"], sd[p + ".fn.fn.to_out.bias"]) + x


def _pixel_unshuffle(x):  # Rearrange('b c (h p1) (w p2) -> b (c p1 p2) h w') :80
    b, c, hh, ww = x.shape
    x = x.reshape(b, c, hh // 2, 2, ww // 2, 2)
    return x.permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, hh // 2, ww // 2)


def sinusoidal_pos_emb(t: torch.Tensor, dim: int):  # SinusoidalPosEmb.forward :127-134
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=t.device) * -emb)
    emb = t[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def positional_encoding(level: torch.Tensor, dim: int):  # PositionalEncoding.forward src/hicdiff_sr3.py:160-165
    count = dim // 2
    step = torch.arange(count, dtype=level.dtype, device=level.device) / count
    enc = level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    return torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)


def time_mlp(sd: SD, emb):  # nn.Sequential(Linear, GELU, Linear) :300-305
    h = F.gelu(F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return F.linear(h, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def unet_forward(sd: SD, x, time, x_self_cond=None, *, dim=64, dim_mults: Sequence[int] = (1, 2, 4, 8),
                 self_condition=True, sr3=False, taps: Optional[dict] = None):
    """Unet.forward  src/hicdiff_condition.py:345-384 (uncond src/hicdiff.py:345-388; SR3 src/hicdiff_sr3.py:406-445).

    `sd` holds the eps-net parameters WITHOUT the `model.` prefix.  `time` is long [B] (or float [B,1] noise levels
    when sr3).  `taps`, when given, collects named intermediates for layer-wise parity.
    """
    def tap(name, v):
        if taps is not None:
            taps[name] = v

    if self_condition:
        x = torch.cat((x_self_cond, x), dim=1)  # :348
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)  # :350
    tap("init_conv", x)
    r = x.clone()
    if sr3:
        t = time_mlp(sd, positional_encoding(time, dim))  # [B,1,256]
    else:
        t = time_mlp(sd, sinusoidal_pos_emb(time, dim))
    L = len(dim_mults)
    hs: List[torch.Tensor] = []
    for i in range(L):  # :357-365
        p = f"downs.{i}"
        x = _resnet_block(sd, p + ".0", x, t, sr3)
        tap(p + ".0", x)
        hs.append(x)
        x = _resnet_block(sd, p + ".1", x, t, sr3)
        tap(p + ".1", x)
        x = _linear_attention(sd, p + ".2", x)
        tap(p + ".2", x)
        hs.append(x)
        if i < L - 1:
            x = F.conv2d(_pixel_unshuffle(x), sd[p + ".3.1.weight"], sd[p + ".3.1.bias"])
        else:
            x = F.conv2d(x, sd[p + ".3.weight"], sd[p + ".3.bias"], padding=1)
        tap(p + ".3", x)
    x = _resnet_block(sd, "mid_block1", x, t, sr3)  # :367-369
    tap("mid_block1", x)
    x = _attention(sd, "mid_attn", x)
    tap("mid_attn", x)
    x = _resnet_block(sd, "mid_block2", x, t, sr3)
    tap("mid_block2", x)
    for k in range(L):  # :371-379
        p = f"ups.{k}"
        x = torch.cat((x, hs.pop()), dim=1)
        x = _resnet_block(sd, p + ".0", x, t, sr3)
        tap(p + ".0", x)
        x = torch.cat((x, hs.pop()), dim=1)
        x = _resnet_block(sd, p + ".1", x, t, sr3)
        tap(p + ".1", x)
        x = _linear_attention(sd, p + ".2", x)
        tap(p + ".2", x)
        if k < L - 1:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[p + ".3.1.weight"], sd[p + ".3.1.bias"], padding=1)
        else:
            x = F.conv2d(x, sd[p + ".3.weight"], sd[p + ".3.bias"], padding=1)
        tap(p + ".3", x)
    x = torch.cat((x, r), dim=1)  # :381
    x = _resnet_block(sd, "final_res_block", x, t, sr3)
    tap("final_res_block", x)
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])  # :384


# --------------------------------------------------------------------------------------------------------------
# HiCEDRN eps-net                                    src/model/hicedrn_Diff.py:169-289, hicedrn_sr3_Diff.py:245-265
# --------------------------------------------------------------------------------------------------------------


def hicedrn_forward(sd: SD, x, time, x_self_cond=None, *, self_condition=False, sr3=False, num_blocks=32,
                    taps: Optional[dict] = None):
    n_feat = 256
    if self_condition:
        x = torch.cat((x_self_cond, x), dim=1)  # :273
    x = F.conv2d(x, sd["head.weight"], sd["head.bias"], padding=1)  # :275
    r = x.clone()
    if sr3:
        t = time_mlp(sd, positional_encoding(time, n_feat))
    else:
        t = time_mlp(sd, sinusoidal_pos_emb(time, n_feat))
    for i in range(num_blocks):  # ResnetBlock.forward :194-208 -- the SAME conv is applied twice
        p = f"body.{i}"
        w, b = sd[p + ".conv.proj.weight"], sd[p + ".conv.proj.bias"]
        h = F.conv2d(x, w, b, padding=1)
        if sr3:
            nf = F.linear(t, sd[p + ".noise_func.noise_func.0.weight"], sd[p + ".noise_func.noise_func.0.bias"])
            h = h + nf.view(x.shape[0], -1, 1, 1)
        else:
            te = F.linear(F.silu(t), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])[:, :, None, None]
            scale, shift = te.chunk(2, dim=1)
            h = h * (scale + 1) + shift
        h = F.silu(h)
        h = F.conv2d(h, w, b, padding=1)
        h = h * 0.1
        x = torch.add(h, x)
        if taps is not None:
            taps[p] = x
    x = F.conv2d(x, sd["body_tail.weight"], sd["body_tail.bias"], padding=1)  # :283
    x = x + r
    if taps is not None:
        taps["body_tail"] = x
    return F.conv2d(x, sd["tail.weight"], sd["tail.bias"], padding=1)  # :287


# --------------------------------------------------------------------------------------------------------------
# Reverse process                                          src/hicdiff_condition.py:526-623
# --------------------------------------------------------------------------------------------------------------


def _extract(a, t, x_shape):  # :388-391
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def p_sample(eps_fn, buf: SD, x, t: int, cond, noise, *, sr3_levels: Optional[torch.Tensor] = None, objective: str = "pred_noise"):
    """One reverse step: p_mean_variance :581-589 + p_sample :591-598 (SR3: src/hicdiff_sr3.py:634-652).

    eps_fn(x, time, cond) -> eps.  `noise` is the z tensor (ignored at t == 0).  Returns (x_{t-1}, x_start, eps).
    """
    b = x.shape[0]
    if sr3_levels is None:
        bt = torch.full((b,), t, dtype=torch.long, device=x.device)
        eps = eps_fn(x, bt, cond)
        if objective == "pred_noise":
            x_start = _extract(buf["sqrt_recip_alphas_cumprod"], bt, x.shape) * x - \
                _extract(buf["sqrt_recipm1_alphas_cumprod"], bt, x.shape) * eps  # :526-530
        elif objective == "pred_x0":  # model_predictions :568-571
            x_start = eps
        else:  # pred_v: predict_start_from_v :544-548
            x_start = _extract(buf["sqrt_alphas_cumprod"], bt, x.shape) * x - \
                _extract(buf["sqrt_one_minus_alphas_cumprod"], bt, x.shape) * eps
        x_start = x_start.clamp(-1.0, 1.0)  # :586
        mean = _extract(buf["posterior_mean_coef1"], bt, x.shape) * x_start + \
            _extract(buf["posterior_mean_coef2"], bt, x.shape) * x  # :550-554
        logvar = _extract(buf["posterior_log_variance_clipped"], bt, x.shape)
    else:
        level = torch.FloatTensor([sr3_levels[t + 1]]).repeat(b, 1).to(x.device)  # sr3 :636
        eps = eps_fn(x, level, cond)
        x_start = buf["sqrt_recip_alphas_cumprod"][t] * x - buf["sqrt_recipm1_alphas_cumprod"][t] * eps
        x_start = x_start.clamp(-1.0, 1.0)
        mean = buf["posterior_mean_coef1"][t] * x_start + buf["posterior_mean_coef2"][t] * x
        logvar = buf["posterior_log_variance_clipped"][t]
    z = noise if t > 0 else torch.zeros_like(x)  # :596
    return mean + (0.5 * logvar).exp() * z, x_start, eps


def p_sample_loop(eps_fn, buf: SD, cond, noise: torch.Tensor, *, timesteps: int,
                  sr3_levels: Optional[torch.Tensor] = None, return_all: bool = False, t_end: int = 0, objective: str = "pred_noise"):
    """p_sample_loop :600-623 with INJECTED noise: noise[0] = x_T (:605), noise[i] = z of step t = T - i (:596)."""
    img = noise[0]
    trace = []
    for t in reversed(range(t_end, timesteps)):
        z = noise[timesteps - t] if t > 0 else None
        img, _, _ = p_sample(eps_fn, buf, img, t, cond, z, sr3_levels=sr3_levels, objective=objective)
        if return_all:
            trace.append(img)
    return (img, trace) if return_all else img


def ddim_sample(eps_fn, buf: SD, noise, *, timesteps: int, sampling_timesteps: int, eta: float = 0.0, return_all=False):
    """ddim_sample src/hicdiff.py:623-664 with the draws supplied: noise[0] = x_T, then one z per step except the last."""
    times = list(reversed(torch.linspace(-1, timesteps - 1, steps=sampling_timesteps + 1).int().tolist()))
    draws = iter(noise)
    img = next(draws).clone()
    imgs = [img]
    for time, time_next in zip(times[:-1], times[1:]):
        t = torch.full((img.shape[0],), time, dtype=torch.long)
        pred_noise = eps_fn(img, t, None)
        x_start = _extract(buf["sqrt_recip_alphas_cumprod"], t, img.shape) * img - _extract(buf["sqrt_recipm1_alphas_cumprod"], t, img.shape) * pred_noise
        x_start = torch.clamp(x_start, min=-1.0, max=1.0)
        if time_next < 0:
            img = x_start
            imgs.append(img)
            continue
        alpha, alpha_next = buf["alphas_cumprod"][time], buf["alphas_cumprod"][time_next]
        sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        img = x_start * alpha_next.sqrt() + c * pred_noise + sigma * next(draws)
        imgs.append(img)
    return torch.stack(imgs, dim=1) if return_all else img


def q_sample(buf: SD, x_start, t, noise):  # :698-704
    return _extract(buf["sqrt_alphas_cumprod"], t, x_start.shape) * x_start + \
        _extract(buf["sqrt_one_minus_alphas_cumprod"], t, x_start.shape) * noise


def p_losses(eps_fn, buf: SD, noisy, clean, t, noise, *, loss_type="l2", self_condition=True, objective="pred_noise"):
    """Conditional p_losses :715-746 with t and noise supplied (the reference draws them :719,721)."""
    x = q_sample(buf, clean, t, noise)
    out = eps_fn(x, t, noisy if self_condition else None)
    if objective == "pred_noise":
        target = noise
    else:
        # reference quirk (:716,:733-737): the conditional p_losses unpacks `x_start, x_end = x_in`, diffuses x_end (the clean
        # tile) but builds the pred_x0 / pred_v targets from `x_start` -- the NOISY conditional input
        ref = noisy if self_condition else clean
        if objective == "pred_x0":
            target = ref
        else:  # predict_v :538-542
            target = _extract(buf["sqrt_alphas_cumprod"], t, ref.shape) * noise - _extract(buf["sqrt_one_minus_alphas_cumprod"], t, ref.shape) * ref
    loss = F.mse_loss(out, target, reduction="none") if loss_type == "l2" else F.l1_loss(out, target, reduction="none")
    loss = loss.reshape(loss.shape[0], -1)  # reduce(loss, 'b ... -> b (...)', 'mean') keeps every element :743
    loss = loss * _extract(buf["p2_loss_weight"], t, loss.shape)
    return loss.mean()


def sr3_p_losses(eps_fn, noisy, clean, level, noise, *, loss_type="l2", self_condition=True):
    """SR3 p_losses (src/hicdiff_sr3.py:750-792) with the noise level and the noise supplied (the reference draws them from
    numpy's global RNG :760-770): x = level * x0 + sqrt(1 - level^2) * noise; plain mean loss, no p2 weight."""
    lv = level.view(-1, 1, 1, 1)
    x = lv * clean + (1 - lv ** 2).sqrt() * noise          # q_sample :735-739
    out = eps_fn(x, level.view(clean.shape[0], -1), noisy if self_condition else None)
    return (F.mse_loss(out, noise, reduction="none") if loss_type == "l2" else F.l1_loss(out, noise, reduction="none")).mean()


def sr3_p_losses_and_grads(sd: SD, noisy, clean, level, noise, *, loss_type="l2", self_condition=True, num_blocks=32, net="hicedrn"):
    """Loss and parameter gradients of one SR3 training iteration over hicedrn_sr3_Diff (pretrain/train_hicedrn_Diff_sr3.py) or
    the SR3 Unet (pretrain/train_unet_Diff_sr3.py)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if torch.is_floating_point(v)}
    if net == "unet":
        eps_fn = lambda x, lv, c: unet_forward(leaves, x, lv, c, self_condition=self_condition, sr3=True)  # noqa: E731
    else:
        eps_fn = lambda x, lv, c: hicedrn_forward(leaves, x, lv, c, self_condition=self_condition, sr3=True, num_blocks=num_blocks)  # noqa: E731
    loss = sr3_p_losses(eps_fn, noisy, clean, level, noise, loss_type=loss_type, self_condition=self_condition)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), dict(zip(leaves.keys(), grads))


def p_losses_and_grads(sd: SD, buf: SD, noisy, clean, t, noise, *, loss_type="l2", self_condition=True, num_blocks=32, net="hicedrn",
                       objective="pred_noise"):
    """One training iteration's loss and d loss / d parameter for the hicedrn_Diff eps-net: what `loss = diffusion(x);
    loss.backward()` leaves in `.grad` (train.py:127-128) -- torch.autograd over the restated forward (the reference has no
    hand-written backward to cite).  Returns (loss, {state_dict key: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if torch.is_floating_point(v)}
    if net == "unet":
        eps_fn = lambda x, tt, c: unet_forward(leaves, x, tt, c, self_condition=self_condition)  # noqa: E731
    else:
        eps_fn = lambda x, tt, c: hicedrn_forward(leaves, x, tt, c, self_condition=self_condition, num_blocks=num_blocks)  # noqa: E731
    loss = p_losses(eps_fn, buf, noisy, clean, t, noise, loss_type=loss_type, self_condition=self_condition, objective=objective)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), dict(zip(leaves.keys(), grads))


# --------------------------------------------------------------------------------------------------------------
# DDRM sampler for the denoising operator     src/functions/denoising.py:6-111, svd_replacement.py:148-168
# --------------------------------------------------------------------------------------------------------------


def ddrm_compute_alpha(beta, t):  # :6-9
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def ddrm_denoising_steps(eps_fn, betas, x, seq, y_0, sigma_0, etaB, etaA, etaC, noise):
    """efficient_generalized_steps with H = Denoising (U = V = I, singulars == 1): the masked updates of :88-97 cover all
    elements, one case per step.  `noise[k]` is the z the reference USES at step k (its third draw when sigma_next > sigma_0,
    its second otherwise).  Returns (xs, x0_preds) like the reference."""
    seq = list(seq)
    n = x.size(0)
    y = y_0.reshape(n, -1)
    la = ddrm_compute_alpha(betas, (torch.ones(n) * seq[-1]).long())
    largest_sigmas = (1 - la).sqrt() / la.sqrt()
    big = bool(largest_sigmas[0, 0, 0, 0] > sigma_0)                                        # :22
    inv = torch.full((1, y.shape[1]), float(sigma_0) if big else 0.0)                        # :25-27 (sigma_0 / 1)
    init_y = (y if big else torch.zeros_like(y)).view(*x.size())                             # :31-33
    remaining_s = (largest_sigmas.view(-1, 1) ** 2 - inv ** 2).view(*x.shape).clamp_min(0.0).sqrt()   # :34-35
    x = (init_y + remaining_s * x) / largest_sigmas                                          # :36-40
    seq_next = [-1] + seq[:-1]
    xs, x0_preds = [x], []
    for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
        t = torch.ones(n) * i
        at = ddrm_compute_alpha(betas, (torch.ones(n) * i).long())
        at_next = ddrm_compute_alpha(betas, (torch.ones(n) * j).long())
        xt = xs[-1]
        et = eps_fn(xt, t)
        x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()                                       # :68
        sigma_next = (1 - at_next).sqrt()[0, 0, 0, 0] / at_next.sqrt()[0, 0, 0, 0]           # :72
        v0, e0, z = x0_t.reshape(n, -1), et.reshape(n, -1), noise[k].reshape(n, -1)
        if bool(sigma_next > sigma_0):                                                       # :96-97
            d = torch.sqrt(sigma_next ** 2 - sigma_0 ** 2 / torch.ones(()) ** 2 * (etaB ** 2))
            nxt = y * etaB + (1 - etaB) * v0 + d * z
        elif bool(sigma_next < sigma_0):                                                     # :92-93
            std = sigma_next * etaA
            nxt = v0 + torch.sqrt(sigma_next ** 2 - std ** 2) * ((y - v0) / sigma_0) + std * z
        else:                                                                                # :89
            std = sigma_next * etaC
            nxt = v0 + torch.sqrt(sigma_next ** 2 - std ** 2) * e0 + std * z
        xs.append((at_next.sqrt()[0, 0, 0, 0] * nxt).view(*x.shape))                         # :100-101
        x0_preds.append(x0_t)
    return xs, x0_preds


# --------------------------------------------------------------------------------------------------------------
# Tiling                                                   processdata/PrepareData_linear.py:25-46
# --------------------------------------------------------------------------------------------------------------


def band_blocks_for(res: int, piece: int = 64) -> int:
    """Largest block distance kept by `abs(i - j) <= int(piece_size * 4 * scal + 1)` with step == piece (:31,42)."""
    scal = int(40000 / res)
    return int(piece * 4 * scal + 1) // piece


def split_pieces(mat: np.ndarray, piece: int = 64, res: int = 40000) -> np.ndarray:
    """splitPieces(fn, piece_size, step=piece_size, resol) on an in-memory matrix (:25-46)."""
    n = mat.shape[0]
    assert mat.shape[1] == n
    rest = n % piece
    if rest != 0:
        pad = piece - rest
        mat = np.pad(mat, ((0, pad), (0, pad)), constant_values=0.0)
    bound = mat.shape[0]
    limit = int(piece * 4 * int(40000 / res) + 1)
    pieces = []
    for i in range(0, bound, piece):
        for j in range(i, bound, piece):
            if abs(i - j) <= limit and i + piece <= bound and j + piece <= bound:
                pieces.append(mat[i:i + piece, j:j + piece])
    if not pieces:
        return np.zeros((0, 1, piece, piece), dtype=mat.dtype)
    return np.expand_dims(np.asarray(pieces), 1)


def reassemble(tiles: np.ndarray, n: int, piece: int = 64, res: int = 40000) -> np.ndarray:
    """Exact inverse of split_pieces' enumeration (the reference has none): tile k goes to block
    (i, j), its transpose to (j, i) when i != j, the padding is cropped, cells outside the band stay 0."""
    P = -(-n // piece)
    limit = int(piece * 4 * int(40000 / res) + 1)
    full = np.zeros((P * piece, P * piece), dtype=tiles.dtype)
    k = 0
    for i in range(0, P * piece, piece):
        for j in range(i, P * piece, piece):
            if abs(i - j) <= limit:
                t = tiles[k, 0]
                full[i:i + piece, j:j + piece] = t
                if i != j:
                    full[j:j + piece, i:i + piece] = t.T
                k += 1
    assert k == tiles.shape[0], f"expected {k} tiles, got {tiles.shape[0]}"
    return full[:n, :n]


# --------------------------------------------------------------------------------------------------------------
# Synthetic inputs: plain data generators shared with the tests / bench
# --------------------------------------------------------------------------------------------------------------
from hicdiff_b200.synthetic import synthetic_noise, synthetic_tiles  # noqa: E402,F401

# --------------------------------------------------------------------------------------------------------------
# Data preparation                                        processdata/PrepareData_linear.py:48-103,183-213
# --------------------------------------------------------------------------------------------------------------


def synthetic_contacts(n_bins: int, res: int = 40000, seed: int = 0, offset_bins: int = 3, empty_every: int = 17,
                       duplicates: int = 50):
    """Seeded (pos1, pos2, value) triples in the text-dump format loadBothConstraints reads (:49-50): genomic positions in
    bp, band-limited symmetric contacts with a decaying count profile, some bins left without a diagonal entry (they get
    removed, :77-85) and a few duplicated cells (later lines overwrite earlier ones, :67-72)."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for i in range(n_bins):
        if empty_every and i % empty_every == empty_every - 1:
            continue
        for j in range(i, min(n_bins, i + 40)):
            if j != i and (empty_every and j % empty_every == empty_every - 1):
                continue
            if j != i and rng.random() < 0.35:
                continue
            rows.append(i)
            cols.append(j)
            vals.append(float(np.float32(rng.gamma(2.0, 30.0 / (1 + (j - i))))))
    rows, cols, vals = np.array(rows), np.array(cols), np.array(vals)
    pick = rng.integers(0, len(rows), duplicates)
    rows = np.concatenate([rows, rows[pick]])
    cols = np.concatenate([cols, cols[pick]])
    vals = np.concatenate([vals, vals[pick] * 0.5 + 1.0])
    perm = rng.permutation(len(rows))
    rows, cols, vals = rows[perm], cols[perm], vals[perm]
    pos = np.stack([(rows + offset_bins) * res + rng.integers(0, res, len(rows)) * 0,      # bin starts, like `cooler dump`
                    (cols + offset_bins) * res, vals], axis=1)
    return pos.astype(np.float64)


def load_constraints(triples_a: np.ndarray, triples_b: np.ndarray, res: int) -> np.ndarray:
    """loadBothConstraints :48-103 on already-loaded `np.loadtxt` arrays: returns the normalised matrix `mata`."""
    rowsa = (triples_a[:, 0] / res).astype(int)
    colsa = (triples_a[:, 1] / res).astype(int)
    valsa = triples_a[:, 2]
    rowsb = (triples_b[:, 0] / res).astype(int)
    colsb = (triples_b[:, 1] / res).astype(int)
    bigbin = np.max((np.max((rowsa, colsa)), np.max((rowsb, colsb))))
    smallbin = np.min((np.min((rowsa, colsa)), np.min((rowsb, colsb))))
    mata = np.zeros((bigbin - smallbin + 1, bigbin - smallbin + 1), dtype="float32")
    for ra, ca, ia in zip(rowsa, colsa, valsa):  # :67-70
        mata[ra - smallbin, ca - smallbin] = ia
        mata[ca - smallbin, ra - smallbin] = ia
    diaga = np.diag(mata)
    removeidx = np.unique(np.concatenate((np.argwhere(diaga == 0)[:, 0], np.argwhere(np.isnan(diaga))[:, 0])))  # :80
    mata = np.delete(mata, removeidx, axis=0)
    mata = np.delete(mata, removeidx, axis=1)
    per_a = np.percentile(mata, 99.0)  # :88
    mata = np.clip(mata, 0, per_a)
    mata = mata / per_a
    mata = 2 * mata - 1.0
    return mata


def add_noise(tiles: torch.Tensor, sigma_0: float, noise: torch.Tensor) -> torch.Tensor:
    """split_numpy :199-204 for deg == 'deno' (H and H_pinv are the identity, svd_replacement.py:148-168)."""
    return tiles + sigma_0 * noise


# --------------------------------------------------------------------------------------------------------------
# Quality metrics used for the 1e-3 SSIM/PSNR bar            src/Utils/loss/SSIM.py:6-74,
# src/datasets/__init__.py:214-223 (inverse_data_transform 'rescaled'), pretrain/train_unet_Diff_cond_n.py:125-133
# --------------------------------------------------------------------------------------------------------------


def to_unit_range(x):
    """inverse_data_transform('rescaled', X): [-1,1] -> clamp((X+1)/2, 0, 1)."""
    return torch.clamp((x + 1.0) / 2.0, 0.0, 1.0)


def ssim(img1, img2, window_size: int = 11):
    """Mean SSIM with an 11x11 gaussian (sigma 1.5) window, C1 = 0.01^2, C2 = 0.03^2, zero 'same' padding."""
    c = img1.shape[1]
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    win = g.mm(g.t()).float()[None, None].expand(c, 1, window_size, window_size).contiguous().to(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, win, padding=pad, groups=c)
    mu2 = F.conv2d(img2, win, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=c) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def psnr(img1, img2):
    """10 * log10(1 / mse) on [0, 1] images."""
    mse = ((img1 - img2) ** 2).mean()
    return 10 * torch.log10(1 / mse)
