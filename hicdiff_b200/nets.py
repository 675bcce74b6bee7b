"""eps-predictor modules with the reference's constructor arguments, attributes and state_dict layout, whose
forward runs on the sm_100a plan (C ABI) instead of ATen.

The classes below are PARAMETER HOLDERS: they create `nn.Conv2d` / `nn.Linear` / `nn.GroupNorm` children under the
reference's attribute names and in the reference's construction order, so that
  * `state_dict()` / `load_state_dict(strict=True)` are key-for-key, shape-for-shape interchangeable with
    /root/reference/src/hicdiff_condition.py:255-343 (Unet), src/hicdiff_sr3.py:310-404 and
    /root/reference/src/model/hicedrn_Diff.py:210-265, hicedrn_sr3_Diff.py:267-330;
  * default initialisation under a given torch seed yields the same weights as the reference (same RNG draws).
None of the holders implements the layer math -- `forward` of the top-level nets hands the whole network to
`EpsPlan` (plan.py), i.e. to libhicdiff_b200.so.  There is no ATen fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib
from .plan import EpsPlan

# ---------------------------------------------------------------------------------------------------- holders


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} only holds parameters; run the enclosing Unet/hicedrn_Diff (sm_100a plan) instead"
        )


class _TimeEmbedding(_Holder):
    """Parameter-free slot 0 of `time_mlp` (SinusoidalPosEmb / PositionalEncoding live in prep.cu)."""

    def __init__(self, dim: int, kind: str):
        super().__init__()
        self.dim = dim
        self.kind = kind


class _NoiseFunc(_Holder):  # FeatureWiseAffine (hicdiff_sr3.py:167-183): keys `noise_func.noise_func.0.*`
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.noise_func = nn.Sequential(nn.Linear(in_features, out_features))


class _NormedConv(_Holder):  # Block (hicdiff_condition.py:155-160): keys `proj.*`, `norm.*`
    def __init__(self, dim: int, dim_out: int, groups: int):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)


class _ResBlock(_Holder):  # ResnetBlock (hicdiff_condition.py:173-183 / hicdiff_sr3.py:235-244)
    def __init__(self, dim: int, dim_out: int, time_dim: int, groups: int, sr3: bool):
        super().__init__()
        if sr3:
            self.noise_func = _NoiseFunc(time_dim, dim_out)
        else:
            self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, dim_out * 2))
        self.block1 = _NormedConv(dim, dim_out, groups)
        self.block2 = _NormedConv(dim_out, dim_out, groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class _Gain(_Holder):  # LayerNorm (hicdiff_condition.py:99-102): key `g`
    def __init__(self, dim: int):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))


class _LinAttn(_Holder):  # LinearAttention (hicdiff_condition.py:199-210)
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), _Gain(dim))


class _Attn(_Holder):  # Attention (hicdiff_condition.py:229-237)
    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


class _PreNorm(_Holder):  # PreNorm (hicdiff_condition.py:110-114): keys `fn.*`, `norm.g`
    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.fn = fn
        self.norm = _Gain(dim)


class _Residual(_Holder):  # Residual (hicdiff_condition.py:64-67): key `fn.*`
    def __init__(self, fn: nn.Module):
        super().__init__()
        self.fn = fn


def _resample(conv: nn.Conv2d) -> nn.Sequential:
    """Downsample / Upsample are Sequential(<parameter-free op>, Conv2d): the conv must sit at index 1."""
    return nn.Sequential(nn.Identity(), conv)


# ---------------------------------------------------------------------------------------------------- base


class _EpsNet(nn.Module):
    """Shared forward plumbing: owns an `EpsPlan` (device-resident weights in GEMM layout + CUDA graphs)."""

    _variant: int = _lib.HD_UNET

    def _plan_config(self) -> dict:
        raise NotImplementedError

    def _init_plan_state(self):
        # not a submodule / not in state_dict
        object.__setattr__(self, "_eps_plan", EpsPlan(self))

    @property
    def eps_plan(self) -> EpsPlan:
        return self._eps_plan

    def forward(self, x, time, x_self_cond=None):
        """eps = net(x, time, x_self_cond); same call signature as the reference's Unet.forward
        (hicdiff_condition.py:345) / hicedrn_Diff.forward (hicedrn_Diff.py:267)."""
        if self.self_condition and x_self_cond is None:
            raise TypeError("this network was built with self_condition=True: x_self_cond is required")
        return self._eps_plan.eps_forward(x, time, x_self_cond if self.self_condition else None)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_eps_plan"):
            self._eps_plan.invalidate()
        return out


class Unet(_EpsNet):
    """Drop-in for the reference `Unet` (all three files); `noise_level_emb=True` selects the SR3 flavour."""

    def __init__(
        self,
        dim,
        init_dim=None,
        out_dim=None,
        dim_mults: Sequence[int] = (1, 2, 4, 8),
        channels=1,
        self_condition=True,
        resnet_block_groups=8,
        learned_variance=False,
        learned_sinusoidal_cond=False,
        random_fourier_features=False,
        learned_sinusoidal_dim=16,
        noise_level_emb=False,
    ):
        super().__init__()
        if learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError(
                "learned/random sinusoidal time embeddings are rejected by GaussianDiffusion in the reference "
                "(hicdiff_condition.py:448) and are not built here"
            )
        if learned_variance:
            raise NotImplementedError("learned_variance is not used by any reference script and is not built")
        if channels != 1:
            raise NotImplementedError("Hi-C tiles are single-channel; channels != 1 is not built")
        if resnet_block_groups != 8:
            raise NotImplementedError("the fused GroupNorm kernel is built for the reference's 8 groups")
        init_dim = dim if init_dim is None else init_dim
        if init_dim != dim:
            raise NotImplementedError("init_dim != dim is not used by any reference script and is not built")

        self.channels = channels
        self.self_condition = self_condition
        self.noise_level_emb = noise_level_emb
        self.random_or_learned_sinusoidal_cond = False
        self.dim = dim
        self.dim_mults = tuple(dim_mults)
        sr3 = bool(noise_level_emb)
        self._variant = _lib.HD_UNET_SR3 if sr3 else _lib.HD_UNET

        in_ch = channels * (2 if self_condition else 1)
        self.init_conv = nn.Conv2d(in_ch, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        time_dim = dim * 4
        self.time_mlp = nn.Sequential(
            _TimeEmbedding(dim, "noise_level" if sr3 else "sinusoidal"),
            nn.Linear(dim, time_dim),
            nn.GELU(),
            nn.Linear(time_dim, time_dim),
        )

        def res(a, b):
            return _ResBlock(a, b, time_dim, resnet_block_groups, sr3)

        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        n = len(in_out)
        for ind, (d_in, d_out) in enumerate(in_out):
            last = ind >= n - 1
            self.downs.append(nn.ModuleList([
                res(d_in, d_in),
                res(d_in, d_in),
                _Residual(_PreNorm(d_in, _LinAttn(d_in))),
                _resample(nn.Conv2d(d_in * 4, d_out, 1)) if not last else nn.Conv2d(d_in, d_out, 3, padding=1),
            ]))
        mid = dims[-1]
        self.mid_block1 = res(mid, mid)
        self.mid_attn = _Residual(_PreNorm(mid, _Attn(mid)))
        self.mid_block2 = res(mid, mid)
        for ind, (d_in, d_out) in enumerate(reversed(in_out)):
            last = ind == n - 1
            self.ups.append(nn.ModuleList([
                res(d_out + d_in, d_out),
                res(d_out + d_in, d_out),
                _Residual(_PreNorm(d_out, _LinAttn(d_out))),
                _resample(nn.Conv2d(d_out, d_in, 3, padding=1)) if not last else nn.Conv2d(d_out, d_in, 3, padding=1),
            ]))
        self.out_dim = channels if out_dim is None else out_dim
        if self.out_dim != 1:
            raise NotImplementedError("out_dim != 1 is not built")
        self.final_res_block = res(dim * 2, dim)
        self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        self._init_plan_state()

    def _plan_config(self) -> dict:
        return dict(variant=self._variant, self_condition=int(self.self_condition), dim=self.dim,
                    dim_mults=self.dim_mults, num_blocks=0)


class _EdrnBlock(_Holder):  # hicedrn ResnetBlock (hicedrn_Diff.py:182-192 / hicedrn_sr3_Diff.py:245-253)
    def __init__(self, n_feat: int, time_dim: int, sr3: bool):
        super().__init__()
        if sr3:
            self.noise_func = _NoiseFunc(time_dim, n_feat)
        else:
            self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, n_feat * 2))
        self.conv = _EdrnConv(n_feat)
        self.res_conv = nn.Identity()


class _EdrnConv(_Holder):  # hicedrn Block (hicedrn_Diff.py:169-172): key `proj.*`
    def __init__(self, n_feat: int):
        super().__init__()
        self.proj = nn.Conv2d(n_feat, n_feat, 3, padding=1)


class _HiCEDRN(_EpsNet):
    N_FEAT = 256

    def _build(self, channels, out_dim, number_resnet, self_condition, learned_sinusoidal_cond, sr3):
        if learned_sinusoidal_cond:
            raise NotImplementedError("learned sinusoidal time embedding is not built (GaussianDiffusion rejects it)")
        if channels != 1:
            raise NotImplementedError("Hi-C tiles are single-channel; channels != 1 is not built")
        n_feat = self.N_FEAT
        self.channels = channels
        self.self_condition = self_condition
        self.random_or_learned_sinusoidal_cond = False
        self.number_resnet = number_resnet
        self._variant = _lib.HD_HICEDRN_SR3 if sr3 else _lib.HD_HICEDRN
        in_ch = channels * (2 if self_condition else 1)
        self.head = nn.Conv2d(in_ch, n_feat, 3, padding=1)
        time_dim = n_feat * 4
        self.time_mlp = nn.Sequential(
            _TimeEmbedding(n_feat, "noise_level" if sr3 else "sinusoidal"),
            nn.Linear(n_feat, time_dim),
            nn.GELU(),
            nn.Linear(time_dim, time_dim),
        )
        self.body = nn.Sequential(*[_EdrnBlock(n_feat, time_dim, sr3) for _ in range(number_resnet)])
        self.body_tail = nn.Conv2d(n_feat, n_feat, 3, padding=1)
        self.out_dim = channels if out_dim is None else out_dim
        if self.out_dim != 1:
            raise NotImplementedError("out_dim != 1 is not built")
        self.tail = nn.Conv2d(n_feat, self.out_dim, 3, padding=1)
        self._init_plan_state()

    def _plan_config(self) -> dict:
        return dict(variant=self._variant, self_condition=int(self.self_condition), dim=0, dim_mults=(),
                    num_blocks=self.number_resnet)

    def init_params(self):
        """Same re-initialisation helper as the reference (hicedrn_Diff.py:291-297)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight.data, 0.0, 0.02)
        self._eps_plan.invalidate()


class hicedrn_Diff(_HiCEDRN):
    """Drop-in for /root/reference/src/model/hicedrn_Diff.py:210 `hicedrn_Diff`."""

    def __init__(self, channels=1, out_dim=None, number_resnet=32, self_condition=False,
                 learned_sinusoidal_cond=False, learned_sinusoidal_dim=16):
        super().__init__()
        self._build(channels, out_dim, number_resnet, self_condition, learned_sinusoidal_cond, sr3=False)


class hicedrn_sr3_Diff(_HiCEDRN):
    """Drop-in for /root/reference/src/model/hicedrn_sr3_Diff.py:267 `hicedrn_Diff` (noise_level_emb=True)."""

    def __init__(self, channels=1, out_dim=None, number_resnet=32, self_condition=False,
                 learned_sinusoidal_cond=False, noise_level_emb=True, learned_sinusoidal_dim=16):
        super().__init__()
        self.noise_level_emb = noise_level_emb
        self._build(channels, out_dim, number_resnet, self_condition, learned_sinusoidal_cond, sr3=bool(noise_level_emb))
