"""`GaussianDiffusion` with the reference's constructor, buffers, methods and return conventions; the reverse
process itself (p_sample_loop and everything under it) runs as CUDA graphs inside libhicdiff_b200.so.

Three flavours, one per reference file:
  conditional   /root/reference/src/hicdiff_condition.py:429-750   (`super_resolution`, list-returning trace)
  unconditional /root/reference/src/hicdiff.py:432-755             (`sample`, stacked trace, p_losses(x_start, t))
  SR3           /root/reference/src/hicdiff_sr3.py:491-796          (noise-level conditioning, numpy RNG in p_losses)

Noise: by default x_T and the per-step z come from an in-kernel Philox stream (seeded from torch's CUDA generator, so
`torch.manual_seed` still controls reproducibility); passing `noise=[T,B,1,64,64]` injects the exact draws the
reference would make (draw 0 = x_T, draw i = z of step T-i, none at t == 0) for parity runs.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start"])

_BUFFERS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2", "p2_loss_weight",
)


def _betas(kind: str, T: int, **kw) -> torch.Tensor:
    """float64 beta schedules (hicdiff_condition.py:393-427)."""
    f64 = torch.float64
    if kind == "linear":
        s = 1000 / T
        return torch.linspace(s * 0.0001, s * 0.02, T, dtype=f64)
    grid = torch.linspace(0, T, T + 1, dtype=f64) / T
    if kind == "cosine":
        s = kw.get("s", 0.008)
        abar = torch.cos((grid + s) / (1 + s) * math.pi * 0.5) ** 2
    elif kind == "sigmoid":
        start, end, tau = kw.get("start", -3), kw.get("end", 3), kw.get("tau", 1)
        lo, hi = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
        abar = (-((grid * (end - start) + start) / tau).sigmoid() + hi) / (hi - lo)
    else:
        raise ValueError(f"unknown beta schedule {kind}")
    abar = abar / abar[0]
    return torch.clip(1 - (abar[1:] / abar[:-1]), 0, 0.999)


class _LossNeedsBackward(torch.autograd.Function):
    """Makes the training loss a graph leaf whose backward explains what is (not) built yet."""

    @staticmethod
    def forward(ctx, loss, *params):
        return loss.clone()

    @staticmethod
    def backward(ctx, g):  # pragma: no cover - exercised only on GPU
        raise NotImplementedError(
            "hicdiff_b200: the backward of this eps-net (Unet / SR3 variants) is not built yet -- SURVEY.md 8(f) N2. "
            "hicedrn_Diff, the model train.py trains, has a full backward (hicdiff_b200/train.py); for the others the "
            "forward loss value is exact and training stays with the reference."
        )


class _GaussianDiffusionBase(nn.Module):
    _flavour = "conditional"
    _default_loss, _default_schedule = "l1", "sigmoid"

    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, loss_type=None,
                 objective="pred_noise", beta_schedule=None, schedule_fn_kwargs=dict(), p2_loss_weight_gamma=0.0,
                 p2_loss_weight_k=1, ddim_sampling_eta=0.0, auto_normalize=False):
        super().__init__()
        if model.channels != model.out_dim:
            raise AssertionError("model.channels must equal model.out_dim")
        if model.random_or_learned_sinusoidal_cond:
            raise AssertionError("random/learned sinusoidal conditioning is not supported")
        loss_type = self._default_loss if loss_type is None else loss_type
        beta_schedule = self._default_schedule if beta_schedule is None else beta_schedule
        if objective not in {"pred_noise", "pred_x0", "pred_v"}:
            raise AssertionError("objective must be pred_noise, pred_x0 or pred_v")
        if objective != "pred_noise" and self._flavour == "sr3":
            raise NotImplementedError("objective != 'pred_noise' is built for the conditional / unconditional flavours only")
        if image_size != 64:
            raise NotImplementedError("the sm_100a path is built for the reference's 64x64 tiles (image_size=64)")

        self.model = model
        self.channels = model.channels
        self.self_condition = model.self_condition
        self.image_size = image_size
        self.objective = objective
        self.loss_type = loss_type

        betas = _betas(beta_schedule, timesteps, **schedule_fn_kwargs)
        alphas = 1.0 - betas
        abar = torch.cumprod(alphas, dim=0)
        abar_prev = F.pad(abar[:-1], (1, 0), value=1.0)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = timesteps if sampling_timesteps is None else sampling_timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
        vals = {
            "betas": betas,
            "alphas_cumprod": abar,
            "alphas_cumprod_prev": abar_prev,
            "sqrt_alphas_cumprod": torch.sqrt(abar),
            "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
            "log_one_minus_alphas_cumprod": torch.log(1.0 - abar),
            "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / abar),
            "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / abar - 1),
            "posterior_variance": post_var,
            "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
            "posterior_mean_coef1": betas * torch.sqrt(abar_prev) / (1.0 - abar),
            "posterior_mean_coef2": (1.0 - abar_prev) * torch.sqrt(alphas) / (1.0 - abar),
            "p2_loss_weight": (p2_loss_weight_k + abar / (1 - abar)) ** -p2_loss_weight_gamma,
        }
        for name in _BUFFERS:  # registration order == reference state_dict order
            self.register_buffer(name, vals[name].to(torch.float32))
        if self._flavour == "sr3":
            # plain float64 attribute, NOT a buffer (hicdiff_sr3.py:536): [1, 1, sqrt(abar_0), ..., sqrt(abar_{T-2})]
            self.sqrt_alphas_cumprod_prev = torch.sqrt(F.pad(abar_prev, (1, 0), value=1.0))
        self.auto_normalize = auto_normalize
        self._sample_calls = 0

    # ------------------------------------------------------------------ helpers
    def normalize(self, img):
        if not self.auto_normalize:
            return img
        if isinstance(img, (list, tuple)):
            return type(img)(i * 2 - 1 for i in img)
        return img * 2 - 1

    def unnormalize(self, t):
        if not self.auto_normalize:
            return t
        if isinstance(t, list):
            return [(i + 1) * 0.5 for i in t]
        return (t + 1) * 0.5

    def _time_values(self) -> torch.Tensor:
        """What the eps-net sees as `time` at step t."""
        T = self.num_timesteps
        if self._flavour == "sr3":
            # torch.FloatTensor([self.sqrt_alphas_cumprod_prev[t + 1]]) -- float64 -> float32 (hicdiff_sr3.py:636)
            return self.sqrt_alphas_cumprod_prev[1:T + 1].to(torch.float32)
        return torch.arange(T, dtype=torch.float32)

    def _x_start_coefficients(self):
        """(a, b) with x_start = a[t] * x_t - b[t] * model_output -- the one place the objective enters the reverse step
        (model_predictions :559-579): pred_noise -> predict_start_from_noise (:526-530), pred_v -> predict_start_from_v
        (:545-549), pred_x0 -> the output itself (0 * x - (-1) * out, exact in fp32)."""
        if self.objective == "pred_noise":
            return self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod
        cached = getattr(self, "_obj_cols", None)
        if cached is None or cached[0].device != self.betas.device:
            if self.objective == "pred_v":
                cached = (self.sqrt_alphas_cumprod.detach().clone(), self.sqrt_one_minus_alphas_cumprod.detach().clone())
            else:
                cached = (torch.zeros_like(self.betas), -torch.ones_like(self.betas))
            object.__setattr__(self, "_obj_cols", cached)      # plain attributes (stable pointers), not buffers / state_dict entries
        return cached

    def _sync_plan(self):
        plan = self.model.eps_plan
        a, b = self._x_start_coefficients()
        sid = (a.data_ptr(), self.posterior_mean_coef1.data_ptr(), self.num_timesteps)
        if plan._schedule_id != sid:
            plan.set_schedule(a, b, self.posterior_mean_coef1, self.posterior_mean_coef2,
                              self.posterior_log_variance_clipped, self._time_values())
        return plan

    def _next_seed(self, device) -> int:
        # derive the Philox seed from torch's generator so torch.manual_seed governs reproducibility
        return int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())

    def _eps_time(self, t_int: int, b: int, device):
        if self._flavour == "sr3":
            return self._time_values()[t_int].repeat(b, 1).to(device)
        return torch.full((b,), t_int, device=device, dtype=torch.long)

    # ------------------------------------------------------------------ reference API: single-step pieces
    def predict_start_from_noise(self, x_t, t, noise):
        if isinstance(t, int):
            return self.sqrt_recip_alphas_cumprod[t] * x_t - self.sqrt_recipm1_alphas_cumprod[t] * noise
        return _gather(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t - _gather(self.sqrt_recipm1_alphas_cumprod, t, x_t) * noise

    def q_posterior(self, x_start, x_t, t):
        if isinstance(t, int):
            return (self.posterior_mean_coef1[t] * x_start + self.posterior_mean_coef2[t] * x_t,
                    self.posterior_variance[t], self.posterior_log_variance_clipped[t])
        mean = _gather(self.posterior_mean_coef1, t, x_t) * x_start + _gather(self.posterior_mean_coef2, t, x_t) * x_t
        return mean, _gather(self.posterior_variance, t, x_t), _gather(self.posterior_log_variance_clipped, t, x_t)

    @torch.no_grad()
    def model_predictions(self, x, t, x_self_cond=None, clip_x_start=False, t_real=None):
        out = self.model(x, t, x_self_cond)
        tt = t if t_real is None else t_real
        clip = (lambda v: v.clamp(-1.0, 1.0)) if clip_x_start else (lambda v: v)
        if self.objective == "pred_noise":
            return ModelPrediction(out, clip(self.predict_start_from_noise(x, tt, out)))
        x_start = clip(out if self.objective == "pred_x0" else self.predict_start_from_v(x, tt, out))
        return ModelPrediction(self.predict_noise_from_start(x, tt, x_start), x_start)

    def predict_noise_from_start(self, x_t, t, x0):  # :532-536
        return (_gather(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t - x0) / _gather(self.sqrt_recipm1_alphas_cumprod, t, x_t)

    def predict_v(self, x_start, t, noise):  # :538-542
        return _gather(self.sqrt_alphas_cumprod, t, x_start) * noise - _gather(self.sqrt_one_minus_alphas_cumprod, t, x_start) * x_start

    def predict_start_from_v(self, x_t, t, v):  # :544-548
        return _gather(self.sqrt_alphas_cumprod, t, x_t) * x_t - _gather(self.sqrt_one_minus_alphas_cumprod, t, x_t) * v

    def _loss_target(self, x_start, t, noise):  # p_losses :731-740
        if self.objective == "pred_noise":
            return noise
        return x_start if self.objective == "pred_x0" else self.predict_v(x_start, t, noise)

    @torch.no_grad()
    def p_mean_variance(self, x, t, x_self_cond=None, clip_denoised=True):
        """(model_mean, posterior_variance, posterior_log_variance, x_start) for a batch of timesteps t [B] (long) -- the
        reference's helper (hicdiff_condition.py:581-589); the eps-net runs on the device plan, the four results are composed
        from the schedule buffers like the reference does."""
        preds = self.model_predictions(x, t, x_self_cond)
        x_start = preds.pred_x_start
        if clip_denoised:
            x_start = x_start.clamp(-1.0, 1.0)
        mean, var, logvar = self.q_posterior(x_start=x_start, x_t=x, t=t)
        return mean, var, logvar, x_start

    @torch.no_grad()
    def interpolate(self, x1, x2, t=None, lam=0.5, noise=None):
        """hicdiff_condition.py:680-696: diffuse both inputs to step t, mix them, run the reverse chain from t - 1 down to 0
        without conditioning (so, like the reference's, only for eps-nets built with self_condition=False)."""
        if self.self_condition:
            raise NotImplementedError("interpolate is unconditional (self_cond = None in the reference): it needs self_condition=False")
        assert x1.shape == x2.shape
        b = x1.shape[0]
        t = self.num_timesteps - 1 if t is None else int(t)
        tb = torch.full((b,), t, device=x1.device, dtype=torch.long)
        img = (1 - lam) * self.q_sample(x1, tb) + lam * self.q_sample(x2, tb)
        if t == 0:
            return img
        plan = self._sync_plan()
        seed = 0 if noise is not None else self._next_seed(None)
        return plan.sample(b, cond=None, noise=noise, seed=seed, t_start=t - 1, t_end=0, init=img)

    @torch.no_grad()
    def p_sample(self, x, t: int, x_self_cond=None, noise=None):
        """One reverse step (hicdiff_condition.py:591-598): returns (x_{t-1}, x_start).  `noise` injects z."""
        t = int(t)
        plan = self._sync_plan()
        eps = self.model(x, self._eps_time(t, x.shape[0], x.device), x_self_cond)
        return plan.ddpm_step(x, eps, t, noise=noise, seed=self._next_seed(x.device) if noise is None else 0,
                              want_x0=True)

    # ------------------------------------------------------------------ reference API: loops
    def _run_chain(self, batch, cond, noise, return_all):
        plan = self._sync_plan()
        seed = 0 if noise is not None else self._next_seed(None)
        self._last_seed = seed
        return plan.sample(batch, cond=cond, noise=noise, seed=seed, trace=return_all)

    def _x_T(self, batch, noise, device):
        """The chain's starting point: injected draw 0, or the Philox stream 0 of the seed just used."""
        if noise is not None:
            return noise[0]
        from .ops import philox_normal

        return philox_normal(batch, self._last_seed, 0, device=device)

    @torch.no_grad()
    def p_sample_loop(self, x_in, return_all_timesteps=False, noise=None):
        if self._flavour == "unconditional" or not self.self_condition:
            shape = tuple(x_in)
            if len(shape) != 4 or shape[1:] != (self.channels, self.image_size, self.image_size):
                raise ValueError(f"shape must be (B, {self.channels}, {self.image_size}, {self.image_size})")
            res = self._run_chain(shape[0], None, noise, return_all_timesteps)
            if not return_all_timesteps:
                return self.unnormalize(res)
            img, trace = res
            first = self._x_T(shape[0], noise, img.device)
            if self._flavour == "unconditional":
                # hicdiff.py:617 stacks [x_T, x_{T-1}, ..., x_0] on dim 1
                return self.unnormalize(torch.cat((first[:, None], trace.transpose(0, 1)), dim=1))
            return self.unnormalize([first] + list(trace.unbind(0)))
        cond = x_in
        res = self._run_chain(cond.shape[0], cond, noise, return_all_timesteps)
        if not return_all_timesteps:
            return self.unnormalize(res)
        _, trace = res
        return self.unnormalize([cond] + list(trace.unbind(0)))  # hicdiff_condition.py:607,617,620

    @torch.no_grad()
    def ddim_sample(self, shape, return_all_timesteps=False, noise=None):
        """DDIM sampling (hicdiff.py:623-664, hicdiff_condition.py:625-668): `sampling_timesteps` strided steps, one eps-net
        call + one `hd_ddim_step` each.  `noise`: optional list / tensor of the draws the reference would make (x_T, then one z
        per step except the last).  Like the reference, this only works for eps-nets without self-conditioning (there the
        reference feeds its own x_start -- None on the first step -- to torch.cat)."""
        from . import _lib

        if self.self_condition:
            raise NotImplementedError("ddim_sample needs self_condition=False (the reference's version cannot run with it either)")
        if self.objective != "pred_noise":
            raise NotImplementedError("ddim_sample is built for objective='pred_noise'")
        if self._flavour == "sr3":
            raise NotImplementedError("the SR3 flavour has no DDIM sampler in the reference")
        batch = shape[0]
        dev = self.betas.device
        if dev.type != "cuda":
            raise RuntimeError("hicdiff_b200 runs on sm_100a GPUs only (there is no CPU fallback)")
        T, S, eta = self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta
        times = list(reversed(torch.linspace(-1, T - 1, steps=S + 1).int().tolist()))
        pairs = list(zip(times[:-1], times[1:]))
        draws = iter(noise) if noise is not None else None
        seed = 0 if noise is not None else self._next_seed(None)
        if draws is not None:
            img = next(draws).to(dev, torch.float32).contiguous().clone()
        else:
            from .ops import philox_normal

            img = philox_normal(batch, seed, 0, device=dev)
        imgs = [img.clone()] if return_all_timesteps else None
        lib = _lib.load()
        ac = self.alphas_cumprod.detach().cpu()
        sr, srm1 = self.sqrt_recip_alphas_cumprod.detach().cpu(), self.sqrt_recipm1_alphas_cumprod.detach().cpu()
        for k, (t, t_next) in enumerate(pairs):
            eps = self.model(img, torch.full((batch,), t, device=dev, dtype=torch.long), None)
            last = t_next < 0
            if last:
                san = c = sigma = 0.0
            else:
                alpha, alpha_next = ac[t], ac[t_next]                          # the reference's fp32 expressions :648-652
                sg = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                c = float((1 - alpha_next - sg ** 2).sqrt())
                san, sigma = float(alpha_next.sqrt()), float(sg)
            z = None
            if not last and draws is not None:
                z = next(draws).to(dev, torch.float32).contiguous()
            with torch.cuda.device(dev):
                _lib.check(lib.hd_ddim_step(img.data_ptr(), eps.data_ptr(), _lib.ptr(z), None, float(sr[t]), float(srm1[t]), san, c, sigma,
                                            1 if last else 0, img.numel(), int(seed), 0, k, _lib.stream_ptr()), "hd_ddim_step")
            if imgs is not None:
                imgs.append(img.clone())
        ret = img if not return_all_timesteps else torch.stack(imgs, dim=1)
        return self.unnormalize(ret)

    @torch.no_grad()
    def sample(self, x, return_all_timesteps=False, noise=None):
        b = x.shape[0]
        shape = (b, self.channels, self.image_size, self.image_size)
        if self.is_ddim_sampling:                                   # hicdiff.py:670-673
            return self.ddim_sample(shape, return_all_timesteps=return_all_timesteps, noise=noise)
        return self.p_sample_loop(shape, return_all_timesteps=return_all_timesteps, noise=noise)

    @torch.no_grad()
    def super_resolution(self, x_in, continous=False, noise=None):
        return self.p_sample_loop(x_in, continous, noise=noise)

    # ------------------------------------------------------------------ training objective
    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return _gather(self.sqrt_alphas_cumprod, t, x_start) * x_start + \
            _gather(self.sqrt_one_minus_alphas_cumprod, t, x_start) * noise

    @property
    def loss_fn(self):
        if self.loss_type == "l1":
            return F.l1_loss
        if self.loss_type == "l2":
            return F.mse_loss
        raise ValueError(f"invalid loss type {self.loss_type}")

    def _finish_loss(self, loss):
        params = [p for p in self.model.parameters() if p.requires_grad]
        if torch.is_grad_enabled() and params:
            return _LossNeedsBackward.apply(loss, *params)
        return loss

    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters())

    def _trained_loss(self, x, t, cond, noise):
        """Forward + loss + backward in one device call when the eps-net has a backward (hicedrn_Diff); else None."""
        from . import train as _train

        if not (self._wants_grad() and _train.supports_training(self.model)):
            return None
        weight = self.p2_loss_weight.gather(-1, t)
        return _train.training_loss(self.model, x, t.to(torch.float32), cond, noise, weight, self.loss_type)

    def p_losses(self, x_in, t=None, noise=None):
        noisy, clean = x_in
        b, c, h, w = clean.shape
        assert h == self.image_size and w == self.image_size, f"height and width of image must be {self.image_size}"
        if t is None:
            t = torch.randint(0, self.num_timesteps, (b,), device=clean.device).long()
        noise = torch.randn_like(clean) if noise is None else noise
        x = self.q_sample(clean, t, noise)
        # reference quirk kept (hicdiff_condition.py:716,733-737): the pred_x0 / pred_v targets are built from `x_start`, which in
        # the conditional p_losses is the NOISY input of the pair, not the clean tile that was diffused
        target = self._loss_target(noisy, t, noise)
        trained = self._trained_loss(x, t, noisy if self.self_condition else None, target)
        if trained is not None:
            return trained
        with torch.no_grad():
            out = self.model(x, t, noisy if self.self_condition else None)
        loss = self.loss_fn(out, target, reduction="none").reshape(b, -1)  # 'b ... -> b (...)' keeps all elements
        loss = loss * _gather(self.p2_loss_weight, t, loss)
        return self._finish_loss(loss.mean())

    def forward(self, img, *args, **kwargs):
        return self.p_losses(self.normalize(img), *args, **kwargs)


def _gather(table, t, like):
    out = table.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (like.dim() - 1)))


class GaussianDiffusionCond(_GaussianDiffusionBase):
    _flavour = "conditional"
    _default_loss, _default_schedule = "l1", "sigmoid"


class GaussianDiffusionUncond(_GaussianDiffusionBase):
    _flavour = "unconditional"
    _default_loss, _default_schedule = "l1", "sigmoid"

    def __init__(self, model, **kw):
        if model.self_condition:
            # torch.cat((None, x)) on the first step (hicdiff.py:352,613): the reference cannot run this either
            raise NotImplementedError("src/hicdiff.py only works with self_condition=False (SURVEY.md C.12)")
        super().__init__(model, **kw)

    def p_losses(self, x_start, t, noise=None):  # hicdiff.py:711-747
        b = x_start.shape[0]
        noise = torch.randn_like(x_start) if noise is None else noise
        x = self.q_sample(x_start, t, noise)
        target = self._loss_target(x_start, t, noise)
        trained = self._trained_loss(x, t, None, target)
        if trained is not None:
            return trained
        with torch.no_grad():
            out = self.model(x, t, None)
        loss = self.loss_fn(out, target, reduction="none").reshape(b, -1)  # 'b ... -> b (...)' keeps all elements
        loss = loss * _gather(self.p2_loss_weight, t, loss)
        return self._finish_loss(loss.mean())

    def forward(self, img, *args, **kwargs):  # hicdiff.py:749-755
        b, c, h, w = img.shape
        assert h == self.image_size and w == self.image_size, f"height and width of image must be {self.image_size}"
        t = torch.randint(0, self.num_timesteps, (b,), device=img.device).long()
        return self.p_losses(self.normalize(img), t, *args, **kwargs)


class GaussianDiffusionSR3(_GaussianDiffusionBase):
    _flavour = "sr3"
    _default_loss, _default_schedule = "l2", "linear"

    def q_sample(self, x_start, continuous_sqrt_alpha_cumprod, noise=None):  # hicdiff_sr3.py:735-739
        noise = torch.randn_like(x_start) if noise is None else noise
        return continuous_sqrt_alpha_cumprod * x_start + (1 - continuous_sqrt_alpha_cumprod ** 2).sqrt() * noise

    def p_losses(self, x_in, t=None, noise=None, level=None):  # hicdiff_sr3.py:750-792 (numpy global RNG, no p2 weight; `t` is ignored there too)
        noisy, clean = x_in
        b, c, h, w = clean.shape
        assert h == self.image_size and w == self.image_size, f"height and width of image must be {self.image_size}"
        if level is None:
            t = np.random.randint(1, self.num_timesteps + 1)
            level = torch.FloatTensor(np.random.uniform(self.sqrt_alphas_cumprod_prev[t - 1],
                                                        self.sqrt_alphas_cumprod_prev[t], size=b)).to(clean.device)
        level = level.view(b, -1)
        noise = torch.randn_like(clean) if noise is None else noise
        x = self.q_sample(clean, level.view(-1, 1, 1, 1), noise)
        from . import train as _train

        if self._wants_grad() and _train.supports_training(self.model):      # hicedrn_sr3_Diff: forward + loss + backward on the device
            return _train.training_loss(self.model, x, level.reshape(-1), noisy if self.self_condition else None, noise,
                                        torch.ones(b, device=clean.device), self.loss_type)
        with torch.no_grad():
            out = self.model(x, level, noisy if self.self_condition else None)
        return self._finish_loss(self.loss_fn(out, noise, reduction="none").mean())
