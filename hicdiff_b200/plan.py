"""Host-side owner of one `hd_plan` (include/hicdiff_b200.h): uploads an eps-net's parameters, the diffusion schedule
and drives eps-forward / DDPM-step / full-sampling calls on the current CUDA stream.

Replaces, for the sampling path, what `nn.Module.to(device)` + ATen dispatch do in the reference
(inference.py:59-98).  All compute happens in libhicdiff_b200.so; a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib


# GEMM precision of the sampling path.  The reference computes in fp32 (hicdiff_condition.py:90,105 pick the fp32
# eps); bf16 operands are the fast default.  "bf16w2" keeps every conv weight as a hi + lo bf16 pair (two MMAs per product,
# ~16 mantissa bits on the weights; activations stay bf16): rounding the WEIGHTS is a fixed perturbation of the model that a
# 1000-step chain accumulates coherently, rounding activations is not (scripts/precision_study.py).  Select per net with
# `net.precision = "bf16w2"` or process-wide with HICDIFF_B200_PRECISION.
PRECISIONS = {"bf16": 0, "bf16w2": 1}


def precision_of(net) -> str:
    p = getattr(net, "precision", None) or os.environ.get("HICDIFF_B200_PRECISION", "bf16")
    if p not in PRECISIONS:
        raise ValueError(f"unknown precision {p!r}: choose one of {sorted(PRECISIONS)}")
    return p


def _params_version(module: torch.nn.Module) -> int:
    v = 0
    for p in module.parameters():
        v += p._version + (hash(p.data_ptr()) & 0xFFFF)
    return v


class EpsPlan:
    """Lazy, self-invalidating binding between a parameter-holder net (nets.py) and a device plan."""

    def __init__(self, net: torch.nn.Module):
        self._net = net
        self._handle: Optional[int] = None
        self._key = None
        self._schedule = None      # dict of fp32 [T] CPU/GPU tensors + 'time_values'
        self._schedule_id = None
        self.debug_keep = False

    # ------------------------------------------------------------------ life cycle
    def invalidate(self):
        self._key = None

    def destroy(self):
        if self._handle is not None:
            _lib.load().hd_plan_destroy(self._handle)
            self._handle = None
            self._key = None

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass

    def set_schedule(self, sqrt_recip, sqrt_recipm1, coef1, coef2, log_var, time_values):
        """Schedule tables (fp32 [T]) of the owning GaussianDiffusion; sigma = exp(0.5 * log_var) is evaluated with
        torch so the table is bit-identical to what p_sample computes (hicdiff_condition.py:597)."""
        # evaluated on the CPU so the table does not depend on the device's exp() (the CPU oracle is the referee)
        sigma = (0.5 * log_var.detach().to("cpu", torch.float32)).exp()
        sched = dict(sqrt_recip=sqrt_recip, sqrt_recipm1=sqrt_recipm1, coef1=coef1, coef2=coef2,
                     sigma=sigma, time_values=time_values)
        self._schedule = {k: v.detach().to("cpu", torch.float32) for k, v in sched.items()}
        self._schedule_id = tuple(int(v.data_ptr()) for v in (sqrt_recip, coef1)) + (int(sqrt_recip.numel()),)
        self.invalidate()

    def _device(self) -> torch.device:
        return next(self._net.parameters()).device

    def _ensure(self) -> int:
        lib = _lib.load()
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                "hicdiff_b200 runs on sm_100a GPUs only: move the module to a CUDA device (there is no CPU fallback)"
            )
        if self._schedule is None:
            # stand-alone eps-net use (no GaussianDiffusion attached): a 1-step identity schedule keeps the plan valid
            one = torch.ones(1)
            self.set_schedule(one, one * 0, one, one * 0, one * 0, one * 0)
        T = int(self._schedule["sqrt_recip"].numel())
        prec = precision_of(self._net)
        # kernel-form option bits of hd_config.reserved[0] (include/hicdiff_b200.h; the precision bits 4-5 belong to `precision`):
        # `net.plan_options = 0x80` or HICDIFF_B200_PLAN_OPTIONS=0x80 -- read per plan, used by the parity tests for the opt-in forms
        po = getattr(self._net, "plan_options", None)
        opts = int(po) if po is not None else int(os.environ.get("HICDIFF_B200_PLAN_OPTIONS", "0"), 0)
        opts &= ~0x30
        key = (str(dev), _params_version(self._net), self._schedule_id, T, self.debug_keep, prec, opts)
        if self._handle is not None and key == self._key:
            return self._handle
        with torch.cuda.device(dev):
            cfgd = self._net._plan_config()
            need_new = self._handle is None or self._key is None or self._key[0] != key[0] or self._key[3:] != key[3:]
            if need_new:
                self.destroy()
                cfg = _lib.hd_config()
                cfg.abi_version = _lib.HD_ABI_VERSION
                cfg.variant = cfgd["variant"]
                cfg.self_condition = cfgd["self_condition"]
                cfg.dim = cfgd["dim"]
                cfg.num_mults = len(cfgd["dim_mults"])
                for i, m in enumerate(cfgd["dim_mults"]):
                    cfg.dim_mults[i] = int(m)
                cfg.image_size = 64
                cfg.timesteps = T
                cfg.num_blocks = cfgd["num_blocks"]
                cfg.debug_keep = 1 if self.debug_keep else 0
                cfg.reserved[0] = (PRECISIONS[prec] << 4) | opts  # bits 4-5: precision (include/hicdiff_b200.h)
                h = C.c_void_p()
                _lib.check(lib.hd_plan_create(C.byref(cfg), C.byref(h)), "hd_plan_create")
                self._handle = h.value
            stream = _lib.stream_ptr()
            keep = []
            for name, t in self._net.state_dict().items():
                if not torch.is_floating_point(t):
                    continue
                t32 = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                keep.append(t32)
                shape = (C.c_int64 * max(t32.dim(), 1))(*t32.shape)
                _lib.check(lib.hd_plan_set_weight(self._handle, name.encode(), t32.data_ptr(), shape, t32.dim(), stream),
                           f"hd_plan_set_weight({name})")
            s = {k: v.to(dev).contiguous() for k, v in self._schedule.items()}
            _lib.check(lib.hd_plan_set_schedule(self._handle, s["sqrt_recip"].data_ptr(), s["sqrt_recipm1"].data_ptr(),
                                                s["coef1"].data_ptr(), s["coef2"].data_ptr(), s["sigma"].data_ptr(),
                                                s["time_values"].data_ptr(), T, stream), "hd_plan_set_schedule")
            _lib.check(lib.hd_plan_finalize(self._handle, stream), "hd_plan_finalize")
            del keep
        self._key = key
        return self._handle

    # ------------------------------------------------------------------ calls
    @staticmethod
    def _tiles(t: torch.Tensor, dev, name: str) -> torch.Tensor:
        if t.dim() != 4 or t.shape[1] != 1 or t.shape[2] != 64 or t.shape[3] != 64:
            raise ValueError(f"{name} must be [B, 1, 64, 64] (got {tuple(t.shape)})")
        if t.device != dev:
            raise RuntimeError(f"{name} is on {t.device}, the module is on {dev}")
        return t.detach().to(torch.float32).contiguous()

    def eps_forward(self, x, time, cond=None):
        lib = _lib.load()
        h = self._ensure()
        dev = self._device()
        x = self._tiles(x, dev, "x")
        B = x.shape[0]
        cond = self._tiles(cond, dev, "x_self_cond") if cond is not None else None
        tv = time.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        if tv.numel() != B:
            raise ValueError(f"time must have one entry per sample ({B}), got {tuple(time.shape)}")
        eps = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(lib.hd_eps_forward(h, x.data_ptr(), _lib.ptr(cond), tv.data_ptr(), eps.data_ptr(), B,
                                          _lib.stream_ptr()), "hd_eps_forward")
        return eps

    def ddpm_step(self, x, eps, t: int, noise=None, seed: int = 0, tile_offset: int = 0, want_x0: bool = False):
        """In-place-free wrapper: returns (x_{t-1}, x_start or None)."""
        lib = _lib.load()
        h = self._ensure()
        dev = self._device()
        x = self._tiles(x, dev, "x").clone()
        eps = self._tiles(eps, dev, "eps")
        noise = self._tiles(noise, dev, "noise") if noise is not None else None
        x0 = torch.empty_like(x) if want_x0 else None
        with torch.cuda.device(dev):
            _lib.check(lib.hd_ddpm_step(h, x.data_ptr(), eps.data_ptr(), _lib.ptr(noise), _lib.ptr(x0), int(t), x.shape[0],
                                        int(seed), int(tile_offset), _lib.stream_ptr()), "hd_ddpm_step")
        return x, x0

    def sample(self, batch: int, cond=None, noise=None, seed: int = 0, tile_offset: int = 0, trace: bool = False,
               t_start: Optional[int] = None, t_end: int = 0, init=None):
        """Full reverse chain on the device.  noise: optional [T,B,1,64,64] injected draws (parity mode)."""
        lib = _lib.load()
        h = self._ensure()
        dev = self._device()
        T = int(self._schedule["sqrt_recip"].numel())
        t_start = T - 1 if t_start is None else int(t_start)
        cond = self._tiles(cond, dev, "cond") if cond is not None else None
        if noise is not None:
            if noise.shape != (T, batch, 1, 64, 64):
                raise ValueError(f"noise must be [T={T}, B={batch}, 1, 64, 64], got {tuple(noise.shape)}")
            if noise.device != dev:
                raise RuntimeError("noise must live on the module's device")
            noise = noise.detach().to(torch.float32).contiguous()
        out = torch.empty(batch, 1, 64, 64, device=dev, dtype=torch.float32)
        if init is not None:
            out.copy_(self._tiles(init, dev, "init"))
        elif t_start != T - 1:
            raise ValueError("a partial chain (t_start < T-1) needs `init`")
        tr = torch.empty(t_start - t_end + 1, batch, 1, 64, 64, device=dev, dtype=torch.float32) if trace else None
        with torch.cuda.device(dev):
            _lib.check(lib.hd_sample(h, _lib.ptr(cond), _lib.ptr(noise), out.data_ptr(), _lib.ptr(tr), batch, int(seed),
                                     int(tile_offset), t_start, int(t_end), _lib.stream_ptr()), "hd_sample")
        return (out, tr) if trace else out

    # ------------------------------------------------------------------ introspection
    def launches_per_step(self, batch: int):
        lib = _lib.load()
        h = self._ensure()
        a, b = C.c_int32(), C.c_int32()
        _lib.check(lib.hd_plan_launches_per_step(h, batch, C.byref(a), C.byref(b)), "hd_plan_launches_per_step")
        return a.value, b.value

    def profile_step(self, batch: int, reps: int = 10):
        """Per-launch average ms + algorithmic FLOPs/bytes of one sampling step (list of dicts)."""
        import json

        lib = _lib.load()
        h = self._ensure()
        buf = C.create_string_buffer(1 << 18)
        with torch.cuda.device(self._device()):
            _lib.check(lib.hd_plan_profile_step(h, batch, reps, buf, len(buf), _lib.stream_ptr()), "hd_plan_profile_step")
        return json.loads(buf.value.decode())

    def device_bytes(self) -> int:
        return int(_lib.load().hd_plan_device_bytes(self._ensure()))

    def debug_names(self, batch: int):
        lib = _lib.load()
        buf = C.create_string_buffer(1 << 16)
        _lib.check(lib.hd_debug_names(self._ensure(), batch, buf, len(buf)), "hd_debug_names")
        return [n for n in buf.value.decode().split("\n") if n]

    def debug_read(self, batch: int, name: str) -> torch.Tensor:
        lib = _lib.load()
        h = self._ensure()
        n = C.c_int64()
        shape = (C.c_int32 * 4)()
        _lib.check(lib.hd_debug_read(h, batch, name.encode(), None, C.byref(n), shape, _lib.stream_ptr()), "hd_debug_read")
        out = torch.empty(tuple(shape), device=self._device(), dtype=torch.float32)
        _lib.check(lib.hd_debug_read(h, batch, name.encode(), out.data_ptr(), C.byref(n), shape, _lib.stream_ptr()),
                   "hd_debug_read")
        return out
