"""Fused Adam for the training loop (train.py:111 `torch.optim.Adam(diffusion.parameters(), lr=2e-5)`, stepped at :129).

`hicdiff_b200.optim.Adam` takes torch.optim.Adam's arguments and is stepped the same way (`opt.step(); opt.zero_grad()`), but
the whole update is ONE launch of `hd_adam_step` over every parameter tensor (torch's default multi-tensor path is ~35 launches
and ~12 passes over the parameters; here each parameter, gradient and moment is read once and written once).  The arithmetic
follows torch's multi-tensor Adam operation by operation, so a run is interchangeable with the stock optimiser.  The stock
`torch.optim.Adam` keeps working unchanged on the same modules -- this class is an opt-in replacement, not a requirement.

CUDA fp32 parameters only; every parameter of a group must have a gradient when `step()` is called (the training step of this
package produces all of them); no amsgrad / maximize / closures.  There is no CPU path: it raises without the library or off-GPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _Span:
    """A device span exposed through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _device_view(ptr, like):
    if like.numel() == 0:
        return torch.empty_like(like)
    return torch.as_tensor(_Span(ptr, like.numel()), device=like.device).view_as(like)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._handles = {}          # group index -> (hd_adam*, key, params)

    def _handle(self, gi: int, group):
        params = [p for p in group["params"] if p.requires_grad]
        key = tuple((p.data_ptr(), p.numel()) for p in params)
        ent = self._handles.get(gi)
        if ent is not None and ent[1] == key:
            return ent[0], params
        if ent is not None:
            raise RuntimeError("hicdiff_b200.optim.Adam: parameter storage moved after the first step (call .to()/.cuda() on the "
                               "module BEFORE constructing the optimiser, as train.py does)")
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("hicdiff_b200.optim.Adam needs contiguous CUDA fp32 parameters (there is no CPU path)")
        lib = _lib.load()
        ptrs = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        nums = (C.c_int64 * len(params))(*[p.numel() for p in params])
        h = C.c_void_p()
        _lib.check(lib.hd_adam_create(ptrs, nums, len(params), C.byref(h)), "hd_adam_create")
        self._handles[gi] = (h, key, params)
        return h, params

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("hicdiff_b200.optim.Adam does not take a closure")
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            h, params = self._handle(gi, group)
            if not params:
                continue
            grads = []
            for p in params:
                g = p.grad
                if g is None:
                    raise RuntimeError("hicdiff_b200.optim.Adam: a parameter has no gradient (every parameter of the group must)")
                if g.is_sparse or g.dtype != torch.float32 or not g.is_cuda:
                    raise RuntimeError("hicdiff_b200.optim.Adam needs dense CUDA fp32 gradients")
                grads.append(g if g.is_contiguous() else g.contiguous())
            gp = (C.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
            b1, b2 = group["betas"]
            _lib.check(lib.hd_adam_step(h, gp, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                        float(group["weight_decay"]), _lib.stream_ptr()), "hd_adam_step")
        return None

    def moments(self, p):
        """(exp_avg, exp_avg_sq, step) of parameter `p`: zero-copy views of the optimiser's flat state (tests / checkpointing)."""
        lib = _lib.load()
        for h, _, params in self._handles.values():
            for i, q in enumerate(params):
                if q is p:
                    m, v, st = C.c_void_p(), C.c_void_p(), C.c_int64()
                    _lib.check(lib.hd_adam_state(h, i, C.byref(m), C.byref(v), C.byref(st)), "hd_adam_state")
                    return _device_view(m.value, p), _device_view(v.value, p), st.value
        raise KeyError("parameter has no state yet (no step taken)")

    def __del__(self):
        try:
            lib = _lib.load()
            for h, _, _ in self._handles.values():
                lib.hd_adam_destroy(h)
        except Exception:
            pass
        self._handles = {}
