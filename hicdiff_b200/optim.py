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


def _bump_versions(params):
    try:
        torch.autograd.graph.increment_version(params)          # torch >= 2.4: accepts an iterable
    except TypeError:
        for p in params:
            torch.autograd.graph.increment_version(p)


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
        self._pending = {}          # group index -> state restored by load_state_dict before the handle exists

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
        if len({p.device for p in params}) > 1:
            raise RuntimeError("hicdiff_b200.optim.Adam: all parameters of a group must live on one device")
        lib = _lib.load()
        ptrs = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        nums = (C.c_int64 * len(params))(*[p.numel() for p in params])
        h = C.c_void_p()
        with self._on(params):
            _lib.check(lib.hd_adam_create(ptrs, nums, len(params), C.byref(h)), "hd_adam_create")
        self._handles[gi] = (h, key, params)
        pending = self._pending.pop(gi, None)
        if pending is not None:
            self._restore(gi, pending)
        return h, params

    @staticmethod
    def _on(params):
        """Device guard: the state, the launch and the stream all belong to the parameters' device."""
        import contextlib

        return torch.cuda.device(params[0].device) if params else contextlib.nullcontext()

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
            with self._on(params):
                _lib.check(lib.hd_adam_step(h, gp, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                            float(group["weight_decay"]), _lib.stream_ptr()), "hd_adam_step")
            # The kernel wrote through raw pointers: tell torch.  Everything keyed on `Tensor._version` -- the sampling plans'
            # weight cache (plan.py::_params_version), autograd's saved-tensor checks -- must see an in-place update exactly as
            # after torch.optim.Adam's `param.addcdiv_()`.
            _bump_versions(params)
        return None

    # ------------------------------------------------------------------ checkpointing (torch.optim.Adam's layout)
    def state_dict(self):
        """Same structure as `torch.optim.Adam.state_dict()` (per-parameter `step`, `exp_avg`, `exp_avg_sq`, indexed in
        param_groups order), so a checkpoint moves between the two optimisers."""
        state, groups, idx = {}, [], 0
        for gi, group in enumerate(self.param_groups):
            ids = list(range(idx, idx + len(group["params"])))
            idx += len(group["params"])
            groups.append({**{k: v for k, v in group.items() if k != "params"}, "params": ids})
            ent = self._handles.get(gi)
            for pid, p in zip(ids, group["params"]):
                if ent is not None and any(q is p for q in ent[2]):
                    m, v, st = self.moments(p)
                    state[pid] = {"step": torch.tensor(float(st)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
                elif gi in self._pending and pid in self._pending[gi]:
                    state[pid] = self._pending[gi][pid]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(g["params"]) != len(s["params"]) for g, s in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        for g, mine in zip(groups, self.param_groups):
            for k, v in g.items():
                if k != "params":
                    mine[k] = tuple(v) if k == "betas" else v
        st = {int(k): v for k, v in state_dict["state"].items()}
        for gi, (g, mine) in enumerate(zip(groups, self.param_groups)):
            per = {i: st[pid] for i, pid in enumerate(g["params"]) if pid in st}
            if not per:
                continue
            if gi in self._handles:
                self._restore(gi, per)
            else:
                self._pending[gi] = per          # applied when the first step creates the handle

    def _restore(self, gi, per):
        h, _, params = self._handles[gi]
        group = self.param_groups[gi]
        steps = set()
        for i, p in enumerate(group["params"]):
            if i not in per or not p.requires_grad:
                continue
            m, v, _ = self.moments(p)
            m.copy_(per[i]["exp_avg"].to(m))
            v.copy_(per[i]["exp_avg_sq"].to(v))
            steps.add(int(float(per[i]["step"])))
        if len(steps) > 1:
            raise ValueError("hicdiff_b200.optim.Adam keeps ONE step count per group; the checkpoint has several")
        if steps:
            _lib.check(_lib.load().hd_adam_set_step(h, steps.pop()), "hd_adam_set_step")

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
