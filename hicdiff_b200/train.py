"""Training step on the sm_100a trainer (`hd_trainer_*`, include/hicdiff_b200.h).

The reference trains hicedrn_Diff through `loss = diffusion(x); loss.backward(); optimizer.step()` (train.py:109-136).
Here `GaussianDiffusion.p_losses` hands the whole forward + loss + backward of the eps-net to libhicdiff_b200.so in ONE call
and returns a scalar whose autograd node delivers the parameter gradients, so the loop above -- including a stock
`torch.optim.Adam(diffusion.parameters())` -- runs unchanged.  Parameters are bound by pointer: the optimizer's in-place
updates are seen by the next step without any re-upload.  There is no ATen fallback.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Optional

import torch

from . import _lib


_TRAINABLE = (_lib.HD_HICEDRN, _lib.HD_HICEDRN_SR3, _lib.HD_UNET, _lib.HD_UNET_SR3)

# Gradient buckets of the Unet's backward (data-parallel training).  The backward walks the modules in reverse:
# bucket k is final when it reaches the first module starting with BUCKET_BOUNDARIES[k] (include/hicdiff_b200.h,
# hd_trainer_set_grad_buckets); the last bucket (the rest, plus everything that depends on the time embedding and init_conv,
# whose gradients accumulate until the end of the step) is reduced after the step.
BUCKET_BOUNDARIES = ("ups.2.", "ups.0.", "downs.2.")
_BUCKET_MODULES = (("final_conv.", "final_res_block.", "ups.3."), ("ups.2.", "ups.1."), ("ups.0.", "mid_", "downs.3."))


def grad_bucket_of(name: str) -> int:
    """Bucket of parameter `name` (state_dict key without `model.`): 0 .. len(BUCKET_BOUNDARIES); the last one is 'late'."""
    late = len(_BUCKET_MODULES)
    if ".mlp." in name or "noise_func" in name or name.startswith(("time_mlp.", "init_conv.")):
        return late        # FiLM / noise-level Linears and the time MLP: back-propagated in one batch at the end of the step
    for k, prefixes in enumerate(_BUCKET_MODULES):
        if name.startswith(prefixes):
            return k
    return late


def grad_bucket_layout(names, numels):
    """(order, ranges): parameter indices in flat-buffer order (bucket by bucket, module order inside a bucket) and the
    [lo, hi) element range of every bucket.  Pure host logic (tests/test_host_cpu.py)."""
    nb = len(_BUCKET_MODULES) + 1
    order = sorted(range(len(names)), key=lambda i: (grad_bucket_of(names[i]), i))
    ranges, off = [], 0
    for b in range(nb):
        lo = off
        off += sum(numels[i] for i in order if grad_bucket_of(names[i]) == b)
        ranges.append((lo, off))
    return order, ranges


class Trainer:
    """One `hd_trainer` for a parameter-holder net (nets.hicedrn_Diff) at a fixed batch size."""

    def __init__(self, net: torch.nn.Module, batch: int):
        lib = _lib.load()
        cfgd = net._plan_config()
        if cfgd["variant"] not in _TRAINABLE:
            raise NotImplementedError("hicdiff_b200: no training backward for this eps-net variant")
        params = dict(net.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("hicdiff_b200 trains on sm_100a GPUs only: move the module to a CUDA device (no CPU fallback)")
        for k, p in params.items():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"parameter {k} must be contiguous fp32 (got {p.dtype})")
        self.net = net
        self.batch = int(batch)
        self.device = dev
        self._ptrs = {k: p.data_ptr() for k, p in params.items()}
        total = sum(p.numel() for p in params.values())
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grads: Dict[str, torch.Tensor] = {}
        names = list(params)
        # Unet: the flat buffer is laid out bucket by bucket in the order the backward completes them, so that each bucket is
        # ONE contiguous all-reduce that can start while the backward is still running (enable_gradient_allreduce)
        self.bucket_ranges = None
        if cfgd["variant"] in (_lib.HD_UNET, _lib.HD_UNET_SR3):
            order, self.bucket_ranges = grad_bucket_layout(names, [params[k].numel() for k in names])
        else:
            order = range(len(names))
        off = 0
        views = {}
        for i in order:
            k = names[i]
            views[k] = self.flat_grad[off:off + params[k].numel()].view_as(params[k])
            off += params[k].numel()
        for k in names:                       # `grads` keeps named_parameters order (the autograd node returns them in it)
            self.grads[k] = views[k]
        self._comm_stream = None
        cfg = _lib.hd_config()
        cfg.abi_version = _lib.HD_ABI_VERSION
        cfg.variant = cfgd["variant"]
        cfg.self_condition = cfgd["self_condition"]
        cfg.image_size = 64
        cfg.timesteps = 1
        cfg.num_blocks = cfgd["num_blocks"]
        cfg.dim = cfgd["dim"]
        cfg.num_mults = len(cfgd["dim_mults"])
        for i, m in enumerate(cfgd["dim_mults"]):
            cfg.dim_mults[i] = int(m)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.hd_trainer_create(C.byref(cfg), self.batch, C.byref(h)), "hd_trainer_create")
            self._handle: Optional[int] = h.value
            if self.bucket_ranges is not None:
                arr = (C.c_char_p * len(BUCKET_BOUNDARIES))(*[b.encode() for b in BUCKET_BOUNDARIES])
                _lib.check(lib.hd_trainer_set_grad_buckets(self._handle, arr, len(BUCKET_BOUNDARIES)), "hd_trainer_set_grad_buckets")
            for k, p in params.items():
                shape = (C.c_int64 * max(p.dim(), 1))(*p.shape)
                _lib.check(lib.hd_trainer_bind(self._handle, k.encode(), p.data_ptr(), self.grads[k].data_ptr(), shape, p.dim()),
                           f"hd_trainer_bind({k})")
            _lib.check(lib.hd_trainer_finalize(self._handle, _lib.stream_ptr()), "hd_trainer_finalize")

    def matches(self, net, batch: int) -> bool:
        if self._handle is None or batch != self.batch:
            return False
        ps = dict(net.named_parameters())
        return ps.keys() == self._ptrs.keys() and all(p.data_ptr() == self._ptrs[k] for k, p in ps.items())

    def destroy(self):
        if self._handle is not None:
            _lib.load().hd_trainer_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass

    def _tiles(self, t, name):
        if t.shape != (self.batch, 1, 64, 64):
            raise ValueError(f"{name} must be [{self.batch}, 1, 64, 64] (got {tuple(t.shape)})")
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device}, the module is on {self.device}")
        return t.detach().to(torch.float32).contiguous()

    def step(self, x_t, cond, time, target, weight, loss_type: str, want_eps: bool = False):
        """Forward + loss + backward.  Returns (loss [] fp32 device tensor, eps or None); gradients land in `self.grads`."""
        lib = _lib.load()
        x_t = self._tiles(x_t, "x_t")
        target = self._tiles(target, "target")
        cond = self._tiles(cond, "cond") if cond is not None else None
        tv = time.detach().reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()
        wv = weight.detach().reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()
        if tv.numel() != self.batch or wv.numel() != self.batch:
            raise ValueError("time and weight must have one entry per sample")
        kind = {"l1": 0, "l2": 1}[loss_type]
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        eps = torch.empty_like(x_t) if want_eps else None
        with torch.cuda.device(self.device):
            _lib.check(lib.hd_trainer_step(self._handle, x_t.data_ptr(), _lib.ptr(cond), tv.data_ptr(), target.data_ptr(),
                                           wv.data_ptr(), kind, _lib.ptr(eps), loss.data_ptr(), _lib.stream_ptr()),
                       "hd_trainer_step")
        return loss, eps

    def profile(self, reps: int = 3) -> dict:
        buf = C.create_string_buffer(1 << 16)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hd_trainer_profile(self._handle, reps, buf, len(buf), _lib.stream_ptr()), "hd_trainer_profile")
        return json.loads(buf.value.decode())

    def device_bytes(self) -> int:
        return int(_lib.load().hd_trainer_device_bytes(self._handle))

    def num_launch_groups(self) -> int:
        return int(_lib.load().hd_trainer_num_launches(self._handle))


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """DDP's gradient exchange for BASELINE config 5: ONE all-reduce over the trainer's flat fp32 gradient buffer (NCCL over
    NVLink on GPUs; any backend works), averaged over the ranks.  No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    return flat


def enable_gradient_allreduce(net, group=None, enabled: bool = True, overlap: bool = False) -> None:
    """Data-parallel replicas (one process per GPU, torch.distributed initialised by the caller): average the gradients
    across ranks inside every training step, as DistributedDataParallel would.  overlap=False (default; and always for the
    hicedrn trainers): one all-reduce of the whole flat buffer after the step.  overlap=True (Unet trainers): the buffer is
    reduced in buckets on a communication stream, each bucket as soon as the backward has finished it (an event the step's CUDA
    graph records), so NCCL runs next to the remaining backward kernels -- DistributedDataParallel's scheme.  Measured on
    8 x B200 (NVSwitch, 143 MB of fp32 gradients, 12.4 ms step): flat 12.45 ms, overlapped 12.65 ms (N = 2: 12.41 / 12.51) --
    the single all-reduce costs < 0.1 ms here, and four smaller ones that share the SMs with the backward cost more than they hide,
    hence the default; the overlapped form is for slower fabrics."""
    object.__setattr__(net, "_grad_allreduce", (group, bool(overlap)) if enabled else None)


def allreduce_buckets_(trainer, group=None) -> None:
    """Bucketed, overlapped form of allreduce_mean_ for a trainer whose step has just been enqueued on the current stream."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    flat, ranges = trainer.flat_grad, trainer.bucket_ranges
    main = torch.cuda.current_stream(trainer.device)
    if trainer._comm_stream is None:
        trainer._comm_stream = torch.cuda.Stream(device=trainer.device)
    comm = trainer._comm_stream
    lib = _lib.load()
    with torch.cuda.device(trainer.device):
        for k, (lo, hi) in enumerate(ranges[:-1]):
            if hi == lo:
                continue
            # the communication stream waits for bucket k's event inside the step that was just launched -- not for the step
            _lib.check(lib.hd_trainer_wait_grad_bucket(trainer._handle, k, comm.cuda_stream), "hd_trainer_wait_grad_bucket")
            with torch.cuda.stream(comm):
                dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
                flat[lo:hi].div_(world)
        lo, hi = ranges[-1]
        if hi > lo:                       # the late bucket: final only when the step is (main stream order)
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
            flat[lo:hi].div_(world)
        main.wait_stream(comm)


class _TrainStep(torch.autograd.Function):
    """loss = trainer.step(...) as an autograd node over the net's parameters (the backward already ran on the device)."""

    @staticmethod
    def forward(ctx, trainer, x_t, cond, time, target, weight, loss_type, *params):
        loss, _ = trainer.step(x_t, cond, time, target, weight, loss_type)
        ar = getattr(trainer.net, "_grad_allreduce", None)
        if ar is not None:
            if len(ar) > 1 and ar[1] and trainer.bucket_ranges is not None:
                allreduce_buckets_(trainer, ar[0])
            else:
                allreduce_mean_(trainer.flat_grad, ar[0])
        # snapshot: the trainer's buffer is rewritten by the next step, and a caller may run several forwards before any
        # backward (e.g. (loss1 + loss2).backward()) -- every loss must keep ITS gradients (one flat copy, ~0.05 ms)
        ctx.flat = trainer.flat_grad.clone()
        ctx.shapes = [(g.storage_offset(), g.shape, g.numel()) for g in trainer.grads.values()]   # named_parameters order
        return loss

    @staticmethod
    def backward(ctx, gout):
        flat = gout * ctx.flat                # ONE launch; a fresh tensor
        grads = [flat[off:off + numel].view(shape) for off, shape, numel in ctx.shapes]
        return (None,) * 7 + tuple(grads)


def supports_training(net) -> bool:
    cfg = getattr(net, "_plan_config", None)
    return cfg is not None and cfg()["variant"] in _TRAINABLE


def training_loss(net, x_t, time, cond, target, weight, loss_type: str):
    """Scalar loss with a grad_fn over `net.parameters()` (named_parameters order).  One trainer is kept per batch size (an
    epoch's ragged last batch must not evict the main one); all are rebuilt if the parameters move (`.to()`, new storage)."""
    batch = x_t.shape[0]
    cache = getattr(net, "_trainers", None)
    if cache is None:
        cache = {}
        object.__setattr__(net, "_trainers", cache)
    tr = cache.get(batch)
    if tr is None or not tr.matches(net, batch):
        if any(not t.matches(net, t.batch) for t in cache.values()):
            for t in cache.values():
                t.destroy()
            cache.clear()
        while len(cache) >= 2:                       # at most two batch sizes resident (activations are GBs per trainer)
            cache.pop(next(iter(cache))).destroy()
        tr = Trainer(net, batch)
        cache[batch] = tr
    object.__setattr__(net, "_trainer", tr)      # the most recently used one (benchmarks / introspection)
    params = [p for _, p in net.named_parameters()]
    return _TrainStep.apply(tr, x_t, cond, time, target, weight, loss_type, *params)
