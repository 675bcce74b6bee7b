"""Data preparation on the GPU: the reference's `loadBothConstraints`
(processdata/PrepareData_linear.py:48-103) and the tile / noisy-tile generation of `split_numpy` (:183-213), as device
kernels behind `hd_coo_to_dense / hd_remove_empty_bins / hd_select_ranks / hd_normalize_contacts / hd_tile_extract /
hd_add_noise`.  Results are bit-identical to the reference's numpy path (tests/test_prepare_gpu.py); the only host work is
parsing the text dump and numpy's own interpolation between the two order statistics the GPU selects."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib, ops

ArrayOrPath = Union[str, os.PathLike, np.ndarray]


def _triples(x: ArrayOrPath) -> np.ndarray:
    a = np.loadtxt(x) if isinstance(x, (str, os.PathLike)) else np.asarray(x)   # :49-50
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("expected (pos1, pos2, value) rows")
    return a


def _cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hicdiff_b200 data preparation runs on sm_100a GPUs only (there is no CPU fallback)")
    return dev


def dense_from_triples(rows, cols, vals, smallbin: int, n: int, device="cuda") -> torch.Tensor:
    """The fill loop :67-72: symmetric fp32 [n, n]; later triples overwrite earlier ones."""
    dev = _cuda(device)
    r = torch.as_tensor(np.ascontiguousarray(rows, dtype=np.int64)).to(dev)
    c = torch.as_tensor(np.ascontiguousarray(cols, dtype=np.int64)).to(dev)
    v = torch.as_tensor(np.ascontiguousarray(vals, dtype=np.float32)).to(dev)      # `mata[...] = ia` rounds to float32
    mat = torch.empty(n, n, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().hd_coo_to_dense(r.data_ptr(), c.data_ptr(), v.data_ptr(), r.numel(), int(smallbin), int(n),
                                               mat.data_ptr(), _lib.stream_ptr()), "hd_coo_to_dense")
    return mat


def remove_empty_bins(mat: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """:77-85: drop rows / columns whose diagonal is 0 or NaN.  Returns (matrix [m, m], kept bin indices [m])."""
    _cuda(mat.device)
    n = mat.shape[0]
    mat = mat.to(torch.float32).contiguous()
    out = torch.empty(n * n, device=mat.device, dtype=torch.float32)
    idx = torch.empty(max(n, 1), device=mat.device, dtype=torch.int64)
    m = C.c_int64(0)
    with torch.cuda.device(mat.device):
        _lib.check(_lib.load().hd_remove_empty_bins(mat.data_ptr(), n, out.data_ptr(), idx.data_ptr(), C.byref(m), _lib.stream_ptr()),
                   "hd_remove_empty_bins")
    m = int(m.value)
    return out[:m * m].view(m, m), idx[:m]


_NUMPY_MAJOR = int(np.__version__.split(".")[0])


def percentile(x: torch.Tensor, q: float) -> np.floating:
    """np.percentile(x, q) ('linear' method, the default the reference uses at :88) of a CUDA fp32 tensor.  The two
    neighbouring order statistics are selected exactly on the device; the virtual index, the interpolation weight and the
    lerp are restated here for numpy's two dtype rules -- numpy >= 2 does this arithmetic in the array's float32, numpy
    1.x (the reference's pinned 1.23) in float64 -- and checked against np.percentile of the installed numpy in
    tests/test_host_cpu.py and tests/test_prepare_gpu.py (no private numpy helpers involved)."""
    _cuda(x.device)
    x = x.to(torch.float32).contiguous()
    n = x.numel()
    if n == 0:
        raise ValueError("percentile of an empty tensor")
    lo, hi, gamma = percentile_plan(n, q)
    ranks = (C.c_int64 * 2)(lo, hi)
    outv = (C.c_float * 2)()
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hd_select_ranks(x.data_ptr(), n, ranks, 2, outv, _lib.stream_ptr()), "hd_select_ranks")
    return percentile_lerp(np.float32(outv[0]), np.float32(outv[1]), gamma)


def percentile_plan(n: int, q: float):
    """(lower rank, upper rank, interpolation weight) of np.percentile(a, q) for len(a) == n, 'linear' method:
    virtual index (n - 1) * q / 100, floor / floor + 1 as the neighbours (clamped to the last index), gamma = the
    fractional part, all in the float type numpy itself uses for a float32 array."""
    ft = np.float32 if _NUMPY_MAJOR >= 2 else np.float64
    quant = np.true_divide(ft(q), ft(100))
    vi = ft(n - 1) * quant                                   # _compute_virtual_index for alpha = beta = 1
    prev = np.floor(vi)
    lo = int(prev)
    hi = lo + 1
    if vi >= n - 1:                                          # at / above the last index: both neighbours are the maximum
        lo = hi = n - 1
    if vi < 0:
        lo = hi = 0
    gamma = ft(vi - prev)
    return lo, min(hi, n - 1), gamma


def percentile_lerp(a, b, t):
    """numpy's `_lerp`: a + (b - a) * t, switched to b - (b - a) * (1 - t) for t >= 0.5 (monotone at both ends);
    evaluated in the dtype numpy promotes to (float32 data with a float32 / float64 weight)."""
    diff = np.subtract(b, a)
    if t >= 0.5:
        return np.subtract(b, diff * (1 - t))
    return np.add(a, diff * t)


def normalize_contacts_(mat: torch.Tensor, per) -> torch.Tensor:
    """:90-92 in place: 2 * (clip(mat, 0, per) / per) - 1."""
    _cuda(mat.device)
    if mat.dtype != torch.float32 or not mat.is_contiguous():
        raise ValueError("expected a contiguous fp32 tensor")
    with torch.cuda.device(mat.device):
        _lib.check(_lib.load().hd_normalize_contacts(mat.data_ptr(), mat.numel(), float(np.float32(per)), _lib.stream_ptr()),
                   "hd_normalize_contacts")
    return mat


def load_both_constraints(stria: ArrayOrPath, strib: ArrayOrPath, res: int, device="cuda") -> torch.Tensor:
    """`loadBothConstraints(stria, strib, res)` (:48-103): returns the normalised matrix in [-1, 1] on the device.  Like the
    reference, the raw-count dump `strib` only contributes to the bin range."""
    a, b = _triples(stria), _triples(strib)
    rowsa, colsa = (a[:, 0] / res).astype(int), (a[:, 1] / res).astype(int)
    rowsb, colsb = (b[:, 0] / res).astype(int), (b[:, 1] / res).astype(int)
    bigbin = int(np.max((np.max((rowsa, colsa)), np.max((rowsb, colsb)))))
    smallbin = int(np.min((np.min((rowsa, colsa)), np.min((rowsb, colsb)))))
    mat = dense_from_triples(rowsa, colsa, a[:, 2], smallbin, bigbin - smallbin + 1, device)
    mat, _ = remove_empty_bins(mat)
    mat = mat.contiguous()
    per = percentile(mat, 99.0)
    return normalize_contacts_(mat, per)


def _fresh_seed() -> int:
    """A Philox key drawn from torch's global CPU generator: every call advances it, so consecutive chromosomes get
    independent noise (the reference draws a fresh `torch.randn_like` per chromosome, :203-204) and `torch.manual_seed`
    still makes a whole run reproducible."""
    return int(torch.randint(0, 2 ** 62, (1,), device="cpu").item())


def add_noise(tiles: torch.Tensor, sigma_0: float, noise: Optional[torch.Tensor] = None, seed: Optional[int] = None) -> torch.Tensor:
    """`data + sigma_0 * torch.randn_like(data)` (:203-204, 'deno').  `noise` injects the draws (parity); otherwise the
    in-kernel Philox generator of the sampling path supplies them (one stream per tile, keyed by `seed`; `seed=None`
    draws a fresh key from torch's generator per call, so no two calls share a noise field)."""
    _cuda(tiles.device)
    x = tiles.detach().to(torch.float32).contiguous()
    if noise is None:
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 64, 64):
            raise ValueError("Philox noise is generated per 64x64 tile: pass [B,1,64,64] or explicit `noise`")
        noise = ops.philox_normal(x.shape[0], _fresh_seed() if seed is None else int(seed), 0, device=x.device)
    z = noise.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if z.shape != x.shape:
        raise ValueError("noise must have the shape of the tiles")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hd_add_noise(x.data_ptr(), z.data_ptr(), float(sigma_0), x.numel(), out.data_ptr(), _lib.stream_ptr()),
                   "hd_add_noise")
    return out


def make_splits(mat: torch.Tensor, res: int = 40000, piece_size: int = 64, sigma_0: float = 0.1, noise=None,
                seed: Optional[int] = None):
    """split_numpy (:183-213) for one chromosome and the 'deno' degradation: (target tiles, noisy tiles), both
    [n_tiles, 1, 64, 64] on the device -- what the reference stores as `*_full_chr_*` / `*_noisy_chr_*`."""
    band = 4 * int(40000 / res)
    target = ops.tile_extract(mat.contiguous(), piece_size, band)
    return target, add_noise(target, sigma_0, noise, seed)
