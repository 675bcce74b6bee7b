"""Measurement side of the sampling path on the GPU.

`ssim` / `psnr` / `inverse_data_transform` have the reference's signatures (src/Utils/loss/SSIM.py:66-74,
src/datasets/__init__.py:214-223, pretrain/train_unet_Diff_cond_n.py:125-133) and run as one CUDA kernel per call
(`hd_ssim_mse_tiles`); `get_metrics` is the loop of src/Utils/metrics_cond.py:95-137 without its per-batch host round
trips: tiles stay on the device, the four result arrays are written once in the reference's `Outputs_diff/<name>/`
wire format (`target.npy`, `noisy.npy`, `predict.npy`, `inds.npy`)."""
from __future__ import annotations

import math
import os
from pathlib import Path
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch

from . import _lib


def gaussian(window_size: int, sigma: float) -> torch.Tensor:  # SSIM.py:6-8
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def create_window(window_size: int = 11, channel: int = 1) -> torch.Tensor:  # SSIM.py:10-14
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def inverse_data_transform(config, X):  # datasets/__init__.py:214-223 ('rescaled' / 'logit_transform' / identity)
    if config == "logit_transform":
        X = torch.sigmoid(X)
    elif config == "rescaled":
        X = (X + 1.0) / 2.0
    return torch.clamp(X, 0.0, 1.0)


def _tiles(t: torch.Tensor) -> torch.Tensor:
    if t.device.type != "cuda":
        raise RuntimeError("hicdiff_b200 metrics run on sm_100a GPUs only (there is no CPU fallback)")
    if t.dim() != 4 or t.shape[1] != 1 or t.shape[2:] != (64, 64):
        raise ValueError(f"expected [B,1,64,64] tiles, got {tuple(t.shape)}")
    return t.detach().to(torch.float32).contiguous()


def ssim_mse_per_tile(img1, img2, rescale: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-tile (SSIM, MSE), fp32 [B] each.  `rescale` applies inverse_data_transform('rescaled') to both inputs first."""
    a, b = _tiles(img1), _tiles(img2)
    if a.shape != b.shape:
        raise ValueError("img1 and img2 must have the same shape")
    B = a.shape[0]
    win = create_window(11, 1).reshape(-1).to(a.device)
    s = torch.empty(B, device=a.device, dtype=torch.float32)
    m = torch.empty(B, device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().hd_ssim_mse_tiles(a.data_ptr(), b.data_ptr(), win.data_ptr(), s.data_ptr(), m.data_ptr(), B,
                                             1 if rescale else 0, _lib.stream_ptr()), "hd_ssim_mse_tiles")
    return s, m


def ssim(img1, img2, window_size: int = 11, size_average: bool = True):  # SSIM.py:66-74
    if window_size != 11:
        raise NotImplementedError("the reference only uses the 11x11 window")
    s, _ = ssim_mse_per_tile(img1, img2)
    return s.mean() if size_average else s


class SSIM(torch.nn.Module):  # SSIM.py:39-63 (same constructor; the window is rebuilt per call like `ssim`)
    def __init__(self, window_size: int = 11, size_average: bool = True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1
        self.window = create_window(window_size, self.channel)

    def forward(self, img1, img2):
        return ssim(img1, img2, self.window_size, self.size_average)


def psnr(img1, img2):
    """10 log10(1 / mse) on [0, 1] images (the validation blocks of the pretrain scripts)."""
    _, m = ssim_mse_per_tile(img1, img2)
    return 10 * torch.log10(1 / m.mean())


@torch.no_grad()
def get_metrics(model, loader: Iterable, out_dir: Optional[os.PathLike] = None, device="cuda") -> Dict[str, object]:
    """metrics_cond.getMetrics' loop: `model` is the bound `diffusion.super_resolution` (inference.py:98); `loader` yields
    (noisy, target, _, inds) batches.  Returns the predictions (numpy, like the reference) plus SSIM / PSNR of the run."""
    preds, hrs, lrs, inds_all, ss, ms = [], [], [], [], [], []
    for lr, hr, _, inds in loader:
        lr = lr.to(device, torch.float32)
        hr = hr.to(device, torch.float32)
        out = model(lr)
        s, m = ssim_mse_per_tile(out, hr, rescale=True)
        preds.append(out); hrs.append(hr); lrs.append(lr); inds_all.append(torch.as_tensor(inds)); ss.append(s); ms.append(m)
    if not preds:
        raise ValueError("empty loader")
    predict = torch.cat(preds).cpu().numpy()
    target = torch.cat(hrs).cpu().numpy()
    low = torch.cat(lrs).cpu().numpy()
    index = torch.cat(inds_all).cpu().numpy()
    s_all, m_all = torch.cat(ss), torch.cat(ms)
    if out_dir is not None:
        d = Path(out_dir)
        d.mkdir(parents=True, exist_ok=True)
        np.save(d / "target", target)
        np.save(d / "noisy", low)
        np.save(d / "predict", predict)
        np.save(d / "inds", index)
    return {"predict": predict, "ssim": float(s_all.mean()), "psnr": float(10 * torch.log10(1 / m_all.mean())),
            "mse": float(m_all.mean()), "nsamples": int(s_all.numel())}
