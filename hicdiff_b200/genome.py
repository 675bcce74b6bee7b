"""Whole-chromosome driver around the sampling path (BASELINE.json config 4): tile -> shard -> denoise -> gather ->
reassemble.  The reference stops at flat tile arrays (`predict.npy` + chromosome ids,
/root/reference/src/Utils/metrics_cond.py:126-134) and has no reassembly; the inverse used here is the exact inverse
of splitPieces' enumeration (/root/reference/processdata/PrepareData_linear.py:25-46), see ops.tile_scatter.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .shard import gather_tiles, nvtx_range, shard_range


def band_blocks_for(res: int, piece: int = 64) -> int:
    """Largest kept block distance of splitPieces' `abs(i - j) <= int(piece_size * 4 * scal + 1)` with step == piece."""
    return int(piece * 4 * int(40000 / res) + 1) // piece


@torch.no_grad()
def denoise_chromosomes(diffusion, mats: Sequence[torch.Tensor], res: int = 40000, max_batch: int = 512,
                        noise: Optional[torch.Tensor] = None, group=None) -> List[torch.Tensor]:
    """Denoise a list of symmetric n_i x n_i contact maps (CUDA fp32, values in [-1, 1]) with a conditional diffusion.

    Every rank must pass the same `mats`.  Tiles of all chromosomes form one global list (chromosome-major, splitPieces
    order inside a chromosome); rank r denoises `shard_range(n_total, r, world)` in batches of <= max_batch with
    Philox streams keyed by the global tile id, then the slices are all-gathered and scattered back.
    `noise` ([T, n_total, 1, 64, 64], optional) injects the draws for parity runs.
    """
    import torch.distributed as dist

    band = band_blocks_for(res)
    tiles = [ops.tile_extract(m, 64, band) for m in mats]
    counts = [t.shape[0] for t in tiles]
    all_tiles = torch.cat(tiles, dim=0) if tiles else None
    n_total = sum(counts)
    if n_total == 0:
        return [torch.zeros_like(m) for m in mats]
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    start, stop = shard_range(n_total, rank, world)
    plan = diffusion._sync_plan()
    seed = diffusion._next_seed(None)
    if world > 1:   # one seed for the whole job
        s = torch.tensor([seed], device=all_tiles.device, dtype=torch.int64)
        dist.broadcast(s, src=0, group=group)
        seed = int(s.item())
    outs = []
    for b0 in range(start, stop, max_batch):
        b1 = min(stop, b0 + max_batch)
        nz = noise[:, b0:b1].contiguous() if noise is not None else None
        outs.append(plan.sample(b1 - b0, cond=all_tiles[b0:b1], noise=nz, seed=seed, tile_offset=b0))
    local = torch.cat(outs, dim=0) if outs else all_tiles[:0]
    full = gather_tiles(local, n_total, group)
    res_mats, k = [], 0
    with nvtx_range("hicdiff_b200: reassemble chromosomes (tile_scatter)", full.is_cuda):
        for m, c in zip(mats, counts):
            res_mats.append(ops.tile_scatter(full[k:k + c], m.shape[0], 64, band))
            k += c
    return res_mats
