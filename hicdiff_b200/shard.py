"""Tile sharding across the GPUs of one box and the single exchange step of the path.

Tiles are independent units for all T steps (no cross-tile op anywhere in the eps-net or the posterior), so the
data path has NO collective: every rank denoises a contiguous slice of the global tile list with Philox streams keyed
by the GLOBAL tile id (results do not depend on the world size).  The only exchange is one all-gather of the finished
tiles before reassembly (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_tiles: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of rank `rank`; sizes differ by at most one, ranks beyond n_tiles get nothing."""
    if n_tiles < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad shard request: n_tiles={n_tiles}, rank={rank}, world={world}")
    base, rem = divmod(n_tiles, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class nvtx_range:
    """NVTX range on CUDA tensors' phases (tile gather / reassembly); a no-op on CPU-only runs (the gloo tests)."""

    def __init__(self, name: str, enabled: bool):
        self.name, self.enabled = name, enabled and torch.cuda.is_available()

    def __enter__(self):
        if self.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def gather_tiles(local: torch.Tensor, n_tiles: int, group=None) -> torch.Tensor:
    with nvtx_range("hicdiff_b200: gather tiles (all_gather)", local.is_cuda):
        return _gather_tiles(local, n_tiles, group)


def _gather_tiles(local: torch.Tensor, n_tiles: int, group=None) -> torch.Tensor:
    """All-gather the per-rank slices produced under `shard_range` back into the global [n_tiles, 1, H, W] order.
    Ragged slices are padded to the largest slice for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[0] != n_tiles:
            raise ValueError(f"single-rank gather expects all {n_tiles} tiles, got {local.shape[0]}")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, stop = shard_range(n_tiles, rank, world)
    if local.shape[0] != stop - start:
        raise ValueError(f"rank {rank} holds {local.shape[0]} tiles, its shard is [{start}, {stop})")
    per = -(-n_tiles // world) if n_tiles else 0
    if per == 0:
        return local[:0]
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        s, e = shard_range(n_tiles, r, world)
        parts.append(out[r * per: r * per + (e - s)])
    return torch.cat(parts, dim=0)
