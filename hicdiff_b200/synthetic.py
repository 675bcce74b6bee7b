"""Seeded synthetic Hi-C-like inputs; shared by the tests, bench.py and the golden generator.
Pure data generators on the CPU -- nothing here computes any part of the sampling path."""
import torch


def synthetic_tiles(batch: int, seed: int = 1234, sigma: float = 0.1):
    """(clean, noisy): clean symmetric band-decay tiles in [-1, 1]; noisy = clean + sigma * randn, mirroring how the
    reference builds its conditional inputs (/root/reference/processdata/PrepareData_linear.py:203-204)."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.arange(64)
    d = (idx[:, None] - idx[None, :]).abs().float()
    u = torch.rand(batch, 1, 64, 64, generator=g)
    u = 0.5 * (u + u.transpose(-1, -2))
    clean = (2 * torch.exp(-d / 8) * (0.6 + 0.4 * u) - 1).clamp(-1, 1)
    noisy = clean + sigma * torch.randn(batch, 1, 64, 64, generator=g)
    return clean, noisy


def synthetic_noise(timesteps: int, batch: int, seed: int = 2024):
    """The T draws of one reverse chain: [0] = x_T, [i] = z of step t = T - i."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(timesteps, batch, 1, 64, 64, generator=g)


def synthetic_chromosome(n: int, seed: int = 0):
    """Symmetric n x n contact map in [-1, 1] with distance decay (stand-in for one normalised chromosome)."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.arange(n)
    d = (idx[:, None] - idx[None, :]).abs().float()
    u = torch.rand(n, n, generator=g)
    u = 0.5 * (u + u.t())
    return (2 * torch.exp(-d / 12) * (0.5 + 0.5 * u) - 1).clamp(-1, 1)
