"""Tile extraction / reassembly on the GPU (hd_tile_extract / hd_tile_scatter): bit-exact against the oracle's
splitPieces restatement; whole-chromosome pipeline (BASELINE config 4) round trip."""
import numpy as np
import pytest
import torch

import helpers
from oracle import hicdiff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 34, 64, 65, 130, 588, 703, 802])
@pytest.mark.parametrize("res", [40000, 10000])
def test_tile_extract_scatter_bit_exact(n, res):
    from hicdiff_b200 import ops

    rng = np.random.default_rng(n + res)
    a = rng.standard_normal((n, n)).astype(np.float32)
    a = (a + a.T) * 0.5
    band = O.band_blocks_for(res)
    ref = O.split_pieces(a, 64, res)
    got = ops.tile_extract(torch.from_numpy(a).cuda(), 64, band)
    assert got.shape == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)                              # bit-exact splitPieces
    back = ops.tile_scatter(got, n, 64, band)
    assert np.array_equal(back.cpu().numpy(), O.reassemble(ref, n, 64, res))  # bit-exact inverse
    assert torch.equal(ops.tile_extract(back, 64, band), got)                  # splitPieces(reassemble(t)) == t


def test_tile_edge_cases():
    from hicdiff_b200 import ops

    assert ops.tile_count(0) == 0
    empty = ops.tile_extract(torch.zeros(0, 0, device="cuda"))
    assert empty.shape == (0, 1, 64, 64)
    with pytest.raises(ValueError):
        ops.tile_scatter(torch.zeros(3, 1, 64, 64, device="cuda"), 64)        # 64x64 has exactly one tile


def test_whole_genome_pipeline_config4():
    """Six Drosophila-sized synthetic chromosomes -> 221 tiles -> HiCEDRN-conditional sampling (short chain, injected
    noise) -> reassembly.  The reassembled matrices equal the scatter of the oracle-ordered tiles bit-for-bit, and
    batching / tile offsets do not change any tile."""
    from hicdiff_b200 import genome, ops
    from hicdiff_b200.synthetic import synthetic_chromosome

    sizes = (588, 632, 703, 802, 34, 588)
    mats = [synthetic_chromosome(n, seed=i).cuda() for i, n in enumerate(sizes)]
    net, v = helpers.build_net("hicedrn_cond")
    T = 3
    diff = helpers.diffusion_cls("hicedrn_cond")(net.cuda(), image_size=64, timesteps=T, loss_type="l2",
                                                  beta_schedule="sigmoid").cuda()
    n_total = sum(ops.tile_count(n) for n in sizes)
    assert n_total == 221
    noise = torch.randn(T, n_total, 1, 64, 64, generator=torch.Generator().manual_seed(4)).cuda()
    out = genome.denoise_chromosomes(diff, mats, res=40000, max_batch=64, noise=noise)
    assert [o.shape[0] for o in out] == list(sizes)
    # reference ordering: the oracle's splitPieces on every chromosome, one batch, then reassemble on the CPU
    tiles = torch.cat([torch.from_numpy(O.split_pieces(m.cpu().numpy(), 64, 40000)) for m in mats]).cuda()
    ref_tiles = diff.super_resolution(tiles[:100], noise=noise[:, :100].contiguous())
    k = 0
    for m, o, n in zip(mats, out, sizes):
        c = ops.tile_count(n)
        if k + c <= 100:
            ref = O.reassemble(ref_tiles[k:k + c].cpu().numpy(), n, 64, 40000)
            assert np.array_equal(o.cpu().numpy(), ref)
        # off-diagonal blocks are mirrored by construction; a denoised DIAGONAL tile need not be symmetric itself
        blk = torch.arange(n, device=o.device) // 64
        off = (blk[:, None] != blk[None, :]).to(o.dtype)
        assert torch.equal(o * off, (o * off).t())
        k += c
    # Philox mode: world-size independent streams -> same result whatever the batching
    torch.manual_seed(5)
    a = genome.denoise_chromosomes(diff, mats[4:5], max_batch=64)
    torch.manual_seed(5)
    b = genome.denoise_chromosomes(diff, mats[4:5], max_batch=1)
    assert torch.equal(a[0], b[0])
