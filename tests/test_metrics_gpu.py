"""GPU-side SSIM / PSNR / metrics loop (hd_ssim_mse_tiles) against the oracle's restatement of
src/Utils/loss/SSIM.py and the reference's PSNR; tolerance 1e-5 absolute on SSIM (fp32 summation order only) -- two
orders below the 1e-3 bar north_star sets for the metrics."""
import numpy as np
import pytest
import torch

from oracle import hicdiff_oracle as O

pytestmark = pytest.mark.gpu


def _pair(b, seed):
    from hicdiff_b200.synthetic import synthetic_tiles

    clean, noisy = synthetic_tiles(b, seed=seed)
    return clean, noisy


@pytest.mark.parametrize("b", [1, 3, 64])
@pytest.mark.parametrize("rescale", [False, True])
def test_ssim_mse_per_tile_matches_oracle(b, rescale):
    from hicdiff_b200 import metrics

    clean, noisy = _pair(b, 7 + b)
    s, m = metrics.ssim_mse_per_tile(noisy.cuda(), clean.cuda(), rescale=rescale)
    ra, rb = (O.to_unit_range(noisy), O.to_unit_range(clean)) if rescale else (noisy, clean)
    for i in range(b):
        ref_s = float(O.ssim(ra[i:i + 1], rb[i:i + 1]))
        ref_m = float(((ra[i:i + 1] - rb[i:i + 1]) ** 2).mean())
        assert abs(float(s[i]) - ref_s) <= 1e-5, (i, float(s[i]), ref_s)
        assert abs(float(m[i]) - ref_m) <= 1e-6 * max(1.0, ref_m)
    # the reference's scalar SSIM / PSNR are the means over equal-size tiles
    assert abs(float(s.mean()) - float(O.ssim(ra, rb))) <= 1e-5
    assert abs(float(10 * torch.log10(1 / m.mean())) - float(O.psnr(ra, rb))) <= 1e-4


def test_reference_signatures_and_edge_cases():
    from hicdiff_b200 import metrics

    clean, noisy = _pair(4, 3)
    a, b = O.to_unit_range(noisy), O.to_unit_range(clean)
    assert abs(float(metrics.ssim(a.cuda(), b.cuda())) - float(O.ssim(a, b))) <= 1e-5
    assert metrics.ssim(a.cuda(), b.cuda(), size_average=False).shape == (4,)
    assert abs(float(metrics.SSIM()(a.cuda(), b.cuda())) - float(O.ssim(a, b))) <= 1e-5
    assert abs(float(metrics.psnr(a.cuda(), b.cuda())) - float(O.psnr(a, b))) <= 1e-4
    assert abs(float(metrics.ssim(a.cuda(), a.cuda())) - 1.0) <= 1e-6          # identical images
    assert torch.equal(metrics.inverse_data_transform("rescaled", noisy), O.to_unit_range(noisy))
    s, m = metrics.ssim_mse_per_tile(torch.zeros(0, 1, 64, 64, device="cuda"), torch.zeros(0, 1, 64, 64, device="cuda"))
    assert s.numel() == 0 and m.numel() == 0
    with pytest.raises(RuntimeError):
        metrics.ssim(a, b)                                                       # CPU tensors: no fallback
    with pytest.raises(ValueError):
        metrics.ssim(a.cuda()[:, :, :32], b.cuda()[:, :, :32])


def test_get_metrics_loop_and_wire_format(tmp_path):
    """metrics_cond.getMetrics' loop with an identity 'model': files in the Outputs_diff wire format, SSIM / PSNR of the run."""
    from hicdiff_b200 import metrics

    clean, noisy = _pair(10, 11)
    inds = torch.arange(40).reshape(10, 4)
    loader = [(noisy[i:i + 4], clean[i:i + 4], None, inds[i:i + 4]) for i in range(0, 10, 4)]
    r = metrics.get_metrics(lambda lr: lr, loader, out_dir=tmp_path / "Outputs_diff" / "run")
    assert r["nsamples"] == 10
    ra, rb = O.to_unit_range(noisy), O.to_unit_range(clean)
    assert abs(r["ssim"] - float(O.ssim(ra, rb))) <= 1e-5
    assert abs(r["psnr"] - float(O.psnr(ra, rb))) <= 1e-4
    d = tmp_path / "Outputs_diff" / "run"
    assert np.array_equal(np.load(d / "target.npy"), clean.numpy())
    assert np.array_equal(np.load(d / "noisy.npy"), noisy.numpy())
    assert np.array_equal(np.load(d / "predict.npy"), noisy.numpy())
    assert np.array_equal(np.load(d / "inds.npy"), inds.numpy())
