"""Data preparation on the GPU (hd_coo_to_dense / hd_remove_empty_bins / hd_select_ranks /
hd_normalize_contacts / hd_add_noise) against the oracle's restatement of loadBothConstraints / split_numpy, which
oracle/make_golden_prepare.py pins bit-for-bit to the unmodified reference.  Everything here is indexing, order statistics or
IEEE fp32 arithmetic in the reference's order: the bar is BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import hicdiff_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n_bins,res,seed", [(150, 40000, 3), (700, 40000, 4), (333, 10000, 5), (65, 40000, 9)])
def test_load_both_constraints_bit_exact(n_bins, res, seed):
    from hicdiff_b200 import prepare

    a = O.synthetic_contacts(n_bins, res, seed)
    b = O.synthetic_contacts(n_bins + 2, res, seed + 100)
    b[:, 2] = np.round(b[:, 2])
    ref = O.load_constraints(a, b, res)
    got = prepare.load_both_constraints(a, b, res, device=DEV)
    assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)


def test_stages_bit_exact_including_duplicates_and_nan_diagonals():
    from hicdiff_b200 import prepare

    rng = np.random.default_rng(0)
    n, nnz = 97, 4000
    rows, cols = rng.integers(5, 5 + n, nnz), rng.integers(5, 5 + n, nnz)          # many duplicated / mirrored cells
    vals = rng.gamma(2.0, 5.0, nnz)
    vals[rng.integers(0, nnz, 20)] = np.nan
    ref = np.zeros((n, n), np.float32)
    for r, c, v in zip(rows, cols, vals):                                           # PrepareData_linear.py:67-70
        ref[r - 5, c - 5] = v
        ref[c - 5, r - 5] = v
    got = prepare.dense_from_triples(rows, cols, vals, 5, n, DEV)
    assert np.array_equal(got.cpu().numpy(), ref, equal_nan=True)
    d = np.diag(ref)
    rm = np.unique(np.concatenate((np.argwhere(d == 0)[:, 0], np.argwhere(np.isnan(d))[:, 0])))
    want = np.delete(np.delete(ref, rm, axis=0), rm, axis=1)
    m2, kept = prepare.remove_empty_bins(got)
    assert np.array_equal(m2.cpu().numpy(), want, equal_nan=True)
    assert np.array_equal(kept.cpu().numpy(), np.setdiff1d(np.arange(n), rm))
    with pytest.raises(RuntimeError):
        prepare.dense_from_triples(np.array([0]), np.array([200]), np.array([1.0]), 0, 10, DEV)   # out of range -> loud


@pytest.mark.parametrize("n", [1, 2, 1000, 250_000, 6_000_000])
@pytest.mark.parametrize("q", [99.0, 50.0, 0.0, 100.0, 99.99])
def test_percentile_matches_numpy_bit_for_bit(n, q):
    from hicdiff_b200 import prepare

    rng = np.random.default_rng(n % 1000 + int(q))
    a = rng.gamma(2.0, 10.0, n).astype(np.float32)
    if n > 10:
        a[rng.integers(0, n, n // 3)] = 0.0                                         # ties, like a sparse contact map
        a[rng.integers(0, n, 5)] *= -1.0
    ref = np.percentile(a, q)
    got = prepare.percentile(torch.from_numpy(a).to(DEV), q)
    assert got == ref and np.asarray(got).dtype == np.asarray(ref).dtype, (got, ref)


def test_normalise_split_and_noise_bit_exact():
    from hicdiff_b200 import prepare

    rng = np.random.default_rng(4)
    a = rng.gamma(2.0, 10.0, (300, 300)).astype(np.float32)
    a = np.maximum(a, a.T)
    per = np.percentile(a, 99.0)
    ref = 2 * (np.clip(a, 0, per) / per) - 1.0
    got = prepare.normalize_contacts_(torch.from_numpy(a.copy()).to(DEV), per)
    assert np.array_equal(got.cpu().numpy(), ref)
    tiles_ref = O.split_pieces(ref, 64, 40000)
    z = torch.randn(tiles_ref.shape, generator=torch.Generator().manual_seed(1))
    noisy_ref = O.add_noise(torch.from_numpy(tiles_ref), 0.1, z)
    target, noisy = prepare.make_splits(got, res=40000, sigma_0=0.1, noise=z.to(DEV))
    assert np.array_equal(target.cpu().numpy(), tiles_ref)
    assert torch.equal(noisy.cpu(), noisy_ref)
    # Philox path: deterministic in the seed, right statistics
    t2, n1 = prepare.make_splits(got, sigma_0=0.1, seed=7)
    _, n2 = prepare.make_splits(got, sigma_0=0.1, seed=7)
    assert torch.equal(n1, n2)
    resid = (n1 - t2) / 0.1
    assert abs(float(resid.mean())) < 0.02 and abs(float(resid.std()) - 1.0) < 0.02
    with pytest.raises(RuntimeError):
        prepare.normalize_contacts_(torch.zeros(4, 4), 1.0)                          # CPU tensor: no fallback
