"""CPU tests of the oracle (the checker) against the golden fixtures written from the UNMODIFIED reference
(oracle/make_golden.py), the schedule known answers and, when /root/reference is present, the
reference itself."""
import hashlib

import numpy as np
import pytest
import torch

import helpers
from oracle import hicdiff_oracle as O


def _time(v, t, B):
    if v["oracle"]["sr3"]:
        lv = O.sr3_noise_levels(v["schedule"], 1000)
        return torch.FloatTensor([lv[t + 1]]).repeat(B, 1)
    return torch.full((B,), t, dtype=torch.long)


@pytest.mark.parametrize("name", ["unet_cond", "unet_uncond", "unet_sr3", "hicedrn_cond", "hicedrn_sr3"])
def test_oracle_eps_reproduces_reference_golden_bit_exact(name):
    net, v = helpers.build_net(name)
    sd = net.state_dict()
    assert helpers.sd_checksum(sd) == v["state_dict_sha256"], "seeded default init differs from the reference's"
    gold = torch.load(helpers.GOLD / f"{name}.pt")
    B = gold["x_t"].shape[0]
    _, noisy = O.synthetic_tiles(B, seed=gold["tile_seed"])
    cond = noisy if v["oracle"]["self_condition"] else None
    t = 37
    with torch.no_grad():
        eps = helpers.oracle_eps_fn(sd, v["oracle"])(gold["x_t"], _time(v, t, B), cond)
    assert torch.equal(eps, gold["eps"][t])


@pytest.mark.parametrize("name", ["unet_cond", "hicedrn_cond"])
def test_oracle_chain_reproduces_reference_golden_bit_exact(name):
    net, v = helpers.build_net(name)
    sd = net.state_dict()
    gold = torch.load(helpers.GOLD / f"{name}.pt")
    T, B = gold["chain_T"], gold["chain_final"].shape[0]
    _, noisy = O.synthetic_tiles(B, seed=gold["tile_seed"])
    noise = O.synthetic_noise(T, B, seed=gold["noise_seed"])
    with torch.no_grad():
        out = O.p_sample_loop(helpers.oracle_eps_fn(sd, v["oracle"]), O.diffusion_buffers(gold["chain_schedule"], T),
                              noisy, noise, timesteps=T)
    assert torch.equal(out, gold["chain_final"])


def test_oracle_training_loss_reproduces_reference_golden():
    net, v = helpers.build_net("unet_cond")
    gold = torch.load(helpers.GOLD / "unet_cond.pt")
    clean, noisy = O.synthetic_tiles(2, seed=gold["tile_seed"])
    nz = torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        loss = O.p_losses(helpers.oracle_eps_fn(net.state_dict(), v["oracle"]), O.diffusion_buffers("sigmoid", 1000),
                          noisy, clean, torch.tensor([500, 20]), nz, loss_type="l2")
    assert float(loss) == gold["loss"]


def test_schedule_known_answers():
    """SURVEY.md Appendix C.9 (derived by running the reference) + the manifest written by make_golden.py."""
    s = O.diffusion_buffers("sigmoid", 1000)
    assert s["betas"][0].item() == pytest.approx(3.002792e-4, rel=1e-6)
    assert s["betas"][999].item() == pytest.approx(0.999, rel=1e-6)
    assert s["sqrt_recip_alphas_cumprod"][999].item() == pytest.approx(1824.8693, rel=1e-6)
    assert s["posterior_log_variance_clipped"][1].item() == pytest.approx(-8.80093, rel=1e-5)
    assert s["posterior_log_variance_clipped"][0].item() == pytest.approx(np.log(1e-20), rel=1e-6)
    assert s["posterior_mean_coef1"][999].item() == pytest.approx(1.7311467e-2, rel=1e-6)
    l = O.diffusion_buffers("linear", 1000)
    assert l["betas"][0].item() == pytest.approx(1e-4, rel=1e-6) and l["betas"][999].item() == pytest.approx(0.02, rel=1e-6)
    assert l["sqrt_recip_alphas_cumprod"][999].item() == pytest.approx(157.41046, rel=1e-6)
    assert l["posterior_mean_coef2"][999].item() == pytest.approx(0.9899487, rel=1e-6)
    kat = helpers.MANIFEST["schedule_kat_T1000_idx_0_1_500_999"]
    for sched, table in kat.items():
        b = O.diffusion_buffers(sched, 1000)
        for k, vals in table.items():
            got = [float(b[k][i]) for i in (0, 1, 500, 999)]
            assert got == vals, (sched, k)
    with pytest.raises(ValueError):
        O.beta_schedule("quadratic", 10)


def test_sr3_noise_level_table_off_by_one():
    lv = O.sr3_noise_levels("linear", 1000)           # hicdiff_sr3.py:535-536
    ac = O.diffusion_buffers("linear", 1000)["alphas_cumprod"].double()
    assert lv.shape == (1001,) and lv[0] == 1 and lv[1] == 1
    assert torch.allclose(lv[2:], ac[:-1].sqrt(), rtol=1e-6)


@pytest.mark.parametrize("n", [34, 64, 130, 588, 703])
@pytest.mark.parametrize("res", [40000, 10000])
def test_split_pieces_matches_reference_fixture_and_round_trips(n, res):
    rng = np.random.default_rng(helpers.MANIFEST["tiles"]["seed"])
    mats = {}
    for m in (34, 64, 130, 588, 703):   # same draw order as make_golden.py
        a = rng.standard_normal((m, m)).astype(np.float32)
        mats[m] = (a + a.T) * 0.5
    a = mats[n]
    tiles = O.split_pieces(a, 64, res)
    case = helpers.MANIFEST["tiles"]["cases"][f"{n}_{res}"]
    assert tiles.shape[0] == case["count"]
    assert hashlib.sha256(np.ascontiguousarray(tiles).tobytes()).hexdigest() == case["sha256"]
    back = O.reassemble(tiles, n, 64, res)
    assert np.array_equal(O.split_pieces(back, 64, res), tiles)                 # splitPieces(reassemble(t)) == t
    band = O.band_blocks_for(res) * 64 + 63
    i, j = np.indices((n, n))
    blk = np.abs(i // 64 - j // 64) <= O.band_blocks_for(res)
    assert np.array_equal(back[blk], a[blk]) and not back[~blk].any()           # reassemble(splitPieces(M)) == M in band
    assert band > 0


def test_split_pieces_edge_cases():
    assert O.split_pieces(np.zeros((0, 0), np.float32)).shape == (0, 1, 64, 64)      # empty chromosome
    one = O.split_pieces(np.ones((1, 1), np.float32))
    assert one.shape == (1, 1, 64, 64) and one.sum() == 1                           # ragged: padded to one tile
    P = 7
    full = O.split_pieces(np.zeros((64 * P, 64 * P), np.float32))
    assert full.shape[0] == 5 * P - 10                                               # 5P - 10 for P >= 5
    assert O.band_blocks_for(40000) == 4 and O.band_blocks_for(10000) == 16


@pytest.mark.reference
def test_oracle_against_live_reference_unet_cond():
    """When the upstream checkout is present, re-pin directly (fresh inputs, not the stored fixture)."""
    import sys

    sys.path.insert(0, "/root/reference")
    from src.hicdiff_condition import Unet as RefUnet

    torch.manual_seed(3)
    ref = RefUnet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True).eval()
    g = torch.Generator().manual_seed(8)
    x, c = torch.randn(1, 1, 64, 64, generator=g), torch.randn(1, 1, 64, 64, generator=g)
    t = torch.tensor([123])
    with torch.no_grad():
        assert torch.equal(ref(x, t, c), O.unet_forward(ref.state_dict(), x, t, c, self_condition=True))


@pytest.mark.parametrize("name", ["cond_l2", "uncond_l1", "sr3_l2", "unet_cond_l2", "unet_uncond_l1", "unet_sr3_l2"])
def test_oracle_training_gradients_reproduce_reference_golden(name):
    """p_losses_and_grads against the per-parameter summaries of the reference's own loss.backward()
    (oracle/make_golden_train.py asserted bit-equality of the FULL gradients when it wrote the fixture)."""
    import json

    gold = json.loads((helpers.GOLD / "hicedrn_train.json").read_text())
    c = gold["cases"][name]
    from hicdiff_b200.model.hicedrn_Diff import hicedrn_Diff

    from hicdiff_b200.model.hicedrn_sr3_Diff import hicedrn_Diff as hicedrn_sr3

    torch.manual_seed(gold["weight_seed"])
    sr3 = c["flavour"] == "sr3"
    unet = name.startswith("unet")
    if unet:
        from hicdiff_b200 import hicdiff, hicdiff_condition

        if sr3:
            from hicdiff_b200 import hicdiff_sr3

            net = hicdiff_sr3.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True, noise_level_emb=True)
        else:
            net = (hicdiff_condition if c["self_condition"] else hicdiff).Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=c["self_condition"])
    else:
        net = (hicedrn_sr3 if sr3 else hicedrn_Diff)(number_resnet=c["blocks"], self_condition=c["self_condition"])
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    clean, noisy = O.synthetic_tiles(c["B"], seed=gold["tile_seed"])
    noise = torch.randn(c["B"], 1, 64, 64, generator=torch.Generator().manual_seed(gold["noise_seed"]))
    if sr3:
        loss, grads = O.sr3_p_losses_and_grads(sd, noisy, clean, torch.tensor(c["level"], dtype=torch.float32), noise,
                                               loss_type=c["loss_type"], self_condition=True, num_blocks=c.get("blocks", 0),
                                               net="unet" if unet else "hicedrn")
    else:
        t = torch.tensor(c["t"], dtype=torch.long)
        buf = O.diffusion_buffers(c["schedule"], c["T"])
        loss, grads = O.p_losses_and_grads(sd, buf, noisy, clean, t, noise, loss_type=c["loss_type"],
                                           self_condition=c["self_condition"], num_blocks=c.get("blocks", 0),
                                           net="unet" if unet else "hicedrn")
    assert abs(float(loss) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"]))
    assert set(grads) == set(c["grads"])
    for k, s in c["grads"].items():
        g = grads[k].reshape(-1)
        got = g[torch.tensor(s["idx"])].double()
        want = torch.tensor(s["val"], dtype=torch.float64)
        scale = max(s["norm"] / max(g.numel(), 1) ** 0.5, 1e-12)      # RMS of the tensor: thread-count-dependent fp32 order only
        assert float((got - want).abs().max()) <= 1e-3 * scale + 1e-9, k
        assert abs(float(g.double().norm()) - s["norm"]) <= 1e-4 * s["norm"] + 1e-12, k


def test_oracle_data_preparation_reproduces_reference_golden():
    """load_constraints against the checksums of the reference's own loadBothConstraints (oracle/make_golden_prepare.py).
    The percentile rule travels with the numpy version (float32 index arithmetic in numpy >= 2), so the checksum is only
    comparable under the numpy major version that wrote it."""
    import json

    gold = json.loads((helpers.GOLD / "prepare.json").read_text())
    if gold["numpy"].split(".")[0] != np.__version__.split(".")[0]:
        pytest.skip("fixture written under a different numpy major version")
    for c in gold["cases"]:
        a = O.synthetic_contacts(c["n_bins"], c["res"], c["seed"])
        b = O.synthetic_contacts(c["n_bins"] + 2, c["res"], c["seed"] + 100)
        b[:, 2] = np.round(b[:, 2])
        m = O.load_constraints(a, b, c["res"])
        assert list(m.shape) == c["shape"] and m.dtype == np.float32
        assert hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest() == c["sha256"]
        assert m.min() == -1.0 and m.max() == 1.0 and np.array_equal(m, m.T)
