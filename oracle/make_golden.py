"""Pins the oracle against the UNMODIFIED reference and writes the golden fixtures under tests/golden/.

Run here (the container that has /root/reference):   python oracle/make_golden.py
It cannot run on the GPU box (no reference there); the fixtures it writes are what travels.

For every eps-net variant it
  1. builds the reference module and the hicdiff_b200 parameter holder under the same torch seed and checks that their
     state_dicts are identical key-for-key and bit-for-bit (so a seed reproduces the weights anywhere);
  2. checks oracle eps == reference eps bit-for-bit (teacher-forced single step, B=2);
  3. checks oracle p_sample_loop == reference super_resolution/sample bit-for-bit with injected noise
     (torch.randn / randn_like patched to pop from one shared iterator -- exactly T draws);
  4. checks oracle split_pieces == reference splitPieces (the four missing imports stubbed);
and stores inputs seeds, weight checksums and the reference outputs.
"""
from __future__ import annotations

import contextlib
import hashlib
import json
import os
import sys
import tempfile
import types
from pathlib import Path

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import hicdiff_oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden"
WEIGHT_SEED = 0


def sd_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@contextlib.contextmanager
def injected_noise(noise):
    it = iter(noise)
    orig = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig


@contextlib.contextmanager
def quiet():
    import io

    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def variants():
    from hicdiff_b200 import hicdiff as B_u, hicdiff_condition as B_c, hicdiff_sr3 as B_s
    from hicdiff_b200.model import hicedrn_Diff as B_h, hicedrn_sr3_Diff as B_hs
    from src import hicdiff as R_u, hicdiff_condition as R_c, hicdiff_sr3 as R_s
    from src.model import hicedrn_Diff as R_h, hicedrn_sr3_Diff as R_hs

    unet = dict(dim=64, dim_mults=(1, 2, 4, 8))
    return [
        # name, ref net ctor, our net ctor, ref diffusion, our diffusion, net kwargs, schedule, oracle forward kwargs
        ("unet_cond", R_c.Unet, B_c.Unet, R_c.GaussianDiffusion, B_c.GaussianDiffusion,
         dict(unet, self_condition=True), "sigmoid", dict(kind="unet", self_condition=True, sr3=False)),
        ("unet_uncond", R_u.Unet, B_u.Unet, R_u.GaussianDiffusion, B_u.GaussianDiffusion,
         dict(unet, self_condition=False), "linear", dict(kind="unet", self_condition=False, sr3=False)),
        ("unet_sr3", R_s.Unet, B_s.Unet, R_s.GaussianDiffusion, B_s.GaussianDiffusion,
         dict(unet, self_condition=True, noise_level_emb=True), "linear", dict(kind="unet", self_condition=True, sr3=True)),
        ("hicedrn_cond", R_h.hicedrn_Diff, B_h.hicedrn_Diff, R_c.GaussianDiffusion, B_c.GaussianDiffusion,
         dict(self_condition=True), "sigmoid", dict(kind="hicedrn", self_condition=True, sr3=False)),
        ("hicedrn_sr3", R_hs.hicedrn_Diff, B_hs.hicedrn_Diff, R_s.GaussianDiffusion, B_s.GaussianDiffusion,
         dict(self_condition=True, noise_level_emb=True), "linear", dict(kind="hicedrn", self_condition=True, sr3=True)),
    ]


def oracle_eps_fn(sd, okw):
    if okw["kind"] == "unet":
        return lambda x, t, c: O.unet_forward(sd, x, t, c, self_condition=okw["self_condition"], sr3=okw["sr3"])
    return lambda x, t, c: O.hicedrn_forward(sd, x, t, c, self_condition=okw["self_condition"], sr3=okw["sr3"])


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    GOLD.mkdir(parents=True, exist_ok=True)
    manifest = {"weight_seed": WEIGHT_SEED, "torch": torch.__version__, "variants": {}}
    B = 2
    clean, noisy = O.synthetic_tiles(B, seed=1234)
    for name, RNet, BNet, RDiff, BDiff, nkw, sched, okw in variants():
        print(f"== {name}")
        torch.manual_seed(WEIGHT_SEED)
        rnet = RNet(**nkw).eval()
        torch.manual_seed(WEIGHT_SEED)
        bnet = BNet(**nkw).eval()
        rsd, bsd = rnet.state_dict(), bnet.state_dict()
        assert list(rsd.keys()) == list(bsd.keys()), f"{name}: state_dict keys differ"
        assert all(torch.equal(rsd[k], bsd[k]) for k in rsd), f"{name}: seeded init differs from the reference"
        T_chain = 40 if okw["kind"] == "unet" else 6
        chain_sched = "sigmoid" if sched == "sigmoid" else "cosine"   # linear is invalid for tiny T
        rdiff = RDiff(rnet, image_size=64, timesteps=1000, loss_type="l2", beta_schedule=sched).eval()
        bdiff = BDiff(bnet, image_size=64, timesteps=1000, loss_type="l2", beta_schedule=sched)
        ra, ba = rdiff.state_dict(), bdiff.state_dict()
        assert list(ra.keys()) == list(ba.keys()) and all(torch.equal(ra[k], ba[k]) for k in ra), f"{name}: diffusion state_dict"
        buf = O.diffusion_buffers(sched, 1000)
        assert all(torch.equal(buf[k], ra[k]) for k in buf), f"{name}: oracle schedule buffers differ"

        # ---- teacher-forced eps at two timesteps
        eps_fn = oracle_eps_fn(rsd, okw)
        g = torch.Generator().manual_seed(77)
        x_t = torch.randn(B, 1, 64, 64, generator=g)
        cond = noisy if okw["self_condition"] else None
        eps_gold = {}
        for t in (999, 37):
            if okw["sr3"]:
                lv = O.sr3_noise_levels(sched, 1000)
                assert torch.equal(lv, rdiff.sqrt_alphas_cumprod_prev), f"{name}: sr3 level table"
                time = torch.FloatTensor([lv[t + 1]]).repeat(B, 1)
            else:
                time = torch.full((B,), t, dtype=torch.long)
            with torch.no_grad():
                ref = rnet(x_t, time, cond)
                ora = eps_fn(x_t, time, cond)
            assert torch.equal(ref, ora), f"{name}: oracle eps != reference eps at t={t} (max {float((ref - ora).abs().max()):.3e})"
            eps_gold[t] = ref.clone()
            print(f"   eps t={t}: bit-exact, rms {ref.pow(2).mean().sqrt():.4f}")

        # ---- free-running short chain with injected noise
        rdiff_s = RDiff(rnet, image_size=64, timesteps=T_chain, loss_type="l2", beta_schedule=chain_sched).eval()
        noise = O.synthetic_noise(T_chain, B, seed=2024)
        with torch.no_grad(), injected_noise(noise), quiet():
            if okw["self_condition"]:
                ref_final = rdiff_s.super_resolution(noisy)
            else:
                ref_final = rdiff_s.sample(noisy)
        bufs = O.diffusion_buffers(chain_sched, T_chain)
        levels = O.sr3_noise_levels(chain_sched, T_chain) if okw["sr3"] else None
        with torch.no_grad():
            ora_final = O.p_sample_loop(eps_fn, bufs, cond, noise, timesteps=T_chain, sr3_levels=levels)
        assert torch.equal(ref_final, ora_final), f"{name}: oracle chain != reference chain (max {float((ref_final - ora_final).abs().max()):.3e})"
        print(f"   chain T={T_chain} ({chain_sched}): bit-exact, range [{ref_final.min():.3f}, {ref_final.max():.3f}]")

        # ---- training objective value (conditional flavours: t / noise injected)
        loss_gold = None
        if name in ("unet_cond", "hicedrn_cond"):
            tt = torch.tensor([500, 20])
            nz = torch.randn(B, 1, 64, 64, generator=torch.Generator().manual_seed(5))
            orig = torch.randint
            torch.randint = lambda *a, **k: tt.clone()
            try:
                with torch.no_grad():
                    ref_loss = rdiff([noisy, clean], noise=nz)
            finally:
                torch.randint = orig
            with torch.no_grad():
                ora_loss = O.p_losses(eps_fn, buf, noisy, clean, tt, nz, loss_type="l2", self_condition=True)
            assert torch.equal(ref_loss, ora_loss), f"{name}: p_losses {ref_loss} vs {ora_loss}"
            loss_gold = float(ref_loss)
            print(f"   p_losses: bit-exact {loss_gold:.6f}")

        torch.save({
            "x_t": x_t, "eps": eps_gold, "chain_final": ref_final, "chain_T": T_chain, "chain_schedule": chain_sched,
            "tile_seed": 1234, "noise_seed": 2024, "x_t_seed": 77, "loss": loss_gold,
        }, GOLD / f"{name}.pt")
        manifest["variants"][name] = {
            "net_kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in nkw.items()},
            "schedule": sched, "state_dict_sha256": sd_checksum(rsd), "n_params": int(sum(v.numel() for v in rsd.values())),
            "oracle": okw,
        }

    # ---- schedule known answers
    kat = {}
    for sched in ("sigmoid", "linear", "cosine"):
        b = O.diffusion_buffers(sched, 1000)
        kat[sched] = {k: [float(v[0]), float(v[1]), float(v[500]), float(v[999])] for k, v in b.items()}
    manifest["schedule_kat_T1000_idx_0_1_500_999"] = kat

    # ---- tiling against the reference's splitPieces
    for m in ("pyrootutils", "pytorch_lightning", "cooler", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["pyrootutils"].setup_root = lambda **k: str(REF)
    sys.modules["pytorch_lightning"].LightningDataModule = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from processdata.PrepareData_linear import splitPieces  # noqa: E402

    tiles_gold = {}
    rng = np.random.default_rng(7)
    for n in (34, 64, 130, 588, 703):
        a = rng.standard_normal((n, n)).astype(np.float32)
        a = (a + a.T) * 0.5
        with tempfile.TemporaryDirectory() as td:
            fn = os.path.join(td, "m.npy")
            np.save(fn, a)
            for res in (40000, 10000):
                ref = splitPieces(fn, 64, 64, res)
                ora = O.split_pieces(a, 64, res)
                assert ref.shape == ora.shape and np.array_equal(ref, ora), f"splitPieces mismatch n={n} res={res}"
                back = O.reassemble(ora, n, 64, res)
                again = O.split_pieces(back, 64, res)
                assert np.array_equal(again, ora), "reassemble is not the inverse of split_pieces"
                tiles_gold[f"{n}_{res}"] = {"count": int(ref.shape[0]),
                                            "sha256": hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest()}
        print(f"   splitPieces n={n}: ok ({tiles_gold[f'{n}_40000']['count']} tiles @40kb)")
    manifest["tiles"] = {"seed": 7, "cases": tiles_gold}
    (GOLD / "manifest.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
