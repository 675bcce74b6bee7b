"""Shared helpers: seeded model construction (weights reproduce the reference's default init under WEIGHT_SEED,
verified by oracle/make_golden.py and re-checked here through the manifest's SHA-256) and oracle plumbing."""
import hashlib
import json
from pathlib import Path

import torch

GOLD = Path(__file__).resolve().parent / "golden"
MANIFEST = json.loads((GOLD / "manifest.json").read_text())


def sd_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_net(name: str):
    """hicdiff_b200 parameter-holder net for a manifest variant, default-initialised under the golden seed (CPU)."""
    from hicdiff_b200 import hicdiff, hicdiff_condition, hicdiff_sr3
    from hicdiff_b200.model import hicedrn_Diff, hicedrn_sr3_Diff

    ctor = {
        "unet_cond": hicdiff_condition.Unet, "unet_uncond": hicdiff.Unet, "unet_sr3": hicdiff_sr3.Unet,
        "hicedrn_cond": hicedrn_Diff.hicedrn_Diff, "hicedrn_sr3": hicedrn_sr3_Diff.hicedrn_Diff,
    }[name]
    v = MANIFEST["variants"][name]
    kw = dict(v["net_kwargs"])
    if "dim_mults" in kw:
        kw["dim_mults"] = tuple(kw["dim_mults"])
    torch.manual_seed(MANIFEST["weight_seed"])
    net = ctor(**kw)
    return net, v


def diffusion_cls(name: str):
    from hicdiff_b200 import hicdiff, hicdiff_condition, hicdiff_sr3

    return {"unet_cond": hicdiff_condition.GaussianDiffusion, "unet_uncond": hicdiff.GaussianDiffusion,
            "unet_sr3": hicdiff_sr3.GaussianDiffusion, "hicedrn_cond": hicdiff_condition.GaussianDiffusion,
            "hicedrn_sr3": hicdiff_sr3.GaussianDiffusion}[name]


def oracle_eps_fn(sd, okw, taps=None):
    from oracle import hicdiff_oracle as O

    if okw["kind"] == "unet":
        return lambda x, t, c: O.unet_forward(sd, x, t, c, self_condition=okw["self_condition"], sr3=okw["sr3"], taps=taps)
    return lambda x, t, c: O.hicedrn_forward(sd, x, t, c, self_condition=okw["self_condition"], sr3=okw["sr3"], taps=taps)


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))
