#!/usr/bin/env python
"""CPU study (oracle + emulated operand rounding): which rounding makes the bf16 path miss the 1e-3 SSIM / PSNR bar on the
CONTRACTIVE (tuned-tail) T = 1000 chain, and what a higher-precision GEMM mode would buy.

The fp32 oracle is run over the last N steps of the chain, restarted from the reference's own snapshot, with F.conv2d's
operands rounded as the device path rounds them:
  w     weights (after standardisation) -> bf16         a     conv inputs (= the stored activations) -> bf16
  wa    both (the shipped bf16 path, to first order)    tf32  both -> tf32 (10-bit mantissa)
  ao    conv inputs and conv outputs -> bf16 (every tensor the device path keeps in HBM), weights fp32
  x3    both split hi + lo in bf16, the lo*lo product dropped (3 MMAs per product: what a 'bf16x3' mode would do)
Prints final-tile RMS vs the reference fixture and |dSSIM| / |dPSNR| vs the clean target.   python scripts/precision_study.py
"""
import contextlib
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import helpers  # noqa: E402
from oracle import hicdiff_oracle as O  # noqa: E402


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def tf32(x):   # round-to-nearest-even onto a 10-bit mantissa
    i = x.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


@contextlib.contextmanager
def rounded_convs(mode):
    orig = F.conv2d
    state = {"i": 0}

    def conv(x, w, b=None, **kw):
        if w.shape[1] <= 2 and w.shape[-1] == 7:      # init_conv: conv #0 of a forward (75 convs; #71-73 = final_res_block, #74 = final_conv)
            state["i"] = 0
        idx = state["i"]
        state["i"] += 1
        if mode.startswith("ao_"):    # 'a' everywhere + the OUTPUT rounding of a subset of the convs
            is_ws = w.shape[-1] == 3 and abs(float(w.mean())) < 1e-6 and abs(float(w.var(unbiased=False)) - 1) < 5e-2   # Block.proj (var / (var + 1e-5) of a ~6e-4 variance)
            is_attn = w.shape[-1] == 1 and (w.shape[0] == 384 or w.shape[1] == 128)                                   # to_qkv / to_out
            sel = {"ao_init": idx == 0, "ao_tail": 71 <= idx <= 73, "ao_rest": 1 <= idx <= 70,
                   "ao_gn": is_ws and 1 <= idx <= 70, "ao_attn": is_attn and 1 <= idx <= 70,
                   "ao_other": (not is_ws) and (not is_attn) and 1 <= idx <= 70,
                   "ao_qkv": w.shape[-1] == 1 and w.shape[0] == 384, "ao_out": w.shape[-1] == 1 and w.shape[1] == 128 and w.shape[0] != 384}[mode]
            out = orig(bf16(x), w, b, **kw)
            return bf16(out) if sel else out
        if mode == "w":
            return orig(x, bf16(w), b, **kw)
        if mode == "a":
            return orig(bf16(x), w, b, **kw)
        if mode == "ao":     # activations rounded where the device path stores them: conv inputs AND conv outputs (not the 1-channel eps)
            out = orig(bf16(x), w, b, **kw)
            return out if out.shape[1] == 1 else bf16(out)
        if mode == "wa":
            return orig(bf16(x), bf16(w), b, **kw)
        if mode == "tf32":
            return orig(tf32(x), tf32(w), b, **kw)
        if mode == "x3":
            xh, wh = bf16(x), bf16(w)
            xl, wl = bf16(x - xh), bf16(w - wh)
            return orig(xh, wh, b, **kw) + orig(xl, wh, None, **kw) + orig(xh, wl, None, **kw)
        return orig(x, w, b, **kw)

    F.conv2d = conv
    try:
        yield
    finally:
        F.conv2d = orig


def main():
    start_t = int(sys.argv[1]) if len(sys.argv) > 1 else 250
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["none", "wa", "w", "a", "tf32", "x3"]
    torch.set_num_threads(8)
    gold = torch.load(helpers.GOLD / "t1000_unet_cond_tuned.pt")
    net, v = helpers.build_net("unet_cond")
    net.load_state_dict(torch.load(helpers.GOLD / "unet_cond_tuned_tail.pt")["tail"], strict=False)
    sd = {k: t.detach().clone() for k, t in net.state_dict().items()}
    T, B = 1000, 2
    clean, noisy = O.synthetic_tiles(B, seed=gold["tile_seed"])
    noise = O.synthetic_noise(T, B, seed=gold["noise_seed"])
    bufs = O.diffusion_buffers("sigmoid", T)
    eps_fn = helpers.oracle_eps_fn(sd, v["oracle"])
    hr = O.to_unit_range(clean)
    ref = gold["final"]
    s_ref, p_ref = float(O.ssim(O.to_unit_range(ref), hr)), float(O.psnr(O.to_unit_range(ref), hr))
    print(f"reference: SSIM {s_ref:.5f} PSNR {p_ref:.4f} dB; restart at t = {start_t}", flush=True)
    for mode in modes:
        x = (noise[0] if start_t >= T else gold["snapshots"][start_t]).clone()
        t0 = time.time()
        with torch.no_grad(), rounded_convs(mode):
            for t in reversed(range(0, min(start_t, T))):
                z = noise[T - t] if t > 0 else None
                x, _, _ = O.p_sample(eps_fn, bufs, x, t, noisy, z)
        rms = float((x - ref).pow(2).mean().sqrt())
        ds = abs(float(O.ssim(O.to_unit_range(x), hr)) - s_ref)
        dp = abs(float(O.psnr(O.to_unit_range(x), hr)) - p_ref)
        print(f"mode {mode:5s}: final RMS {rms:.3e}  |dSSIM| {ds:.2e}  |dPSNR| {dp:.2e} dB   ({time.time() - t0:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
