"""Per-kernel parity tests (B200 only): every sm_100a kernel the plan launches, called through the C ABI
(`hd_op_*`) and compared with a float64 torch evaluation of the reference's op on the SAME bf16-rounded inputs.

Tolerances (stated, bf16 storage with fp32 accumulation): the only error sources are accumulation order and the
final bf16 rounding of the output (2^-9 relative), so
    rel-RMS(out - ref) <= 4e-3   and   max|out - ref| <= 2e-2 * max|ref| + 1e-3.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from hicdiff_b200 import ops

    return ops


def _rand_nhwc(B, H, W, C, g, scale=1.0):
    return (torch.randn(B, H, W, C, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _nchw64(x):  # NHWC bf16 -> NCHW float64
    return x.permute(0, 3, 1, 2).to(torch.float64)


def _check(out_nhwc, ref_nchw, what, rel_rms=4e-3, max_rel=2e-2):
    out = _nchw64(out_nhwc)
    assert out.shape == ref_nchw.shape, f"{what}: shape {tuple(out.shape)} vs {tuple(ref_nchw.shape)}"
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    err = out - ref_nchw
    rms = err.pow(2).mean().sqrt().item()
    ref_rms = ref_nchw.pow(2).mean().sqrt().item()
    mx = err.abs().max().item()
    assert rms <= rel_rms * ref_rms + 1e-6, f"{what}: rel-RMS {rms / max(ref_rms, 1e-30):.3e} (rms {rms:.3e}, ref {ref_rms:.3e})"
    assert mx <= max_rel * ref_nchw.abs().max().item() + 1e-3, f"{what}: max abs err {mx:.3e}"


def _ws(w):  # WeightStandardizedConv2d, fp64
    w = w.to(torch.float64)
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + 1e-5).rsqrt()


def _bf16_round(w):
    return w.to(torch.bfloat16).to(torch.float64)


CONV_CASES = [
    # B, H, C0, C1, Cout, k, standardize
    (2, 64, 64, 0, 64, 3, True),
    (1, 64, 64, 64, 64, 3, True),       # concat 128 -> 64
    (2, 32, 128, 64, 128, 3, True),     # concat 192 -> 128
    (3, 16, 256, 0, 256, 3, True),
    (2, 8, 512, 0, 512, 3, True),
    (1, 8, 256, 0, 512, 3, False),      # M = 64 < one tile (TMA OOB rows), plain conv (downs.3.3)
    (3, 8, 512, 256, 512, 3, True),     # M = 192: partial last tile, concat 768
    (2, 64, 64, 0, 384, 1, False),      # to_qkv
    (2, 32, 128, 0, 64, 1, False),      # to_out
    (2, 16, 256, 128, 256, 1, False),   # res_conv on a concat
    (2, 64, 256, 0, 256, 3, False),     # HiCEDRN body conv
    (3, 32, 64, 0, 64, 3, True),        # level-1 64 -> 64: padded-slab path with 34-pixel row pitch, 9 tiles per image
    (2, 32, 64, 64, 64, 3, True),
    (5, 64, 64, 0, 64, 3, False),       # odd batch on the padded-slab path (33 tiles per image)
]


@pytest.mark.parametrize("B,H,C0,C1,Cout,k,std", CONV_CASES)
def test_conv_gemm(B, H, C0, C1, Cout, k, std):
    ops = _ops()
    g = torch.Generator().manual_seed(1000 + B * 7 + H + C0 + C1 + Cout + k)
    x0 = _rand_nhwc(B, H, H, C0, g)
    x1 = _rand_nhwc(B, H, H, C1, g) if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).to(DEV)
    bias = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x0, w, bias, x1=x1, standardize=std)
    torch.cuda.synchronize()
    wr = _bf16_round(_ws(w).float() if std else w)
    xin = _nchw64(x0) if x1 is None else torch.cat((_nchw64(x0), _nchw64(x1)), dim=1)
    ref = F.conv2d(xin, wr, bias.to(torch.float64), padding=k // 2)
    _check(out, ref, f"conv {C0}+{C1}->{Cout} k{k} @{H}")


@pytest.mark.parametrize("B,Hl,C,Cout", [(2, 32, 128, 64), (2, 16, 256, 128), (2, 8, 512, 256), (3, 8, 512, 256)])
def test_upsample_conv(B, Hl, C, Cout):
    """Upsample = nn.Upsample(scale_factor=2, mode='nearest') + Conv2d(3, padding=1) (hicdiff_condition.py:72-76), run as
    four 2x2 phase convs over the low-res input.  The folded weights are sums of up to four taps rounded to bf16 once,
    so the reference convolves with the fp64 weights and the tolerance carries one extra bf16 weight rounding."""
    ops = _ops()
    g = torch.Generator().manual_seed(300 + Hl + B)
    x = _rand_nhwc(B, Hl, Hl, C, g)
    w = (torch.randn(Cout, C, 3, 3, generator=g) / math.sqrt(C * 9)).to(DEV)
    bias = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x, w, bias, upsample=True)
    torch.cuda.synchronize()
    up = F.interpolate(_nchw64(x), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w.to(torch.float64), bias.to(torch.float64), padding=1)
    _check(out, ref, f"upsample conv {C}->{Cout} @{Hl}->{2 * Hl}", rel_rms=6e-3)


@pytest.mark.parametrize("B,H,C0,C1", [(2, 64, 64, 0), (5, 64, 64, 0), (1, 64, 64, 64), (3, 32, 64, 0), (2, 32, 64, 64)])
def test_conv_gemm_padded_slab(B, H, C0, C1):
    """Opt-in padded-slab form of the 3x3, Cout = 64 conv: one TMA box {64 ch, W + 2, rows + 2} per chunk, the nine taps are
    row offsets into it (A descriptors start at arbitrary 128-byte rows: the tensor core takes the swizzle phase from the
    absolute smem address), halo positions are masked / compacted away in the epilogue."""
    ops = _ops()
    g = torch.Generator().manual_seed(4242 + B + H + C1)
    x0 = _rand_nhwc(B, H, H, C0, g)
    x1 = _rand_nhwc(B, H, H, C1, g) if C1 else None
    Cin = C0 + C1
    w = (torch.randn(64, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).to(DEV)
    bias = (torch.randn(64, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x0, w, bias, x1=x1, standardize=True, pad_mode=2)
    torch.cuda.synchronize()
    xin = _nchw64(x0) if x1 is None else torch.cat((_nchw64(x0), _nchw64(x1)), dim=1)
    ref = F.conv2d(xin, _bf16_round(_ws(w).float()), bias.to(torch.float64), padding=1)
    _check(out, ref, f"padded-slab conv {Cin}->64 @{H}")
    same = ops.conv2d_nhwc(x0, w, bias, x1=x1, standardize=True, pad_mode=0)
    assert (out.float() - same.float()).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,C0,C1,res", [(2, 64, 64, 0, False), (5, 64, 64, 64, True), (16, 64, 64, 64, False), (24, 64, 64, 0, True),
                                           (3, 32, 64, 0, False), (9, 32, 64, 64, True), (8, 16, 64, 0, False), (40, 16, 64, 64, True)])
def test_conv_gemm_dx_stacked(B, H, C0, C1, res):
    """Default form of the 3x3, Cout = 64 conv (hicdiff_condition.py:84-97 at the 64-channel level): the three dx taps of a filter
    row are ONE tcgen05.mma of N = 192 over the un-shifted pixel slab, and the epilogue adds the three 64-column groups shifted by
    -1 / 0 / +1 pixel (neighbours dropped on the zero padding; across TMEM lane quarters the rows go through shared memory at
    W = 64).  Checked against fp64 torch and against the one-MMA-per-tap form; several tiles per persistent CTA (accumulator and
    slab rings wrap), concat sources, residual epilogue, W = 64 / 32 / 16."""
    ops = _ops()
    g = torch.Generator().manual_seed(5150 + B + H + C1)
    x0 = _rand_nhwc(B, H, H, C0, g)
    x1 = _rand_nhwc(B, H, H, C1, g) if C1 else None
    r = _rand_nhwc(B, H, H, 64, g) if res else None
    Cin = C0 + C1
    w = (torch.randn(64, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).to(DEV)
    bias = (torch.randn(64, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x0, w, bias, x1=x1, res=r, standardize=True)
    torch.cuda.synchronize()
    xin = _nchw64(x0) if x1 is None else torch.cat((_nchw64(x0), _nchw64(x1)), dim=1)
    ref = F.conv2d(xin, _bf16_round(_ws(w).float()), bias.to(torch.float64), padding=1)
    if res:
        ref = ref + _nchw64(r)
    _check(out, ref, f"dx-stacked conv {Cin}->64 @{H}")
    per_tap = ops.conv2d_nhwc(x0, w, bias, x1=x1, res=r, standardize=True, dx_stack=False)
    _check(per_tap, ref, f"per-tap conv {Cin}->64 @{H}")
    # same products, fp32 accumulation in a different order, one bf16 rounding at the end: at most one bf16 ulp apart
    assert (out.float() - per_tap.float()).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()
    again = ops.conv2d_nhwc(x0, w, bias, x1=x1, res=r, standardize=True)
    assert torch.equal(out, again), "dx-stacked conv is not bit-reproducible"
    one_group = ops.conv2d_nhwc(x0, w, bias, x1=x1, res=r, standardize=True, dx_stack=1)
    _check(one_group, ref, f"dx-stacked (one epilogue group) conv {Cin}->64 @{H}")
    assert (out.float() - one_group.float()).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,C0,C1,Cout,k", [(2, 64, 64, 0, 64, 3), (3, 32, 128, 64, 128, 3), (3, 8, 512, 256, 512, 3), (2, 16, 256, 0, 256, 3),
                                               (2, 64, 64, 0, 384, 1), (1, 8, 256, 0, 512, 3)])
def test_conv_gemm_cta_pairs(B, H, C0, C1, Cout, k):
    """Opt-in CTA-pair form (cluster of 2, tcgen05.mma.cta_group::2): M = 256 tile pairs, each CTA loads its 128 rows of A and
    half of the weight rows, TMA completions land on the leader's barriers, commits are multicast.  Includes an odd number of
    M tiles (phantom tile in the last pair) and the resident / slab / general kinds."""
    ops = _ops()
    g = torch.Generator().manual_seed(777 + B + H + C0 + Cout)
    x0 = _rand_nhwc(B, H, H, C0, g)
    x1 = _rand_nhwc(B, H, H, C1, g) if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).to(DEV)
    bias = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x0, w, bias, x1=x1, cta_pairs=True)
    torch.cuda.synchronize()
    xin = _nchw64(x0) if x1 is None else torch.cat((_nchw64(x0), _nchw64(x1)), dim=1)
    ref = F.conv2d(xin, _bf16_round(w), bias.to(torch.float64), padding=k // 2)
    _check(out, ref, f"cta-pair conv {Cin}->{Cout} k{k} @{H}")


def test_conv_gemm_residual_epilogue():
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    x = _rand_nhwc(2, 8, 8, 128, g)
    res = _rand_nhwc(2, 8, 8, 512, g)
    w = (torch.randn(512, 128, 1, 1, generator=g) / math.sqrt(128)).to(DEV)
    bias = (torch.randn(512, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x, w, bias, res=res)
    ref = F.conv2d(_nchw64(x), _bf16_round(w), bias.to(torch.float64)) + _nchw64(res)
    _check(out, ref, "conv 1x1 + residual")


@pytest.mark.parametrize("B,Hout,C,Cout", [(2, 32, 64, 64), (2, 16, 64, 128), (3, 8, 128, 256)])
def test_downsample_unshuffle(B, Hout, C, Cout):
    from einops import rearrange

    ops = _ops()
    g = torch.Generator().manual_seed(11 + Hout)
    x = _rand_nhwc(B, 2 * Hout, 2 * Hout, C, g)
    w = (torch.randn(Cout, 4 * C, 1, 1, generator=g) / math.sqrt(4 * C)).to(DEV)
    bias = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    out = ops.conv2d_nhwc(x, w, bias, unshuffle=True, ksize=1)
    xin = rearrange(_nchw64(x), "b c (h p1) (w p2) -> b (c p1 p2) h w", p1=2, p2=2)
    ref = F.conv2d(xin, _bf16_round(w), bias.to(torch.float64))
    _check(out, ref, f"downsample {C}->{Cout} @{Hout}")


@pytest.mark.parametrize("B,H,C,film,res", [(2, 64, 64, True, False), (3, 32, 128, False, True), (2, 16, 256, True, True),
                                            (5, 8, 512, True, False), (1, 32, 256, False, False), (1, 64, 256, False, True)])
def test_groupnorm_film_silu(B, H, C, film, res):
    ops = _ops()
    g = torch.Generator().manual_seed(5 + C)
    x = _rand_nhwc(B, H, H, C, g, scale=1.7) + 0.3
    x = x.to(torch.bfloat16)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    scale = (0.3 * torch.randn(C, generator=g)).to(DEV) if film else None
    shift = (0.3 * torch.randn(C, generator=g)).to(DEV) if film else None
    r = _rand_nhwc(B, H, H, C, g) if res else None
    out = ops.groupnorm_silu_nhwc(x, gamma, beta, scale, shift, r)
    ref = F.group_norm(_nchw64(x), 8, gamma.double(), beta.double(), eps=1e-5)
    if film:
        ref = ref * (scale.double().view(1, -1, 1, 1) + 1) + shift.double().view(1, -1, 1, 1)
    ref = F.silu(ref)
    if res:
        ref = ref + _nchw64(r)
    _check(out, ref, f"groupnorm C{C} @{H}")


@pytest.mark.parametrize("B,H,C0,C1,Cout,film,res", [
    (3, 64, 64, 0, 64, True, False),      # block1 at level 0 (weights resident in smem)
    (2, 64, 64, 64, 64, True, False),     # concat 128 -> 64 (K = 1152, resident)
    (2, 64, 64, 0, 64, False, True),      # block2: no FiLM, + residual
    (5, 32, 128, 64, 128, True, False),   # 192 -> 128, streaming weights, 128-wide tile
    (2, 32, 128, 0, 128, False, True),
    (3, 16, 128, 0, 128, True, True),
    (160, 16, 128, 0, 128, False, True),  # more tiles than SMs: several phase-B rounds per CTA
])
def test_conv_groupnorm_fused(B, H, C0, C1, Cout, film, res):
    """conv -> GroupNorm(8) -> FiLM -> SiLU (+res) in ONE launch vs fp64 torch on the same bf16 inputs."""
    ops = _ops()
    g = torch.Generator().manual_seed(900 + B + H + C0 + C1 + Cout)
    x0 = _rand_nhwc(B, H, H, C0, g)
    x1 = _rand_nhwc(B, H, H, C1, g) if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).to(DEV)
    bias = (torch.randn(Cout, generator=g) * 0.3).to(DEV)
    gamma = (1 + 0.2 * torch.randn(Cout, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(Cout, generator=g)).to(DEV)
    scale = (0.3 * torch.randn(Cout, generator=g)).to(DEV) if film else None
    shift = (0.3 * torch.randn(Cout, generator=g)).to(DEV) if film else None
    r = _rand_nhwc(B, H, H, Cout, g) if res else None
    out = ops.conv_gn_nhwc(x0, w, bias, gamma, beta, x1=x1, scale=scale, shift=shift, res=r, standardize=True)
    torch.cuda.synchronize()
    xin = _nchw64(x0) if x1 is None else torch.cat((_nchw64(x0), _nchw64(x1)), dim=1)
    y = F.conv2d(xin, _bf16_round(_ws(w).float()), bias.to(torch.float64), padding=1)
    y = F.group_norm(y, 8, gamma.to(torch.float64), beta.to(torch.float64), eps=1e-5)
    if film:
        y = y * (scale.to(torch.float64).view(1, -1, 1, 1) + 1) + shift.to(torch.float64).view(1, -1, 1, 1)
    y = F.silu(y)
    if res:
        y = y + _nchw64(r)
    _check(out, y, f"fused conv+GN {Cin}->{Cout} @{H}")


def test_groupnorm_rejects_unsupported_channel_counts():
    """Channel counts the kernel is not built for must fail loudly, never silently compute something else."""
    ops = _ops()
    x = torch.zeros(1, 64, 64, 96, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError):
        ops.groupnorm_silu_nhwc(x, torch.ones(96, device=DEV), torch.zeros(96, device=DEV))


@pytest.mark.parametrize("B,H,C,res,up", [(2, 64, 64, True, False), (2, 32, 128, False, False), (2, 16, 256, True, True),
                                          (3, 8, 512, True, False), (1, 32, 128, True, True)])
def test_channel_layernorm(B, H, C, res, up):
    ops = _ops()
    g = torch.Generator().manual_seed(9 + C)
    x = _rand_nhwc(B, H, H, C, g, scale=2.0)
    gain = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    r = _rand_nhwc(B, H, H, C, g) if res else None
    out = ops.channel_layernorm_nhwc(x, gain, r, upsample2x=up)
    xx = _nchw64(x)
    var = xx.var(dim=1, unbiased=False, keepdim=True)
    mean = xx.mean(dim=1, keepdim=True)
    ref = (xx - mean) * (var + 1e-5).rsqrt() * gain.double().view(1, -1, 1, 1)
    if res:
        ref = ref + _nchw64(r)
    if up:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    _check(out, ref, f"layernorm C{C} @{H}")


@pytest.mark.parametrize("B,H", [(2, 64), (3, 32), (2, 16), (5, 8)])
def test_linear_attention(B, H):
    ops = _ops()
    g = torch.Generator().manual_seed(21 + H)
    qkv = _rand_nhwc(B, H, H, 384, g, scale=1.5)
    out = ops.linear_attention_nhwc(qkv)
    q, k, v = _nchw64(qkv).chunk(3, dim=1)
    q, k, v = (t.reshape(B, 4, 32, H * H) for t in (q, k, v))
    q = q.softmax(dim=-2) * 32 ** -0.5
    k = k.softmax(dim=-1)
    v = v / (H * H)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, 128, H, H)
    _check(out, o, f"linear attention n={H * H}", rel_rms=6e-3, max_rel=3e-2)


def _linattn_block_ref(x_nchw, g1, wqkv, wo, bo, g2):
    """fp64 Residual(PreNorm(LinearAttention)) exactly as hicdiff_condition.py:64-70,99-118,199-227."""
    def ln(t, g):
        var = t.var(dim=1, unbiased=False, keepdim=True)
        mean = t.mean(dim=1, keepdim=True)
        return (t - mean) * (var + 1e-5).rsqrt() * g.reshape(1, -1, 1, 1)

    b, c, h, w = x_nchw.shape
    qkv = F.conv2d(ln(x_nchw, g1), wqkv.reshape(384, c, 1, 1)).chunk(3, dim=1)
    q, k, v = (t.reshape(b, 4, 32, h * w) for t in qkv)
    q = q.softmax(dim=-2) * 32 ** -0.5
    k = k.softmax(dim=-1)
    v = v / (h * w)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, 128, h, w)
    out = F.conv2d(out, wo.reshape(c, 128, 1, 1), bo)
    return ln(out, g2) + x_nchw


@pytest.mark.parametrize("B,H,C,kscale", [(2, 64, 64, 1.0), (3, 32, 64, 1.0), (2, 32, 128, 1.0), (5, 16, 128, 1.0),
                                          (2, 32, 64, 8.0), (300, 16, 64, 1.0),
                                          # long units (16 / 4 pixel tiles per unit) and several units per persistent CTA
                                          (160, 64, 64, 1.0), (150, 32, 128, 1.0)])
def test_linattn_block_fused(B, H, C, kscale):
    """The fused three-launch block vs the fp64 reference on the same bf16 input.  Error sources: bf16 rounding of the
    gain-folded weights, of exp(k - shift), of the softmaxed q, of the mixed to_out matrix, and of the output (the
    unfused chain rounds five intermediate tensors instead).  kscale = 8 pushes the analytic k bound past 40 so the
    softmax shift is exercised."""
    ops = _ops()
    g = torch.Generator().manual_seed(77 + B + H + C)
    x = (_rand_nhwc(B, H, H, C, g, scale=1.3).float() + 0.2).to(torch.bfloat16)
    g1 = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    g2 = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    wqkv = (torch.randn(384, C, generator=g) / math.sqrt(C))
    wqkv[128:256] *= kscale
    wqkv = wqkv.to(DEV)
    wo = (torch.randn(C, 128, generator=g) / math.sqrt(128)).to(DEV)
    bo = (0.1 * torch.randn(C, generator=g)).to(DEV)
    y, bound = ops.linattn_block_nhwc(x, g1, wqkv, wo, bo, g2)
    torch.cuda.synchronize()
    assert (bound > 40.0) == (kscale > 1.0), f"bound {bound}"
    xr = _nchw64(x)
    ref = _linattn_block_ref(xr, g1.double(), wqkv.double(), wo.double(), bo.double(), g2.double())
    _check(y, ref, f"linattn block C={C} n={H * H}", rel_rms=8e-3, max_rel=4e-2)
    # the attention branch alone (output minus the residual), where all the approximation lives
    br = _nchw64(y) - xr
    br_ref = ref - xr
    rel = ((br - br_ref).pow(2).mean().sqrt() / br_ref.pow(2).mean().sqrt()).item()
    assert rel <= 3e-2, f"attention branch rel-RMS {rel:.3e}"


def test_full_attention():
    ops = _ops()
    g = torch.Generator().manual_seed(33)
    B, H = 3, 8
    qkv = _rand_nhwc(B, H, H, 384, g, scale=1.5)
    out = ops.full_attention_nhwc(qkv)
    q, k, v = _nchw64(qkv).chunk(3, dim=1)
    q, k, v = (t.reshape(B, 4, 32, H * H) for t in (q, k, v))
    sim = torch.einsum("bhdi,bhdj->bhij", q * 32 ** -0.5, k)
    attn = sim.softmax(dim=-1)
    o = torch.einsum("bhij,bhdj->bhid", attn, v)          # b h n d
    o = o.permute(0, 1, 3, 2).reshape(B, 128, H, H)       # 'b h (x y) d -> b (h d) x y'
    _check(out, o, "full attention", rel_rms=6e-3, max_rel=3e-2)


@pytest.mark.parametrize("Cin,Cout,k", [(2, 64, 7), (1, 64, 7), (2, 256, 3), (1, 256, 3)])
def test_stem_conv(Cin, Cout, k):
    ops = _ops()
    g = torch.Generator().manual_seed(41 + Cin + k)
    B = 3
    x0 = torch.randn(B, 1, 64, 64, generator=g).to(DEV)
    x1 = torch.randn(B, 1, 64, 64, generator=g).to(DEV) if Cin == 2 else None
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).to(DEV)
    bias = (0.1 * torch.randn(Cout, generator=g)).to(DEV)
    out = ops.stem_conv(x0, x1, w, bias)
    xin = x0 if x1 is None else torch.cat((x0, x1), dim=1)
    ref = F.conv2d(xin.double(), w.double(), bias.double(), padding=k // 2)
    _check(out, ref, f"stem conv {Cin}->{Cout} k{k}")


def test_philox_normal_statistics_and_sharding_invariance():
    ops = _ops()
    z = ops.philox_normal(64, seed=1234, tile_offset=0)
    assert abs(z.mean().item()) < 0.01
    assert abs(z.std().item() - 1.0) < 0.01
    assert abs((z ** 3).mean().item()) < 0.05          # skewness
    assert abs((z ** 4).mean().item() - 3.0) < 0.1     # kurtosis
    # tile k of a batch starting at offset o equals tile (o + k) of the whole job, whatever the sharding
    z2 = ops.philox_normal(16, seed=1234, tile_offset=40)
    assert torch.equal(z2, z[40:56])
    assert not torch.equal(ops.philox_normal(4, seed=1235), z[:4])
