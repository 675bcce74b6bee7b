"""Single-operator entry points of the C ABI on torch tensors (used by the parity tests and for debugging).

Activations are NHWC bf16 `[B, H, W, C]` CUDA tensors -- the layout the kernels keep in HBM.  Nothing here falls back
to torch: every function launches the sm_100a kernel through `libhicdiff_b200.so` or raises.
"""
from __future__ import annotations

import torch

from . import _lib


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hicdiff_b200 ops need CUDA tensors (there is no CPU path)")


def _bf16c(t):
    if t is None:
        return None
    assert t.dtype == torch.bfloat16
    return t.contiguous()


def _f32c(t):
    if t is None:
        return None
    return t.detach().to(torch.float32).contiguous()


def conv2d_nhwc(x0, w, bias=None, x1=None, res=None, ksize=None, standardize=False, unshuffle=False, upsample=False,
                pad_mode=0, cta_pairs=False, dx_stack=True):
    """Implicit-GEMM conv. x0/x1: [B,H,W,C] bf16 (x1 = second operand of a channel concat); w: fp32
    [Cout, Cin, k, k] in the reference layout.  `unshuffle=True` is Downsample (pixel-unshuffle + 1x1),
    `upsample=True` is Upsample (nearest x2 + 3x3, folded into four 2x2 phase convs over the low-res input).
    `pad_mode` 1/2 opts 3x3, Cout == 64 convs into the padded-slab form (one halo'd box per chunk serves all nine taps);
    `cta_pairs` runs the GEMM on clusters of two CTAs (tcgen05 cta_group::2, M = 256 tiles, each CTA holds half of B);
    `dx_stack=False` makes the 3x3, Cout == 64 conv issue one MMA per tap instead of the default N = 192 (three dx taps) form;
    `dx_stack=1` selects that form's one-epilogue-group variant (default: two groups on alternating tiles)."""
    lib = _lib.load()
    _need_cuda(x0, w)
    x0, x1, res = _bf16c(x0), _bf16c(x1), _bf16c(res)
    w, bias = _f32c(w), _f32c(bias)
    B, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    Cout = w.shape[0]
    k = int(w.shape[-1]) if ksize is None else ksize
    Ho, Wo = (H // 2, W // 2) if unshuffle else ((2 * H, 2 * W) if upsample else (H, W))
    out = torch.empty(B, Ho, Wo, Cout, device=x0.device, dtype=torch.bfloat16)
    _lib.check(
        lib.hd_op_conv2d(_lib.ptr(x0), C0, _lib.ptr(x1), C1, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out),
                         B, Ho, Wo, Cout, k, 1 if unshuffle else (2 if upsample else 0), (1 if standardize else 0) | (int(pad_mode) << 1) | ((1 if cta_pairs else 0) << 3) | ((2 if dx_stack == 1 and dx_stack is not True else (0 if dx_stack else 1)) << 4),
                         _lib.stream_ptr()),
        "hd_op_conv2d",
    )
    return out


def conv_gn_nhwc(x0, w, bias, gamma, beta, x1=None, scale=None, shift=None, res=None, standardize=True):
    """Block = WeightStandardizedConv2d(3x3) -> GroupNorm(8) -> [x*(scale+1)+shift] -> SiLU [-> + res] in one launch."""
    lib = _lib.load()
    _need_cuda(x0, w)
    x0, x1, res = _bf16c(x0), _bf16c(x1), _bf16c(res)
    w, bias, gamma, beta, scale, shift = (_f32c(t) for t in (w, bias, gamma, beta, scale, shift))
    B, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    Cout = w.shape[0]
    out = torch.empty(B, H, W, Cout, device=x0.device, dtype=torch.bfloat16)
    _lib.check(lib.hd_op_conv_gn(_lib.ptr(x0), C0, _lib.ptr(x1), C1, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta),
                                 _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(res), _lib.ptr(out), B, H, W, Cout,
                                 1 if standardize else 0, _lib.stream_ptr()), "hd_op_conv_gn")
    return out


def groupnorm_silu_nhwc(x, gamma, beta, scale=None, shift=None, res=None):
    lib = _lib.load()
    _need_cuda(x)
    x, res = _bf16c(x), _bf16c(res)
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    gamma, beta, scale, shift = _f32c(gamma), _f32c(beta), _f32c(scale), _f32c(shift)
    _lib.check(
        lib.hd_op_groupnorm_silu(_lib.ptr(x), _lib.ptr(y), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(scale),
                                 _lib.ptr(shift), _lib.ptr(res), B, H * W, Cc, _lib.stream_ptr()),
        "hd_op_groupnorm_silu",
    )
    return y


def channel_layernorm_nhwc(x, g, res=None, upsample2x=False):
    lib = _lib.load()
    _need_cuda(x)
    x, res = _bf16c(x), _bf16c(res)
    B, H, W, Cc = x.shape
    y = torch.empty(B, H * (2 if upsample2x else 1), W * (2 if upsample2x else 1), Cc, device=x.device, dtype=torch.bfloat16)
    g = _f32c(g).reshape(-1)
    _lib.check(
        lib.hd_op_channel_layernorm(_lib.ptr(x), _lib.ptr(y), _lib.ptr(g), _lib.ptr(res), B, H, W, Cc,
                                    1 if upsample2x else 0, _lib.stream_ptr()),
        "hd_op_channel_layernorm",
    )
    return y


def linear_attention_nhwc(qkv):
    lib = _lib.load()
    _need_cuda(qkv)
    qkv = _bf16c(qkv)
    B, H, W, Cc = qkv.shape
    assert Cc == 384
    out = torch.empty(B, H, W, 128, device=qkv.device, dtype=torch.bfloat16)
    _lib.check(lib.hd_op_linear_attention(_lib.ptr(qkv), _lib.ptr(out), B, H * W, _lib.stream_ptr()), "hd_op_linear_attention")
    return out


def linattn_block_nhwc(x, g1, wqkv, wo, bo, g2):
    """Fused Residual(PreNorm(LinearAttention)): x [B,H,W,C] bf16 (C in {64,128}, H*W >= 128); g1/g2 [C] LayerNorm gains,
    wqkv [384,C(,1,1)], wo [C,128(,1,1)], bo [C].  Returns (y [B,H,W,C] bf16, analytic softmax bound)."""
    import ctypes

    lib = _lib.load()
    _need_cuda(x)
    x = _bf16c(x)
    B, H, W, C = x.shape
    g1, wqkv, wo, bo, g2 = (_f32c(t.reshape(t.shape[0], -1) if t.dim() > 1 else t) for t in (g1, wqkv, wo, bo, g2))
    y = torch.empty_like(x)
    bound = ctypes.c_float(0.0)
    _lib.check(lib.hd_op_linattn_block(_lib.ptr(x), _lib.ptr(g1.reshape(-1)), _lib.ptr(wqkv), _lib.ptr(wo), _lib.ptr(bo),
                                       _lib.ptr(g2.reshape(-1)), _lib.ptr(y), B, H * W, C, ctypes.byref(bound),
                                       _lib.stream_ptr()), "hd_op_linattn_block")
    return y, bound.value


def full_attention_nhwc(qkv):
    lib = _lib.load()
    _need_cuda(qkv)
    qkv = _bf16c(qkv)
    B, H, W, Cc = qkv.shape
    assert Cc == 384
    out = torch.empty(B, H, W, 128, device=qkv.device, dtype=torch.bfloat16)
    _lib.check(lib.hd_op_full_attention(_lib.ptr(qkv), _lib.ptr(out), B, H * W, _lib.stream_ptr()), "hd_op_full_attention")
    return out


def stem_conv(x0, x1, w, bias):
    """x0/x1: fp32 [B,1,64,64] planes (x1 may be None); w fp32 [Cout, Cin, k, k] -> NHWC bf16 [B,64,64,Cout]."""
    lib = _lib.load()
    _need_cuda(x0, w)
    x0, x1, w, bias = _f32c(x0), _f32c(x1), _f32c(w), _f32c(bias)
    B = x0.shape[0]
    Cout, Cin, k, _ = w.shape
    y = torch.empty(B, 64, 64, Cout, device=x0.device, dtype=torch.bfloat16)
    _lib.check(lib.hd_op_stem_conv(_lib.ptr(x0), _lib.ptr(x1), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), B, Cout, Cin, k,
                                   _lib.stream_ptr()), "hd_op_stem_conv")
    return y


def philox_normal(n_tiles, seed, tile_offset=0, device="cuda"):
    lib = _lib.load()
    out = torch.empty(n_tiles, 1, 64, 64, device=device, dtype=torch.float32)
    _lib.check(lib.hd_op_philox_normal(_lib.ptr(out), out.numel(), seed, tile_offset, _lib.stream_ptr()), "hd_op_philox_normal")
    return out


def tile_count(n, piece=64, band_blocks=4):
    return int(_lib.load().hd_tile_count(n, piece, band_blocks))


def tile_extract(mat, piece=64, band_blocks=4):
    """splitPieces on the GPU: mat fp32 [n, n] -> tiles fp32 [k, 1, piece, piece]."""
    lib = _lib.load()
    _need_cuda(mat)
    mat = _f32c(mat)
    n = mat.shape[0]
    assert mat.shape == (n, n)
    k = tile_count(n, piece, band_blocks)
    tiles = torch.empty(k, 1, piece, piece, device=mat.device, dtype=torch.float32)
    _lib.check(lib.hd_tile_extract(_lib.ptr(mat), n, _lib.ptr(tiles), piece, band_blocks, _lib.stream_ptr()), "hd_tile_extract")
    return tiles


def tile_scatter(tiles, n, piece=64, band_blocks=4):
    """Inverse of tile_extract: tiles fp32 [k,1,piece,piece] -> symmetric band matrix fp32 [n, n]."""
    lib = _lib.load()
    _need_cuda(tiles)
    tiles = _f32c(tiles)
    k = tile_count(n, piece, band_blocks)
    if tiles.shape[0] != k:
        raise ValueError(f"expected {k} tiles for n={n}, got {tiles.shape[0]}")
    mat = torch.zeros(n, n, device=tiles.device, dtype=torch.float32)
    _lib.check(lib.hd_tile_scatter(_lib.ptr(tiles), _lib.ptr(mat), n, piece, band_blocks, _lib.stream_ptr()), "hd_tile_scatter")
    return mat


def conv3x3_wgrad_nhwc(x, dy):
    """d/dW of conv2d(x, W, padding=1) for a 256 -> 256 3x3 conv over 64x64 tiles: x, dy [B,64,64,256] bf16 -> fp32
    [256,256,3,3] (tcgen05 wgrad kernel over MN-major NHWC tiles + fixed-order split reduction)."""
    lib = _lib.load()
    _need_cuda(x, dy)
    x, dy = _bf16c(x), _bf16c(dy)
    if x.shape != dy.shape or tuple(x.shape[1:]) != (64, 64, 256):
        raise ValueError(f"expected two [B,64,64,256] tensors, got {tuple(x.shape)} / {tuple(dy.shape)}")
    dw = torch.empty(256, 256, 3, 3, device=x.device, dtype=torch.float32)
    _lib.check(lib.hd_op_conv3x3_wgrad(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), x.shape[0], _lib.stream_ptr()), "hd_op_conv3x3_wgrad")
    return dw


def conv3x3_dgrad_nhwc(dy, w):
    """d/dx of conv2d(x, W, padding=1): dy [B,64,64,256] bf16, W fp32 [256,256,3,3] -> dx [B,64,64,256] bf16 (the forward
    implicit-GEMM kernel over the flipped, transposed filter)."""
    lib = _lib.load()
    _need_cuda(dy, w)
    dy, w = _bf16c(dy), _f32c(w)
    if tuple(dy.shape[1:]) != (64, 64, 256) or tuple(w.shape) != (256, 256, 3, 3):
        raise ValueError("expected dy [B,64,64,256] and w [256,256,3,3]")
    dx = torch.empty_like(dy)
    _lib.check(lib.hd_op_conv3x3_dgrad(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), dy.shape[0], _lib.stream_ptr()), "hd_op_conv3x3_dgrad")
    return dx


def conv_wgrad_nhwc(x, dy, ksize, dw=None, cin_total=None, ci0=0):
    """General conv weight gradient: x [B,H,W,Cin], dy [B,H,W,Cout] bf16 -> dw fp32 [Cout, cin_total, k, k] (columns
    [ci0, ci0 + Cin) are written; pass `dw` to fill the slices of a channel concat one operand at a time)."""
    lib = _lib.load()
    _need_cuda(x, dy)
    x, dy = _bf16c(x), _bf16c(dy)
    B, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    cin_total = Cin if cin_total is None else cin_total
    if dw is None:
        dw = torch.zeros(Cout, cin_total, ksize, ksize, device=x.device, dtype=torch.float32)
    _lib.check(lib.hd_op_conv_wgrad(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), B, H, W, Cin, Cout, ksize, cin_total, ci0,
                                    _lib.stream_ptr()), "hd_op_conv_wgrad")
    return dw


def groupnorm_silu_bwd_nhwc(y, ds, gamma, beta, scale=None, shift=None, want_conv_bias=False):
    """Backward of SiLU(GroupNorm_8(y) * (scale + 1) + shift): y, ds [B,H,W,C] bf16 -> (dy bf16, dgamma, dbeta, dscale, dshift)."""
    lib = _lib.load()
    _need_cuda(y, ds)
    y, ds = _bf16c(y), _bf16c(ds)
    B, H, W, C = y.shape
    gamma, beta, scale, shift = _f32c(gamma), _f32c(beta), _f32c(scale), _f32c(shift)
    dy = torch.empty_like(y)
    dgamma = torch.empty(C, device=y.device, dtype=torch.float32)
    dbeta = torch.empty_like(dgamma)
    dscale = torch.empty(B, C, device=y.device, dtype=torch.float32) if scale is not None else None
    dshift = torch.empty_like(dscale) if scale is not None else None
    dcb = torch.empty_like(dgamma) if want_conv_bias else None
    _lib.check(lib.hd_op_groupnorm_silu_bwd(_lib.ptr(y), _lib.ptr(ds), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(scale), _lib.ptr(shift),
                                            _lib.ptr(dy), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dscale), _lib.ptr(dshift),
                                            _lib.ptr(dcb), B, H * W, C, _lib.stream_ptr()), "hd_op_groupnorm_silu_bwd")
    if want_conv_bias:
        return dy, dgamma, dbeta, dscale, dshift, dcb
    return dy, dgamma, dbeta, dscale, dshift


def channel_layernorm_bwd_nhwc(x, dz, g):
    """Backward of the channel LayerNorm: x, dz [..., C] bf16, g [C] -> (dx bf16, dg fp32 [C])."""
    lib = _lib.load()
    _need_cuda(x, dz)
    x, dz, g = _bf16c(x), _bf16c(dz), _f32c(g).reshape(-1)
    C = x.shape[-1]
    dx = torch.empty_like(x)
    dg = torch.empty(C, device=x.device, dtype=torch.float32)
    _lib.check(lib.hd_op_channel_layernorm_bwd(_lib.ptr(x), _lib.ptr(dz), _lib.ptr(g), _lib.ptr(dx), _lib.ptr(dg), x.numel() // C, C,
                                               _lib.stream_ptr()), "hd_op_channel_layernorm_bwd")
    return dx, dg


def weight_standardize_bwd(w, dwt):
    """Gradient w.r.t. the raw weight from the gradient w.r.t. WeightStandardizedConv2d's standardised weight."""
    lib = _lib.load()
    _need_cuda(w, dwt)
    w, dwt = _f32c(w), _f32c(dwt)
    dw = torch.empty_like(w)
    _lib.check(lib.hd_op_weight_standardize_bwd(_lib.ptr(w), _lib.ptr(dwt), _lib.ptr(dw), w.shape[0], w[0].numel(), _lib.stream_ptr()),
               "hd_op_weight_standardize_bwd")
    return dw


def attention_bwd(qkv, dout, linear=True):
    """Backward of LinearAttention's core (linear=True) or of the 8x8 softmax attention: qkv [B,n,384], dout [B,n,128] bf16."""
    lib = _lib.load()
    _need_cuda(qkv, dout)
    qkv, dout = _bf16c(qkv), _bf16c(dout)
    B, n = qkv.shape[0], qkv.shape[1]
    dqkv = torch.empty_like(qkv)
    _lib.check(lib.hd_op_attention_bwd(_lib.ptr(qkv), _lib.ptr(dout), _lib.ptr(dqkv), B, n, 1 if linear else 0, _lib.stream_ptr()),
               "hd_op_attention_bwd")
    return dqkv
