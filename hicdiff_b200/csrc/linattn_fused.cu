// Fused Residual(PreNorm(LinearAttention)) block of the HiCDiff UNet for C <= 128 channels and n >= 128 pixels.
//
// reference: Residual :64-70, PreNorm :110-118, LayerNorm :99-108, LinearAttention :199-227 of
//            /root/reference/src/hicdiff_condition.py  (heads = 4, dim_head = 32)
//     y = x + LN_g2( Wo * lin_attn( Wqkv * LN_g1(x) ) + bo )
//     q = softmax_d(q) * 32^-0.5 ; k = softmax_n(k) ; v = v / n ; ctx[d,e] = sum_n k[d,n] v[e,n] ; out[e,n] = sum_d ctx[d,e] q[d,n]
//
// The unfused chain (LayerNorm, to_qkv GEMM, two attention kernels, to_out GEMM, LayerNorm + residual) moves
// ~27 B per input byte through HBM because the [B, n, 384] qkv tensor is 6x wider than x; here qkv never leaves the SM:
//
//   linattn_kv_kernel   (one CTA per (image, pixel range); linattn_kv2_kernel is its persistent, double-buffered form): for every
//       64-pixel tile, TMA-load x, LayerNorm it IN PLACE in shared memory (gain folded into the weights, statistics by shuffles),
//       tcgen05 GEMMs K^T[d, px] = Wk' z^T and V^T[e, px] = Wv' z^T, p = exp(k - shift_d) written bf16 K-major to shared
//       memory, then ctx[d, e] += P V^T as a third tcgen05 GEMM accumulating in TMEM over
//       the CTA's tiles.  Writing the GEMMs transposed makes every operand K-major and gives each epilogue thread one
//       (d or e) row, so the softmax denominators are in-thread sums.  softmax_n is shift-invariant: instead of the
//       data-dependent max it uses the analytic bound |k_d| <= sqrt(C) * ||Wk'_d||_2 (shift_d = max(0, bound_d - 40)),
//       which cannot overflow and only underflows for bound_d > ~120; the plan falls back to the unfused kernels then.
//   linattn_mix_kernel  (one CTA per image): merges the partial contexts, normalises (1/S[d], 1/n, 32^-0.5) and folds
//       to_out into them: Mb[co, hd] = sum_e Wo[co, h*32+e] ctx_h[d, e], bf16 [C, 128].
//   linattn_out_kernel  (persistent, 2 CTAs / SM at C = 64): per 128-pixel tile Q[px, hd] = x Wq'^T (tcgen05) -> per-row
//       LayerNorm fix-up + softmax over each head's 32 columns in registers -> bf16 A operand -> Y[px, co] = Q Mb^T
//       (tcgen05) -> + bias, LayerNorm g2, + x (still in shared memory) -> TMA store.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

constexpr int HD = 128;                 // heads * dim_head
constexpr int TILE = 128;               // pixels per tile
constexpr uint32_t SPAN_BYTES = TILE * 128;   // one 64-channel K span of a 128-row operand: 16 KiB

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// byte offset of 16-byte chunk c of row r inside a K-major, 128B-swizzled tile with 128-byte rows
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(c ^ (r & 7)) << 4); }

// ------------------------------------------------------------------------------------------------ weight prep
// out[row, c] = bf16(w[row, c] * g1[c]); rowsum[row] = sum_c out[row, c]; for the k rows additionally the softmax shift.
__global__ void __launch_bounds__(128)
linattn_prep_kernel(const float* __restrict__ wqkv, const float* __restrict__ g1, bf16* __restrict__ out,
                    float* __restrict__ rowsum, float* __restrict__ kshift, int* __restrict__ max_bound_bits, int C) {
    __shared__ float s_a[4], s_b[4];
    const int row = blockIdx.x;
    const int c = threadIdx.x;
    float v = 0.f;
    if (c < C) {
        const bf16 q = __float2bfloat16(wqkv[static_cast<size_t>(row) * C + c] * g1[c]);
        out[static_cast<size_t>(row) * C + c] = q;
        v = __bfloat162float(q);
    }
    float s = v, ss = v * v;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if ((c & 31) == 0) { s_a[c >> 5] = s; s_b[c >> 5] = ss; }
    __syncthreads();
    if (c == 0) {
        const float ts = s_a[0] + s_a[1] + s_a[2] + s_a[3];
        const float tss = s_b[0] + s_b[1] + s_b[2] + s_b[3];
        rowsum[row] = ts;
        if (row >= HD && row < 2 * HD) {
            const float bound = sqrtf(static_cast<float>(C) * tss);      // |k_d| <= ||z||_2 ||Wk'_d||_2, ||z||_2 <= sqrt(C)
            kshift[row - HD] = fmaxf(bound - 40.0f, 0.f);
            atomicMax(max_bound_bits, __float_as_int(bound));           // non-negative floats order like ints
        }
    }
}

// ------------------------------------------------------------------------------------------------ kernel 1: context
struct KvArgs {
    int n, tiles_per_unit, parts;
    const float* kshift;     // [128]
    float* ctx_part;         // [B * parts][128 (h*32+d)][32 (e)]
    float* s_part;           // [B * parts][128]
    float eps;
};

constexpr int KV_PX = 64;                // pixels per kv tile: K^T / V^T accumulators are 64 columns each

// LayerNorm of the 64 pixels of a kv tile IN PLACE in shared memory (the gain is folded into the weights), by NT epilogue threads:
// NT / 64 consecutive lanes share a pixel, each keeps its share of the row's channels in registers (two-pass statistics), the partial
// sums meet in xor-shuffles (a + b == b + a: every lane of a pixel ends with the same bits), and the normalised row goes back to the
// same swizzled chunks as bf16.  The MMAs then produce k and v themselves -- the epilogue needs no per-pixel constants (the earlier
// form corrected the raw-x products per element: W LN(x) = rstd * (W'x - mean * rowsum(W')), ~60 % of its instructions and, through
// the shared-memory loads of the per-pixel constants, most of its exposed latency: profiles/r02_notes.md 10).
template <int C, int NT, uint32_t SPAN_STRIDE>
__device__ __forceinline__ void tile_normalize(uint8_t* sx, int te, float eps) {
    constexpr int TPP = NT / KV_PX;              // threads per pixel: 4 or 8
    constexpr int NCH = (C / 8) / TPP;           // 16-byte chunks per thread
    static_assert(NCH >= 1, "tile_normalize: more threads than chunks");
    const int p = te / TPP, sub = te % TPP;
    ptx::f32x2 xh[NCH * 4];
    uint4* src[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = sub + i * TPP;             // chunk of the row: span c >> 3, chunk c & 7 inside it
        src[i] = reinterpret_cast<uint4*>(sx + (c >> 3) * SPAN_STRIDE + sw_off(p, c & 7));
        const uint4 u = *src[i];
        xh[i * 4] = ptx::bf16x2_to_f32x2(u.x); xh[i * 4 + 1] = ptx::bf16x2_to_f32x2(u.y);
        xh[i * 4 + 2] = ptx::bf16x2_to_f32x2(u.z); xh[i * 4 + 3] = ptx::bf16x2_to_f32x2(u.w);
    }
    ptx::f32x2 s0 = xh[0], s1 = xh[1];
#pragma unroll
    for (int i = 2; i < NCH * 4; i += 2) { s0 = ptx::add2(s0, xh[i]); s1 = ptx::add2(s1, xh[i + 1]); }
    s0 = ptx::add2(s0, s1);
    float sum = ptx::lo(s0) + ptx::hi(s0);
#pragma unroll
    for (int off = 1; off < TPP; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const ptx::f32x2 nm = ptx::dup2(-(sum * (1.0f / C)));
    ptx::f32x2 q0 = ptx::mk2(0.f, 0.f), q1 = q0;
#pragma unroll
    for (int i = 0; i < NCH * 4; i += 2) {
        xh[i] = ptx::add2(xh[i], nm);
        xh[i + 1] = ptx::add2(xh[i + 1], nm);
        q0 = ptx::fma2(xh[i], xh[i], q0);
        q1 = ptx::fma2(xh[i + 1], xh[i + 1], q1);
    }
    q0 = ptx::add2(q0, q1);
    float ss = ptx::lo(q0) + ptx::hi(q0);
#pragma unroll
    for (int off = 1; off < TPP; off <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const ptx::f32x2 RS = ptx::dup2(rsqrtf(ss * (1.0f / C) + eps));
#pragma unroll
    for (int i = 0; i < NCH; ++i)
        *src[i] = make_uint4(ptx::pack_bf16x2(ptx::mul2(xh[i * 4], RS)), ptx::pack_bf16x2(ptx::mul2(xh[i * 4 + 1], RS)),
                             ptx::pack_bf16x2(ptx::mul2(xh[i * 4 + 2], RS)), ptx::pack_bf16x2(ptx::mul2(xh[i * 4 + 3], RS)));
}

// Eight pixels (one 16-byte shared-memory chunk) of the kv epilogue, packed fp32 (ptx.cuh): v = accumulator registers OFF .. OFF + 7.
//   K path: p = exp(k - shift) = exp2(k * log2e - shift * log2e); ssum2 accumulates the unrounded p (the bf16 rounding of the MMA
//           operand is unbiased: over n >= 1024 pixels the two sums agree to ~1e-5)
//   V path: the accumulator is v
template <int OFF, int N>
__device__ __forceinline__ uint4 k_chunk8(const uint32_t (&v)[N], ptx::f32x2 L2E, ptx::f32x2 NSHIFT, ptx::f32x2& ssum2) {
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const ptx::f32x2 arg = ptx::fma2(ptx::mk2(v[OFF + 2 * e], v[OFF + 2 * e + 1]), L2E, NSHIFT);
        const ptx::f32x2 pp = ptx::mk2(ptx::ex2(ptx::lo(arg)), ptx::ex2(ptx::hi(arg)));
        ssum2 = ptx::add2(ssum2, pp);
        o[e] = ptx::pack_bf16x2(pp);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
template <int OFF, int N>
__device__ __forceinline__ uint4 v_chunk8(const uint32_t (&v)[N]) {
    return make_uint4(ptx::pack_bf16x2(__uint_as_float(v[OFF]), __uint_as_float(v[OFF + 1])),
                      ptx::pack_bf16x2(__uint_as_float(v[OFF + 2]), __uint_as_float(v[OFF + 3])),
                      ptx::pack_bf16x2(__uint_as_float(v[OFF + 4]), __uint_as_float(v[OFF + 5])),
                      ptx::pack_bf16x2(__uint_as_float(v[OFF + 6]), __uint_as_float(v[OFF + 7])));
}

// linattn_kv_kernel: one CTA per unit (image, pixel range), two CTAs per SM at C = 64 (one CTA's MMAs run under the other's epilogue).
// K^T / V^T accumulators 64 columns each + ctx 128 = 256 TMEM columns.  Per tile t the epilogue warps first normalise tile t + 1 in
// place (under the MMAs of tile t), then drain tile t.
template <int C>
struct KvCfg {
    static constexpr int SPANS = C / 64;
    static constexpr uint32_t W_BYTES = SPANS * SPAN_BYTES;          // one of Wk' / Wv' [128 rows][C]
    static constexpr uint32_t XSPAN_BYTES = KV_PX * 128;             // one 64-channel span of a 64-pixel x tile
    static constexpr uint32_t X_BYTES = SPANS * XSPAN_BYTES;
    static constexpr int X_STAGES = 3;
    static constexpr uint32_t PV_BYTES = SPAN_BYTES;                 // [128 rows][64 px]
    static constexpr uint32_t SMALL_BYTES = 2 * 512 + 256;           // s_S [2][128], barriers
    static constexpr int SMEM_BYTES = 2 * W_BYTES + X_STAGES * X_BYTES + 2 * PV_BYTES + SMALL_BYTES + 1024;
    static constexpr int CTAS_PER_SM = C == 64 ? 2 : 1;
};
constexpr int KV_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue: (TMEM lane quarter, 32-pixel half)

template <int C>
__global__ void __launch_bounds__(KV_THREADS, KvCfg<C>::CTAS_PER_SM)
linattn_kv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const KvArgs a) {
    using Cf = KvCfg<C>;
    constexpr int SPANS = Cf::SPANS;
    constexpr int XS = Cf::X_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    uint8_t* sWk = smem;
    uint8_t* sWv = sWk + Cf::W_BYTES;
    uint8_t* sX = sWv + Cf::W_BYTES;                 // X_STAGES stages
    uint8_t* sP = sX + XS * Cf::X_BYTES;
    uint8_t* sV = sP + Cf::PV_BYTES;
    float* s_S = reinterpret_cast<float*>(sV + Cf::PV_BYTES);        // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_S + 256);
    uint64_t* w_bar = bars;
    uint64_t* x_full = bars + 1;                     // [3] TMA -> normaliser
    uint64_t* xn_full = bars + 4;                    // [3] normaliser -> MMA
    uint64_t* x_empty = bars + 7;                    // [3] MMA done with the stage -> TMA
    uint64_t* d_full = bars + 10;
    uint64_t* pv_ready = bars + 11;
    uint64_t* ctx_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        ptx::mbar_init(w_bar, 1);
        for (int i = 0; i < XS; ++i) { ptx::mbar_init(&x_full[i], 1); ptx::mbar_init(&xn_full[i], 8); ptx::mbar_init(&x_empty[i], 1); }
        ptx::mbar_init(d_full, 1);
        ptx::mbar_init(pv_ready, 8);
        ptx::mbar_init(ctx_full, 1);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    constexpr uint32_t COL_K = 0, COL_V = 64, COL_CTX = 128;

    const int unit = blockIdx.x;
    const int b = unit / a.parts;
    const int part = unit - b * a.parts;
    const int T = a.tiles_per_unit;                  // 64-pixel tiles of this unit
    const int m0 = b * a.n + part * T * KV_PX;

    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(w_bar, 2 * Cf::W_BYTES);
            for (int sp = 0; sp < SPANS; ++sp) {
                ptx::tma_load_2d(sWk + sp * SPAN_BYTES, &tmW, w_bar, sp * 64, HD);
                ptx::tma_load_2d(sWv + sp * SPAN_BYTES, &tmW, w_bar, sp * 64, 2 * HD);
            }
        }
        __syncwarp();
        for (int t = 0; t < T; ++t) {
            const int st = t % XS;
            ptx::mbar_wait(&x_empty[st], ((t / XS) & 1u) ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&x_full[st], Cf::X_BYTES);
                for (int sp = 0; sp < SPANS; ++sp)
                    ptx::tma_load_2d(sX + st * Cf::X_BYTES + sp * Cf::XSPAN_BYTES, &tmX, &x_full[st], sp * 64, m0 + t * KV_PX);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_kv = ptx::make_idesc_bf16(128, KV_PX);     // K^T / V^T: M = 128 (d | e), N = 64 pixels
        constexpr uint32_t idesc_ctx = ptx::make_idesc_bf16(128, 128);      // ctx: M = 128 d, N = 128 e, K = 64 pixels
        const uint64_t dWk = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWk));
        const uint64_t dWv = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWv));
        const uint64_t dX = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sX));
        const uint64_t dP = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sP));
        const uint64_t dV = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sV));
        ptx::mbar_wait(w_bar, 0);
        for (int t = 0; t <= T; ++t) {
            if (t < T) ptx::mbar_wait(&xn_full[t % XS], (t / XS) & 1u);   // tile t normalised in place
            if (t > 0) ptx::mbar_wait(pv_ready, (t - 1) & 1u);     // epilogue t-1: D_K / D_V drained, P / V written
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                if (t > 0) {                                        // ctx[d, e] += P V^T over the 64 pixels of tile t-1
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16(tmem_base + COL_CTX, dP + 2u * k, dV + 2u * k, idesc_ctx, (t > 1 || k != 0) ? 1u : 0u);
                }
                if (t < T) {                                        // K^T = Wk' z^T, V^T = Wv' z^T of tile t (z = normalised x)
                    const uint64_t xoff = static_cast<uint64_t>(((t % XS) * Cf::X_BYTES) >> 4);
#pragma unroll
                    for (int sp = 0; sp < SPANS; ++sp)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t offw = static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                            const uint64_t offx = static_cast<uint64_t>((sp * Cf::XSPAN_BYTES) >> 4) + 2u * k;
                            ptx::umma_bf16(tmem_base + COL_K, dWk + offw, dX + xoff + offx, idesc_kv, (sp | k) != 0 ? 1u : 0u);
                        }
#pragma unroll
                    for (int sp = 0; sp < SPANS; ++sp)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t offw = static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                            const uint64_t offx = static_cast<uint64_t>((sp * Cf::XSPAN_BYTES) >> 4) + 2u * k;
                            ptx::umma_bf16(tmem_base + COL_V, dWv + offw, dX + xoff + offx, idesc_kv, (sp | k) != 0 ? 1u : 0u);
                        }
                    ptx::umma_commit(d_full);
                } else {
                    ptx::umma_commit(ctx_full);
                }
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue: 8 warps, thread = (row, 32-pixel half)
        const int te = (warp - 2) * 32 + lane;          // 0..255
        const int q = warp & 3;                         // TMEM lane quarter
        const int hcol = (warp - 2) >> 2;               // which 32 pixels of the tile
        const int r = q * 32 + lane;                    // d (K^T, ctx) / e (V^T) row
        constexpr float LOG2E = 1.4426950408889634f;
        const ptx::f32x2 L2E = ptx::dup2(LOG2E), NSHIFT = ptx::dup2(-(a.kshift[r] * LOG2E));
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int px0 = hcol * 32;
        ptx::f32x2 ssum2 = ptx::mk2(0.f, 0.f);
        auto normalise = [&](int t) {                   // tile t: wait for the TMA, LayerNorm in place, hand the stage to the MMA warp
            const int st = t % XS;
            ptx::mbar_wait(&x_full[st], (t / XS) & 1u);
            tile_normalize<C, 256, Cf::XSPAN_BYTES>(sX + st * Cf::X_BYTES, te, a.eps);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&xn_full[st]);
        };
        normalise(0);
        for (int t = 0; t < T; ++t) {
            if (t + 1 < T) normalise(t + 1);                      // under the MMAs of tile t
            ptx::mbar_wait(d_full, t & 1u);                       // K^T / V^T of tile t ready (and P / V of tile t-1 consumed)
            ptx::tc_fence_after();
            if (te == 0) ptx::mbar_arrive(&x_empty[t % XS]);      // the MMAs have read the stage
            uint32_t v[32];
            ptx::tmem_ld32(tlane + COL_K + px0, v);
            ptx::tmem_ld_wait();
            *reinterpret_cast<uint4*>(sP + sw_off(r, hcol * 4 + 0)) = k_chunk8<0>(v, L2E, NSHIFT, ssum2);
            *reinterpret_cast<uint4*>(sP + sw_off(r, hcol * 4 + 1)) = k_chunk8<8>(v, L2E, NSHIFT, ssum2);
            *reinterpret_cast<uint4*>(sP + sw_off(r, hcol * 4 + 2)) = k_chunk8<16>(v, L2E, NSHIFT, ssum2);
            *reinterpret_cast<uint4*>(sP + sw_off(r, hcol * 4 + 3)) = k_chunk8<24>(v, L2E, NSHIFT, ssum2);
            ptx::tmem_ld32(tlane + COL_V + px0, v);
            ptx::tmem_ld_wait();
            *reinterpret_cast<uint4*>(sV + sw_off(r, hcol * 4 + 0)) = v_chunk8<0>(v);
            *reinterpret_cast<uint4*>(sV + sw_off(r, hcol * 4 + 1)) = v_chunk8<8>(v);
            *reinterpret_cast<uint4*>(sV + sw_off(r, hcol * 4 + 2)) = v_chunk8<16>(v);
            *reinterpret_cast<uint4*>(sV + sw_off(r, hcol * 4 + 3)) = v_chunk8<24>(v);
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(pv_ready);
        }
        s_S[hcol * 128 + r] = ptx::lo(ssum2) + ptx::hi(ssum2);
        named_bar_sync(1, 256);
        if (hcol == 0) {
            a.s_part[static_cast<size_t>(unit) * HD + r] = s_S[r] + s_S[128 + r];
            ptx::mbar_wait(ctx_full, 0);
            ptx::tc_fence_after();
            uint32_t v[32];
            ptx::tmem_ld32(tlane + COL_CTX + q * 32, v);     // row d = q*32 + lane belongs to head q: columns e of head q
            ptx::tmem_ld_wait();
            float4* dst = reinterpret_cast<float4*>(a.ctx_part + (static_cast<size_t>(unit) * HD + r) * 32);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                dst[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                          __uint_as_float(v[j + 3]));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ kernel 1, pipelined form
// linattn_kv2_kernel: the same math, one persistent CTA per SM walking its units, as three decoupled pipelines that only meet in
// mbarriers: eight LayerNorm warps normalise the x stages in place (two groups on alternating tiles at C = 64), the MMA warp runs one
// tile ahead into double-buffered K^T / V^T accumulators, sixteen epilogue warps (TMEM lane quarter x 16-pixel column quarter) drain
// tile g into double-buffered P / V tiles while the MMAs of tile g + 1 run; weights and TMEM are set up once per CTA.  (In the
// one-CTA-per-unit form every tile is a serial chain normalise -> MMA -> drain whose latencies only the second CTA of the SM hides.)
//   TMEM (512 columns): [K0 | V0 | K1 | V1] 4 x 64, ctx 128 at column 256.
//   barriers: x_full / xn_full / x_empty[3] (TMA -> normaliser -> MMA -> TMA), kv_full / kv_empty[2] (accumulators), pv_full /
//   pv_empty[2] (P / V tiles), ctx_full / ctx_empty (context accumulator of one unit).  g counts tiles across units.
template <int C>
struct Kv2Cfg {
    static constexpr int SPANS = C / 64;
    static constexpr uint32_t W_BYTES = SPANS * SPAN_BYTES;
    static constexpr uint32_t XSPAN_BYTES = KV_PX * 128;
    static constexpr uint32_t X_BYTES = SPANS * XSPAN_BYTES;
    static constexpr int X_STAGES = 4;
    static constexpr int NORM_GROUPS = C == 64 ? 2 : 1;              // normaliser warp groups, alternating tiles (C = 128: one group,
                                                                     // four threads per pixel, so that a row share fits the registers)
    static constexpr uint32_t PV_BYTES = SPAN_BYTES;                 // [128 rows][64 px]
    static constexpr uint32_t SMALL_BYTES = 8 * 512 + 256;           // s_S [2][4][128] (by unit parity), barriers
    static constexpr int SMEM_BYTES = 2 * W_BYTES + X_STAGES * X_BYTES + 4 * PV_BYTES + SMALL_BYTES + 1024;
};
constexpr int KV2_EPI_WARPS = 16;
constexpr int KV2_NORM_WARPS = 8;
constexpr int KV2_CTX_WARP = 2 + KV2_EPI_WARPS + KV2_NORM_WARPS;            // warp 26: issues the context MMAs
constexpr int KV2_THREADS = (KV2_CTX_WARP + 1) * 32;      // warp 0 TMA, warp 1 K^T / V^T MMAs, 2..17 epilogue, 18..25 LayerNorm, 26 context MMAs

template <int C>
__global__ void __launch_bounds__(KV2_THREADS, 1)
linattn_kv2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const KvArgs a, const int units) {
    using Cf = Kv2Cfg<C>;
    constexpr int SPANS = Cf::SPANS;
    constexpr int XS = Cf::X_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sWk = smem;
    uint8_t* sWv = sWk + Cf::W_BYTES;
    uint8_t* sX = sWv + Cf::W_BYTES;                 // X_STAGES stages
    uint8_t* sP = sX + XS * Cf::X_BYTES;             // [2]
    uint8_t* sV = sP + 2 * Cf::PV_BYTES;             // [2]
    float* s_S = reinterpret_cast<float*>(sV + 2 * Cf::PV_BYTES);    // [2][4][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_S + 1024);
    uint64_t* w_bar = bars;
    uint64_t* x_full = bars + 1;                     // [4]
    uint64_t* xn_full = bars + 5;                    // [4]
    uint64_t* x_empty = bars + 9;                    // [4]
    uint64_t* kv_full = bars + 13;                   // [2]
    uint64_t* kv_empty = bars + 15;                  // [2]
    uint64_t* pv_full = bars + 17;                   // [2]
    uint64_t* pv_empty = bars + 19;                  // [2]
    uint64_t* ctx_full = bars + 21;
    uint64_t* ctx_empty = bars + 22;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        ptx::mbar_init(w_bar, 1);
        for (int i = 0; i < XS; ++i) {
            ptx::mbar_init(&x_full[i], 1);
            ptx::mbar_init(&xn_full[i], KV2_NORM_WARPS / Cf::NORM_GROUPS);
            ptx::mbar_init(&x_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], KV2_EPI_WARPS / 2);     // the eight warps of epilogue group i
            ptx::mbar_init(&pv_full[i], KV2_EPI_WARPS / 2);
            ptx::mbar_init(&pv_empty[i], 1);
        }
        ptx::mbar_init(ctx_full, 1);
        ptx::mbar_init(ctx_empty, 4);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();
    constexpr uint32_t COL_CTX = 256;

    const int T = a.tiles_per_unit;                  // 64-pixel tiles per unit
    const int my_units = units > static_cast<int>(blockIdx.x) ? (units - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int G = my_units * T;                      // tiles of this CTA: g = (unit index of this CTA) * T + t
    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(w_bar, 2 * Cf::W_BYTES);
            for (int sp = 0; sp < SPANS; ++sp) {
                ptx::tma_load_2d(sWk + sp * SPAN_BYTES, &tmW, w_bar, sp * 64, HD);
                ptx::tma_load_2d(sWv + sp * SPAN_BYTES, &tmW, w_bar, sp * 64, 2 * HD);
            }
        }
        __syncwarp();
        ptx::grid_dep_wait();
        int g = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int b = unit / a.parts;
            const int part = unit - b * a.parts;
            const int m0 = b * a.n + part * T * KV_PX;
            for (int t = 0; t < T; ++t, ++g) {
                const int st = g % XS;
                ptx::mbar_wait(&x_empty[st], ((g / XS) & 1u) ^ 1u);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&x_full[st], Cf::X_BYTES);
                    for (int sp = 0; sp < SPANS; ++sp)
                        ptx::tma_load_2d(sX + st * Cf::X_BYTES + sp * Cf::XSPAN_BYTES, &tmX, &x_full[st], sp * 64, m0 + t * KV_PX);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_kv = ptx::make_idesc_bf16(128, KV_PX);
        const uint64_t dWk = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWk));
        const uint64_t dWv = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWv));
        const uint64_t dX = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sX));
        ptx::mbar_wait(w_bar, 0);
        for (int g = 0; g < G; ++g) {                // K^T / V^T of tile g into accumulator pair g & 1, as far ahead as the buffers allow
            const int st = g % XS, bb = g & 1;
            ptx::mbar_wait(&xn_full[st], (g / XS) & 1u);         // tile g normalised in place
            ptx::mbar_wait(&kv_empty[bb], ((g >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t xoff = static_cast<uint64_t>((st * Cf::X_BYTES) >> 4);
                const uint32_t colk = tmem_base + bb * 128, colv = colk + 64;
#pragma unroll
                for (int sp = 0; sp < SPANS; ++sp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offw = static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                        const uint64_t offx = static_cast<uint64_t>((sp * Cf::XSPAN_BYTES) >> 4) + 2u * k;
                        ptx::umma_bf16(colk, dWk + offw, dX + xoff + offx, idesc_kv, (sp | k) != 0 ? 1u : 0u);
                    }
#pragma unroll
                for (int sp = 0; sp < SPANS; ++sp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offw = static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                        const uint64_t offx = static_cast<uint64_t>((sp * Cf::XSPAN_BYTES) >> 4) + 2u * k;
                        ptx::umma_bf16(colv, dWv + offw, dX + xoff + offx, idesc_kv, (sp | k) != 0 ? 1u : 0u);
                    }
                ptx::umma_commit(&kv_full[bb]);
                ptx::umma_commit(&x_empty[st]);                  // the stage may be refilled once these MMAs have read it
            }
            __syncwarp();
        }
    } else if (warp == KV2_CTX_WARP) {
        // ------------------------------------------------------------------ context MMAs: ctx += P V^T of tile g (its own issuing warp,
        // so that waiting for an epilogue never holds back the K^T / V^T MMAs of the tiles behind it)
        constexpr uint32_t idesc_ctx = ptx::make_idesc_bf16(128, 128);
        const uint64_t dP = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sP));
        const uint64_t dV = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sV));
        int g = 0;
        for (int u = 0; u < my_units; ++u) {
            for (int t = 0; t < T; ++t, ++g) {
                const int bb = g & 1;
                ptx::mbar_wait(&pv_full[bb], (g >> 1) & 1u);
                if (t == 0) ptx::mbar_wait(ctx_empty, (u & 1u) ^ 1u);   // the previous unit's context has been read out
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t off = static_cast<uint64_t>((bb * Cf::PV_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16(tmem_base + COL_CTX, dP + off + 2u * k, dV + off + 2u * k, idesc_ctx, (t | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&pv_empty[bb]);
                    if (t == T - 1) ptx::umma_commit(ctx_full);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 2 + KV2_EPI_WARPS) {
        // ------------------------------------------------------------------ LayerNorm warps: tile g of group (g % NORM_GROUPS), in place
        constexpr int NG = Cf::NORM_GROUPS;
        constexpr int NT = KV2_NORM_WARPS * 32 / NG;    // threads per tile: 128 (two per pixel) or 256 (four per pixel)
        const int wn = warp - 2 - KV2_EPI_WARPS;
        const int grp = wn / (KV2_NORM_WARPS / NG);
        const int tn = (wn % (KV2_NORM_WARPS / NG)) * 32 + lane;
        for (int g = grp; g < G; g += NG) {
            const int st = g % XS;
            ptx::mbar_wait(&x_full[st], (g / XS) & 1u);
            tile_normalize<C, NT, Cf::XSPAN_BYTES>(sX + st * Cf::X_BYTES, tn, a.eps);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&xn_full[st]);
        }
    } else {
        // ------------------------------------------------------------------ epilogue: two groups of eight warps on alternating tiles
        // (group = tile parity = accumulator / P-V buffer index; T is even).  Thread = (row, 32-pixel half).
        const int wd = warp - 2;
        const int grp = wd >> 3;
        const int q = warp & 3;                         // TMEM lane quarter
        const int half = (wd >> 2) & 1;                 // which 32 pixels of the tile
        const int r = q * 32 + lane;                    // d (K^T, ctx) / e (V^T) row
        constexpr float LOG2E = 1.4426950408889634f;
        const ptx::f32x2 L2E = ptx::dup2(LOG2E), NSHIFT = ptx::dup2(-(a.kshift[r] * LOG2E));
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * 128;
        uint8_t* pb = sP + grp * Cf::PV_BYTES;
        uint8_t* vb = sV + grp * Cf::PV_BYTES;
        ptx::grid_dep_wait();
        int u = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u) {
            ptx::f32x2 ssum2 = ptx::mk2(0.f, 0.f);
            for (int t = grp; t < T; t += 2) {
                const uint32_t n = static_cast<uint32_t>(u * T + t) >> 1;     // uses of this group's buffers so far
                ptx::mbar_wait(&kv_full[grp], n & 1u);               // K^T / V^T of the tile ready
                ptx::tc_fence_after();
                // all TMEM reads first, so that the accumulators go back to the MMA warp (tile g + 2) before the exponentials start:
                // V^T (only packed: 32 -> 16 registers), then K^T
                uint32_t vt[32];
                ptx::tmem_ld32(tlane + 64 + half * 32, vt);
                ptx::tmem_ld_wait();
                const uint4 v0 = v_chunk8<0>(vt), v1 = v_chunk8<8>(vt), v2 = v_chunk8<16>(vt), v3 = v_chunk8<24>(vt);
                ptx::tmem_ld32(tlane + half * 32, vt);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&kv_empty[grp]);
                const uint4 p0 = k_chunk8<0>(vt, L2E, NSHIFT, ssum2);
                const uint4 p1 = k_chunk8<8>(vt, L2E, NSHIFT, ssum2);
                const uint4 p2 = k_chunk8<16>(vt, L2E, NSHIFT, ssum2);
                const uint4 p3 = k_chunk8<24>(vt, L2E, NSHIFT, ssum2);
                ptx::mbar_wait(&pv_empty[grp], (n & 1u) ^ 1u);       // the context MMA two tiles back has consumed P / V
                const int ch = half * 4;
                *reinterpret_cast<uint4*>(vb + sw_off(r, ch)) = v0;
                *reinterpret_cast<uint4*>(vb + sw_off(r, ch + 1)) = v1;
                *reinterpret_cast<uint4*>(vb + sw_off(r, ch + 2)) = v2;
                *reinterpret_cast<uint4*>(vb + sw_off(r, ch + 3)) = v3;
                *reinterpret_cast<uint4*>(pb + sw_off(r, ch)) = p0;
                *reinterpret_cast<uint4*>(pb + sw_off(r, ch + 1)) = p1;
                *reinterpret_cast<uint4*>(pb + sw_off(r, ch + 2)) = p2;
                *reinterpret_cast<uint4*>(pb + sw_off(r, ch + 3)) = p3;
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&pv_full[grp]);
            }
            // ---- end of the unit: softmax denominators and the context accumulator
            float* sS = s_S + (u & 1) * 512;            // by unit parity: the slots are rewritten two units on, behind the next barrier
            sS[(grp * 2 + half) * 128 + r] = ptx::lo(ssum2) + ptx::hi(ssum2);
            named_bar_sync(1, KV2_EPI_WARPS * 32);
            if (wd < 4) {
                a.s_part[static_cast<size_t>(unit) * HD + r] = (sS[r] + sS[128 + r]) + (sS[256 + r] + sS[384 + r]);
                ptx::mbar_wait(ctx_full, u & 1u);
                ptx::tc_fence_after();
                uint32_t v[32];
                ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + COL_CTX + q * 32, v);   // row d of head q: columns e of head q
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ctx_empty);
                float4* dst = reinterpret_cast<float4*>(a.ctx_part + (static_cast<size_t>(unit) * HD + r) * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    dst[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                              __uint_as_float(v[j + 3]));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ kernel 2: mix
// One CTA per image, 256 threads, two CTAs per SM: the whole grid (B <= 296 images) is one wave.  Phase 1: the image's partial
// contexts ([parts][128][32] fp32, contiguous) are summed with coalesced 16-byte loads (thread i owns float4 i + 256 k of the 16 KiB
// block; the loads of four parts are in flight together), normalised and parked in shared memory.  Phase 2: thread (hd = t % 128,
// group t / 128) computes C / 2 output channels of column hd; Wo is read through warp-uniform __ldg (one transaction per warp).
// The first form (grid (B, 4), a thread per context ROW) read every partial four times with 128-byte-strided loads and ran as
// 1.4 waves: 24 us per launch, five launches a step (profiles/r02_notes.md 10).
constexpr int MIX_THREADS = 256;
__global__ void __launch_bounds__(MIX_THREADS, 2)
linattn_mix_kernel(const float* __restrict__ ctx_part, const float* __restrict__ s_part, const float* __restrict__ wo,
                   bf16* __restrict__ mb, int parts, int C, float inv_n_scale) {
    __shared__ float s_ctx[HD][33];
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    const int b = blockIdx.x;
    const int t = threadIdx.x;
    {
        const float4* base = reinterpret_cast<const float4*>(ctx_part + static_cast<size_t>(b) * parts * HD * 32);
        const float* sbase = s_part + static_cast<size_t>(b) * parts * HD;
        float4 acc[4];
        float S[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k] = make_float4(0.f, 0.f, 0.f, 0.f); S[k] = 0.f; }
        for (int p0 = 0; p0 < parts; p0 += 4) {            // parts in a fixed order: bit-reproducible
            float4 v[4][4];
            float sv[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const bool on = p0 + pp < parts;
                const int p = on ? p0 + pp : p0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int f = t + k * MIX_THREADS;     // float4 index inside one [128][32] block: row f >> 3, columns (f & 7) * 4 ..
                    v[pp][k] = on ? __ldg(base + static_cast<size_t>(p) * (HD * 8) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                    sv[pp][k] = on ? __ldg(sbase + p * HD + (f >> 3)) : 0.f;
                }
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[k].x += v[pp][k].x; acc[k].y += v[pp][k].y; acc[k].z += v[pp][k].z; acc[k].w += v[pp][k].w;
                    S[k] += sv[pp][k];
                }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int f = t + k * MIX_THREADS;
            const float norm = inv_n_scale / S[k];
            float* dst = &s_ctx[f >> 3][(f & 7) * 4];
            dst[0] = acc[k].x * norm; dst[1] = acc[k].y * norm; dst[2] = acc[k].z * norm; dst[3] = acc[k].w * norm;
        }
    }
    __syncthreads();
    const int hd = t & (HD - 1);
    const int h = hd >> 5;
    const int per = C >> 1;                                // output channels of this thread group (a multiple of 4)
    const int co0 = (t >> 7) * per;
    float c[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) c[e] = s_ctx[hd][e];
    // four output channels at a time: four independent FMA chains; the sum order inside a channel is e = 0 .. 31
    for (int cl = 0; cl < per; cl += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wo + static_cast<size_t>(co0 + cl + u) * HD + h * 32 + 4 * e));
                acc[u] = fmaf(w4.x, c[4 * e], acc[u]); acc[u] = fmaf(w4.y, c[4 * e + 1], acc[u]);
                acc[u] = fmaf(w4.z, c[4 * e + 2], acc[u]); acc[u] = fmaf(w4.w, c[4 * e + 3], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            mb[(static_cast<size_t>(b) * C + co0 + cl + u) * HD + hd] = __float2bfloat16(acc[u]);
    }
}

// ------------------------------------------------------------------------------------------------ kernel 3: output
struct OutArgs {
    int n, tiles_per_img, num_tiles;
    const float* rowsum;     // [384] (q rows first)
    const float* bo;         // [C]
    const float* g2;         // [C]
    float eps;
};

template <int C>
struct OutCfg {
    static constexpr int SPANS = C / 64;
    static constexpr uint32_t WQ_BYTES = SPANS * SPAN_BYTES;
    static constexpr uint32_t X_BYTES = SPANS * SPAN_BYTES;      // one of two input stages; also the output staging (in place)
    static constexpr uint32_t MB_SPAN = C * 128;                 // one 64-wide K span of Mb [C rows]
    static constexpr uint32_t MB_BYTES = 2 * MB_SPAN;
    static constexpr uint32_t A2_BYTES = 2 * SPAN_BYTES;
    static constexpr uint32_t SMALL_BYTES = 512 + 2 * 4 * C + 128 + 4096;    // rowsum, bo, g2, barriers, four [2][128] half-row exchange slots
    static constexpr int SMEM_BYTES = WQ_BYTES + 2 * X_BYTES + MB_BYTES + A2_BYTES + SMALL_BYTES + 1024;
    static constexpr int CTAS_PER_SM = C == 64 ? 2 : 1;
    static constexpr int X_EMPTY_ARRIVALS = C == 64 ? 4 : 8;     // storing warps per tile
};
constexpr int OUT_THREADS = 320;         // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue: (TMEM lane quarter, half of the work)

template <int C>
__global__ void __launch_bounds__(OUT_THREADS, OutCfg<C>::CTAS_PER_SM)
linattn_out_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmY, const OutArgs a) {
    using Cf = OutCfg<C>;
    constexpr int SPANS = Cf::SPANS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    uint8_t* sWq = smem;
    uint8_t* sX = sWq + Cf::WQ_BYTES;                // 2 stages
    uint8_t* sMb = sX + 2 * Cf::X_BYTES;
    uint8_t* sA2 = sMb + Cf::MB_BYTES;
    float* s_sq = reinterpret_cast<float*>(sA2 + Cf::A2_BYTES);
    float* s_bo = s_sq + 128;
    float* s_g2 = s_bo + C;
    float* s_xch = s_g2 + C;                         // [4][2][128]: the two column halves of a row exchange their LayerNorm partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xch + 1024);
    uint64_t* wq_bar = bars;
    uint64_t* x_full = bars + 1;                     // [2]
    uint64_t* x_empty = bars + 3;                    // [2]
    uint64_t* mb_full = bars + 5;
    uint64_t* q_full = bars + 6;
    uint64_t* a2_full = bars + 7;
    uint64_t* y_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmM); ptx::prefetch_tmap(&tmY); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        ptx::mbar_init(wq_bar, 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&x_full[i], 1); ptx::mbar_init(&x_empty[i], Cf::X_EMPTY_ARRIVALS); }
        ptx::mbar_init(mb_full, 1);
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(a2_full, 256);
        ptx::mbar_init(y_full, 1);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    constexpr uint32_t COL_Q = 0, COL_Y = 128;
    const int ntile = a.num_tiles > static_cast<int>(blockIdx.x) ? (a.num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;   // tiles of this CTA

    if (warp == 0) {
        auto load_x = [&](int it) {      // tile `it` of this CTA into stage it & 1 (the stage's previous store has drained)
            const int st = it & 1;
            ptx::mbar_wait(&x_empty[st], ((it >> 1) & 1u) ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&x_full[st], Cf::X_BYTES);
                const int tile = blockIdx.x + it * gridDim.x;
                for (int sp = 0; sp < SPANS; ++sp)
                    ptx::tma_load_2d(sX + st * Cf::X_BYTES + sp * SPAN_BYTES, &tmX, &x_full[st], sp * 64, tile * TILE);
            }
            __syncwarp();
        };
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(wq_bar, Cf::WQ_BYTES);
            for (int sp = 0; sp < SPANS; ++sp) ptx::tma_load_2d(sWq + sp * SPAN_BYTES, &tmW, wq_bar, sp * 64, 0);
        }
        __syncwarp();
        if (ntile > 0) load_x(0);
        for (int it = 0; it < ntile; ++it) {
            if (it > 0) ptx::mbar_wait(y_full, (it - 1) & 1u);       // MMA2 of the previous tile has consumed Mb
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(mb_full, Cf::MB_BYTES);
                const int b = (blockIdx.x + it * gridDim.x) / a.tiles_per_img;
                for (int sp = 0; sp < 2; ++sp) ptx::tma_load_2d(sMb + sp * Cf::MB_SPAN, &tmM, mb_full, sp * 64, b * C);
            }
            __syncwarp();
            if (it + 1 < ntile) load_x(it + 1);                       // prefetch the next tile behind this one's compute
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_q = ptx::make_idesc_bf16(128, 128);
        constexpr uint32_t idesc_y = ptx::make_idesc_bf16(128, C);
        const uint64_t dX = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sX));
        const uint64_t dWq = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWq));
        const uint64_t dA2 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA2));
        const uint64_t dMb = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sMb));
        ptx::mbar_wait(wq_bar, 0);
        for (int it = 0; it < ntile; ++it) {
            const int st = it & 1;
            ptx::mbar_wait(&x_full[st], (it >> 1) & 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t xoff = static_cast<uint64_t>((st * Cf::X_BYTES) >> 4);
#pragma unroll
                for (int sp = 0; sp < SPANS; ++sp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t off = static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                        ptx::umma_bf16(tmem_base + COL_Q, dX + xoff + off, dWq + off, idesc_q, (sp | k) != 0 ? 1u : 0u);
                    }
                ptx::umma_commit(q_full);
            }
            __syncwarp();
            ptx::mbar_wait(a2_full, it & 1u);
            ptx::mbar_wait(mb_full, it & 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offa = static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                        const uint64_t offb = static_cast<uint64_t>((sp * Cf::MB_SPAN) >> 4) + 2u * k;
                        ptx::umma_bf16(tmem_base + COL_Y, dA2 + offa, dMb + offb, idesc_y, (sp | k) != 0 ? 1u : 0u);
                    }
                ptx::umma_commit(y_full);
            }
            __syncwarp();
        }
    } else {
        // epilogue warp (q, hf): pixel row r = q*32 + lane of the tile; hf picks two of the four heads (softmax) and half
        // of the output channels (both LayerNorms); the halves exchange partial sums through shared memory
        const int q = warp & 3;
        const int hf = (warp - 2) >> 2;
        const int te = (warp - 2) * 32 + lane;          // 0..255
        const int r = q * 32 + lane;
        if (te < 128) s_sq[te] = a.rowsum[te];
        if (te < C) { s_bo[te] = a.bo[te]; s_g2[te] = a.g2[te]; }
        named_bar_sync(1, 256);
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // Arithmetic below uses the packed fp32 forms (ptx.cuh: two IEEE lanes per FFMA2 / FADD2 / FMUL2): the epilogue is issue-bound.
        constexpr int HALF = C / 2;                      // channels of a row this thread owns: [hf * HALF, (hf + 1) * HALF)
        constexpr int HCH = HALF / 8;                    // ... as 16-byte chunks: half of span 0 (C = 64) or span hf (C = 128)
        constexpr float LOG2E = 1.4426950408889634f;
        float* xa = s_xch;                               // four [2][128] exchange slots: x sum, x centred squares, y sum, y squares
        float* xb = s_xch + 256;
        float* xc = s_xch + 512;
        float* xd = s_xch + 768;
        const int mine = hf * 128 + r, other = (hf ^ 1) * 128 + r;
        for (int it = 0; it < ntile; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int st = it & 1;
            uint8_t* xs = sX + st * Cf::X_BYTES;
            ptx::mbar_wait(&x_full[st], (it >> 1) & 1u);
            // ---- LayerNorm statistics of pixel row r (two-pass): each of the row's two threads reduces its half of the channels
            // from registers and the halves exchange their partial sums through shared memory (a + b == b + a: both get the same bits)
            float mean, rstd;
            {
                const uint8_t* xrow = xs + (C == 64 ? 0 : hf * SPAN_BYTES);
                const int ch0 = C == 64 ? hf * 4 : 0;
                ptx::f32x2 xh[HALF / 2];
#pragma unroll
                for (int c = 0; c < HCH; ++c) {
                    const uint4 u = *reinterpret_cast<const uint4*>(xrow + sw_off(r, ch0 + c));
                    xh[c * 4] = ptx::bf16x2_to_f32x2(u.x); xh[c * 4 + 1] = ptx::bf16x2_to_f32x2(u.y);
                    xh[c * 4 + 2] = ptx::bf16x2_to_f32x2(u.z); xh[c * 4 + 3] = ptx::bf16x2_to_f32x2(u.w);
                }
                ptx::f32x2 s0 = xh[0], s1 = xh[1];
#pragma unroll
                for (int i = 2; i < HALF / 2; i += 2) { s0 = ptx::add2(s0, xh[i]); s1 = ptx::add2(s1, xh[i + 1]); }
                s0 = ptx::add2(s0, s1);
                const float psum = ptx::lo(s0) + ptx::hi(s0);
                xa[mine] = psum;
                named_bar_sync(2 + q, 64);
                mean = (psum + xa[other]) * (1.0f / C);
                const ptx::f32x2 nm = ptx::dup2(-mean);
                ptx::f32x2 q0 = ptx::mk2(0.f, 0.f), q1 = q0;
#pragma unroll
                for (int i = 0; i < HALF / 2; i += 2) {
                    const ptx::f32x2 d0 = ptx::add2(xh[i], nm), d1 = ptx::add2(xh[i + 1], nm);
                    q0 = ptx::fma2(d0, d0, q0);
                    q1 = ptx::fma2(d1, d1, q1);
                }
                q0 = ptx::add2(q0, q1);
                const float pss = ptx::lo(q0) + ptx::hi(q0);
                xb[mine] = pss;
                named_bar_sync(2 + q, 64);
                rstd = rsqrtf((pss + xb[other]) * (1.0f / C) + a.eps);
            }
            ptx::mbar_wait(q_full, it & 1u);
            ptx::tc_fence_after();
            // q = rstd * (acc - mean * rowsum) and exp(q - max) = exp2(q * log2e - max'): al = rstd * log2e, ncl = -mean * al
            // (softmax is invariant to the common shift, the max is taken after scaling)
            const ptx::f32x2 AL = ptx::dup2(rstd * LOG2E), NCL = ptx::dup2(-(mean * (rstd * LOG2E)));
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                const int h = 2 * hf + hh;
                uint32_t v[32];
                ptx::tmem_ld32(tlane + COL_Q + h * 32, v);
                ptx::tmem_ld_wait();
                ptx::f32x2 f[16];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float4 s4 = *reinterpret_cast<const float4*>(s_sq + h * 32 + 2 * j);
                    f[j] = ptx::fma2(ptx::mk2(v[2 * j], v[2 * j + 1]), AL, ptx::mul2(NCL, ptx::mk2(s4.x, s4.y)));
                    f[j + 1] = ptx::fma2(ptx::mk2(v[2 * j + 2], v[2 * j + 3]), AL, ptx::mul2(NCL, ptx::mk2(s4.z, s4.w)));
                    mx = fmaxf(fmaxf(fmaxf(mx, ptx::lo(f[j])), fmaxf(ptx::hi(f[j]), ptx::lo(f[j + 1]))), ptx::hi(f[j + 1]));
                }
                const ptx::f32x2 NMX = ptx::dup2(-mx);
                ptx::f32x2 a0 = ptx::mk2(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const ptx::f32x2 d0 = ptx::add2(f[j], NMX), d1 = ptx::add2(f[j + 1], NMX);
                    f[j] = ptx::mk2(ptx::ex2(ptx::lo(d0)), ptx::ex2(ptx::hi(d0)));
                    f[j + 1] = ptx::mk2(ptx::ex2(ptx::lo(d1)), ptx::ex2(ptx::hi(d1)));
                    a0 = ptx::add2(a0, f[j]);
                    a1 = ptx::add2(a1, f[j + 1]);
                }
                a0 = ptx::add2(a0, a1);
                const ptx::f32x2 INV = ptx::dup2(ptx::rcp(ptx::lo(a0) + ptx::hi(a0)));   // the sum is in [1, 32]
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint4 o;
                    o.x = ptx::pack_bf16x2(ptx::mul2(f[jj * 4], INV));
                    o.y = ptx::pack_bf16x2(ptx::mul2(f[jj * 4 + 1], INV));
                    o.z = ptx::pack_bf16x2(ptx::mul2(f[jj * 4 + 2], INV));
                    o.w = ptx::pack_bf16x2(ptx::mul2(f[jj * 4 + 3], INV));
                    *reinterpret_cast<uint4*>(sA2 + hf * SPAN_BYTES + sw_off(r, hh * 4 + jj)) = o;   // span = h >> 1 = hf
                }
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(a2_full);

            ptx::mbar_wait(y_full, it & 1u);
            ptx::tc_fence_after();
            // LayerNorm over the C output channels of this pixel.  This thread owns the channels [hf * C/2, (hf + 1) * C/2) of row r:
            // one TMEM read into registers (+ bias), then the two halves of the row exchange their partial sums through shared
            // memory (mean first, then the centred sum of squares: two-pass accuracy without reading the other half).
            ptx::f32x2 y[HALF / 2];
            ptx::f32x2 ys0 = ptx::mk2(0.f, 0.f), ys1 = ys0;
#pragma unroll
            for (int cc = 0; cc < HALF / 32; ++cc) {
                const int c32 = hf * (HALF / 32) + cc;
                uint32_t v[32];
                ptx::tmem_ld32(tlane + COL_Y + c32 * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bo + c32 * 32 + 2 * j);
                    y[cc * 16 + j] = ptx::add2(ptx::mk2(v[2 * j], v[2 * j + 1]), ptx::mk2(b4.x, b4.y));
                    y[cc * 16 + j + 1] = ptx::add2(ptx::mk2(v[2 * j + 2], v[2 * j + 3]), ptx::mk2(b4.z, b4.w));
                    ys0 = ptx::add2(ys0, y[cc * 16 + j]);
                    ys1 = ptx::add2(ys1, y[cc * 16 + j + 1]);
                }
            }
            ys0 = ptx::add2(ys0, ys1);
            const float sum = ptx::lo(ys0) + ptx::hi(ys0);
            xc[mine] = sum;
            named_bar_sync(2 + q, 64);
            const float ymean = (sum + xc[other]) * (1.0f / C);
            const ptx::f32x2 NYM = ptx::dup2(-ymean);
            ptx::f32x2 yq0 = ptx::mk2(0.f, 0.f), yq1 = yq0;
#pragma unroll
            for (int j = 0; j < HALF / 2; j += 2) {
                y[j] = ptx::add2(y[j], NYM);
                y[j + 1] = ptx::add2(y[j + 1], NYM);
                yq0 = ptx::fma2(y[j], y[j], yq0);
                yq1 = ptx::fma2(y[j + 1], y[j + 1], yq1);
            }
            yq0 = ptx::add2(yq0, yq1);
            const float ss = ptx::lo(yq0) + ptx::hi(yq0);
            xd[mine] = ss;
            named_bar_sync(2 + q, 64);
            const ptx::f32x2 YR = ptx::dup2(rsqrtf((ss + xd[other]) * (1.0f / C) + a.eps));
            // normalise, + x, and write the result IN PLACE over the x tile (same row, same swizzled chunk), which then doubles
            // as the TMA-store staging buffer
#pragma unroll
            for (int cc = 0; cc < HALF / 32; ++cc) {
                const int c32 = hf * (HALF / 32) + cc;
                const int sp = c32 >> 1;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int ch = (c32 & 1) * 4 + jj;                       // 16-byte chunk inside the span
                    uint4* px = reinterpret_cast<uint4*>(xs + sp * SPAN_BYTES + sw_off(r, ch));
                    const uint4 xu = *px;
                    const float4 ga = *reinterpret_cast<const float4*>(s_g2 + c32 * 32 + jj * 8);
                    const float4 gb = *reinterpret_cast<const float4*>(s_g2 + c32 * 32 + jj * 8 + 4);
                    const int k = cc * 16 + jj * 4;
                    uint4 u;                                                  // o = (y * rstd) * g2 + x
                    u.x = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k], YR), ptx::mk2(ga.x, ga.y), ptx::bf16x2_to_f32x2(xu.x)));
                    u.y = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k + 1], YR), ptx::mk2(ga.z, ga.w), ptx::bf16x2_to_f32x2(xu.y)));
                    u.z = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k + 2], YR), ptx::mk2(gb.x, gb.y), ptx::bf16x2_to_f32x2(xu.z)));
                    u.w = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k + 3], YR), ptx::mk2(gb.z, gb.w), ptx::bf16x2_to_f32x2(xu.w)));
                    *px = u;
                }
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            if constexpr (C == 64) {
                named_bar_sync(2 + q, 64);               // both halves of this row block are in place
                if (hf == 0 && lane == 0) {
                    ptx::tma_store_2d(&tmY, xs + (q * 32) * 128, 0, tile * TILE + q * 32);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<0>();             // the store has read the stage: it may be refilled
                    ptx::mbar_arrive(&x_empty[st]);
                }
            } else {
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmY, xs + hf * SPAN_BYTES + (q * 32) * 128, hf * 64, tile * TILE + q * 32);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<0>();
                    ptx::mbar_arrive(&x_empty[st]);
                }
            }
        }
        if (lane == 0) ptx::bulk_wait_all();
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ kernel 3, pipelined form (C = 64)
// linattn_out2_kernel: the same math as linattn_out_kernel as five decoupled pipelines inside one persistent CTA per SM, meeting only
// in mbarriers, so that no role's latency chain is on another's critical path (the first form runs LayerNorm statistics -> MMA ->
// softmax -> MMA -> LayerNorm -> store of a tile back to back in the same eight warps, and only the second CTA of the SM fills the gaps):
//   warp 0        TMA: x tiles (3 stages; a stage doubles as the output staging buffer) and the tile's image matrix Mb (2 buffers)
//   warps 19..26  LayerNorm of the tile, two threads per pixel row, written as bf16 z = LN(x) / gain into a second buffer (2 buffers):
//                 Q = z Wq'^T then needs no per-row correction, and x itself stays intact for the residual
//   warp 1        MMA 1: Q[i & 1] = z Wq'^T                (TMEM columns 0 / 128)
//   warps 3..10   softmax over each head's 32 columns of Q (thread = pixel row, two heads) -> bf16 A2[i & 1]
//   warp 2        MMA 2: Y[i & 1] = A2 Mb^T                (TMEM columns 256 / 320)
//   warps 11..18  + bias, LayerNorm g2, + x, in place into the x stage -> TMA store (thread = pixel row, half of the channels)
constexpr int OUT2_C = 64;
struct Out2Cfg {
    static constexpr uint32_t WQ_BYTES = SPAN_BYTES;
    static constexpr uint32_t X_BYTES = SPAN_BYTES;                   // one stage: [128 px][64 ch] bf16
    static constexpr int X_STAGES = 3;
    static constexpr uint32_t MB_SPAN = OUT2_C * 128;                 // one 64-wide K span of Mb [C rows]
    static constexpr uint32_t MB_BYTES = 2 * MB_SPAN;
    static constexpr uint32_t A2_BYTES = 2 * SPAN_BYTES;
    static constexpr uint32_t SMALL_BYTES = 2 * 4 * OUT2_C + 2048 + 256;     // bo, g2, two [2][128] exchange slots, barriers
    static constexpr int SMEM_BYTES = WQ_BYTES + X_STAGES * X_BYTES + 2 * X_BYTES + 2 * MB_BYTES + 2 * A2_BYTES + SMALL_BYTES + 1024;
};
constexpr int OUT2_SM_WARP0 = 3, OUT2_LN_WARP0 = 11, OUT2_NORM_WARP0 = 19;
constexpr int OUT2_THREADS = 27 * 32;

// LayerNorm (no gain) of ROWS pixel rows of a swizzled [ROWS x C] bf16 tile by NT threads (NT / ROWS consecutive lanes per row), from
// src to dst (same layout; dst == src works in place); see tile_normalize.
template <int C, int NT, int ROWS, uint32_t SPAN_STRIDE>
__device__ __forceinline__ void rows_normalize(const uint8_t* src, uint8_t* dst, int te, float eps) {
    constexpr int TPP = NT / ROWS;
    constexpr int NCH = (C / 8) / TPP;
    static_assert(NCH >= 1 && TPP >= 1 && TPP <= 8, "rows_normalize: thread split");
    const int p = te / TPP, sub = te % TPP;
    ptx::f32x2 xh[NCH * 4];
    uint32_t off[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = sub + i * TPP;
        off[i] = (c >> 3) * SPAN_STRIDE + sw_off(p, c & 7);
        const uint4 u = *reinterpret_cast<const uint4*>(src + off[i]);
        xh[i * 4] = ptx::bf16x2_to_f32x2(u.x); xh[i * 4 + 1] = ptx::bf16x2_to_f32x2(u.y);
        xh[i * 4 + 2] = ptx::bf16x2_to_f32x2(u.z); xh[i * 4 + 3] = ptx::bf16x2_to_f32x2(u.w);
    }
    ptx::f32x2 s0 = xh[0], s1 = xh[1];
#pragma unroll
    for (int i = 2; i < NCH * 4; i += 2) { s0 = ptx::add2(s0, xh[i]); s1 = ptx::add2(s1, xh[i + 1]); }
    s0 = ptx::add2(s0, s1);
    float sum = ptx::lo(s0) + ptx::hi(s0);
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const ptx::f32x2 nm = ptx::dup2(-(sum * (1.0f / C)));
    ptx::f32x2 q0 = ptx::mk2(0.f, 0.f), q1 = q0;
#pragma unroll
    for (int i = 0; i < NCH * 4; i += 2) {
        xh[i] = ptx::add2(xh[i], nm);
        xh[i + 1] = ptx::add2(xh[i + 1], nm);
        q0 = ptx::fma2(xh[i], xh[i], q0);
        q1 = ptx::fma2(xh[i + 1], xh[i + 1], q1);
    }
    q0 = ptx::add2(q0, q1);
    float ss = ptx::lo(q0) + ptx::hi(q0);
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const ptx::f32x2 RS = ptx::dup2(rsqrtf(ss * (1.0f / C) + eps));
#pragma unroll
    for (int i = 0; i < NCH; ++i)
        *reinterpret_cast<uint4*>(dst + off[i]) =
            make_uint4(ptx::pack_bf16x2(ptx::mul2(xh[i * 4], RS)), ptx::pack_bf16x2(ptx::mul2(xh[i * 4 + 1], RS)),
                       ptx::pack_bf16x2(ptx::mul2(xh[i * 4 + 2], RS)), ptx::pack_bf16x2(ptx::mul2(xh[i * 4 + 3], RS)));
}

__global__ void __launch_bounds__(OUT2_THREADS, 1)
linattn_out2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmY, const OutArgs a) {
    using Cf = Out2Cfg;
    constexpr int C = OUT2_C;
    constexpr int XS = Cf::X_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sWq = smem;
    uint8_t* sX = sWq + Cf::WQ_BYTES;                // [3] raw x / output staging
    uint8_t* sZ = sX + XS * Cf::X_BYTES;             // [2] normalised x
    uint8_t* sMb = sZ + 2 * Cf::X_BYTES;             // [2]
    uint8_t* sA2 = sMb + 2 * Cf::MB_BYTES;           // [2]
    float* s_bo = reinterpret_cast<float*>(sA2 + 2 * Cf::A2_BYTES);
    float* s_g2 = s_bo + C;
    float* s_xch = s_g2 + C;                         // [2 slots][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xch + 512);
    uint64_t* wq_bar = bars;
    uint64_t* x_full = bars + 1;                     // [3] TMA -> LayerNorm warps (and the epilogue's residual read)
    uint64_t* x_empty = bars + 4;                    // [3] output store has read the stage -> TMA
    uint64_t* z_full = bars + 7;                     // [2] LayerNorm warps -> MMA 1
    uint64_t* z_empty = bars + 9;                    // [2] MMA 1 -> LayerNorm warps
    uint64_t* q_full = bars + 11;                    // [2] MMA 1 -> softmax warps
    uint64_t* q_empty = bars + 13;                   // [2] softmax warps -> MMA 1
    uint64_t* a2_full = bars + 15;                   // [2] softmax warps -> MMA 2
    uint64_t* a2_empty = bars + 17;                  // [2] MMA 2 -> softmax warps
    uint64_t* mb_full = bars + 19;                   // [2] TMA -> MMA 2
    uint64_t* mb_empty = bars + 21;                  // [2] MMA 2 -> TMA
    uint64_t* y_full = bars + 23;                    // [2] MMA 2 -> epilogue warps
    uint64_t* y_empty = bars + 25;                   // [2] epilogue warps -> MMA 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmM); ptx::prefetch_tmap(&tmY); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        ptx::mbar_init(wq_bar, 1);
        for (int i = 0; i < XS; ++i) { ptx::mbar_init(&x_full[i], 1); ptx::mbar_init(&x_empty[i], 4); }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&z_full[i], 8);  ptx::mbar_init(&z_empty[i], 1);
            ptx::mbar_init(&q_full[i], 1);  ptx::mbar_init(&q_empty[i], 8);
            ptx::mbar_init(&a2_full[i], 8); ptx::mbar_init(&a2_empty[i], 1);
            ptx::mbar_init(&mb_full[i], 1); ptx::mbar_init(&mb_empty[i], 1);
            ptx::mbar_init(&y_full[i], 1);  ptx::mbar_init(&y_empty[i], 8);
        }
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; each role waits for the predecessor before global memory
    constexpr uint32_t COL_Q = 0, COL_Y = 256;
    const int ntile = a.num_tiles > static_cast<int>(blockIdx.x) ? (a.num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;   // tiles of this CTA

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(wq_bar, Cf::WQ_BYTES);
            ptx::tma_load_2d(sWq, &tmW, wq_bar, 0, 0);
        }
        __syncwarp();
        ptx::grid_dep_wait();
        for (int it = 0; it < ntile; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int st = it % XS, mb = it & 1;
            ptx::mbar_wait(&x_empty[st], ((it / XS) & 1u) ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&x_full[st], Cf::X_BYTES);
                ptx::tma_load_2d(sX + st * Cf::X_BYTES, &tmX, &x_full[st], 0, tile * TILE);
            }
            __syncwarp();
            ptx::mbar_wait(&mb_empty[mb], ((it >> 1) & 1u) ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&mb_full[mb], Cf::MB_BYTES);
                const int b = tile / a.tiles_per_img;
                for (int sp = 0; sp < 2; ++sp) ptx::tma_load_2d(sMb + mb * Cf::MB_BYTES + sp * Cf::MB_SPAN, &tmM, &mb_full[mb], sp * 64, b * C);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA 1: Q = z Wq'^T
        constexpr uint32_t idesc_q = ptx::make_idesc_bf16(128, 128);
        const uint64_t dZ = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sZ));
        const uint64_t dWq = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWq));
        ptx::mbar_wait(wq_bar, 0);
        for (int it = 0; it < ntile; ++it) {
            const int bb = it & 1;
            const uint32_t ph = (it >> 1) & 1u;
            ptx::mbar_wait(&z_full[bb], ph);
            ptx::mbar_wait(&q_empty[bb], ph ^ 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t zoff = static_cast<uint64_t>((bb * Cf::X_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_bf16(tmem_base + COL_Q + bb * 128, dZ + zoff + 2u * k, dWq + 2u * k, idesc_q, k != 0 ? 1u : 0u);
                ptx::umma_commit(&q_full[bb]);
                ptx::umma_commit(&z_empty[bb]);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ MMA 2: Y = A2 Mb^T
        constexpr uint32_t idesc_y = ptx::make_idesc_bf16(128, C);
        const uint64_t dA2 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA2));
        const uint64_t dMb = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sMb));
        for (int it = 0; it < ntile; ++it) {
            const int bb = it & 1;
            const uint32_t ph = (it >> 1) & 1u;
            ptx::mbar_wait(&a2_full[bb], ph);
            ptx::mbar_wait(&mb_full[bb], ph);
            ptx::mbar_wait(&y_empty[bb], ph ^ 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t aoff = static_cast<uint64_t>((bb * Cf::A2_BYTES) >> 4);
                const uint64_t moff = static_cast<uint64_t>((bb * Cf::MB_BYTES) >> 4);
#pragma unroll
                for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offa = aoff + static_cast<uint64_t>((sp * SPAN_BYTES) >> 4) + 2u * k;
                        const uint64_t offb = moff + static_cast<uint64_t>((sp * Cf::MB_SPAN) >> 4) + 2u * k;
                        ptx::umma_bf16(tmem_base + COL_Y + bb * 64, dA2 + offa, dMb + offb, idesc_y, (sp | k) != 0 ? 1u : 0u);
                    }
                ptx::umma_commit(&y_full[bb]);
                ptx::umma_commit(&a2_empty[bb]);
                ptx::umma_commit(&mb_empty[bb]);
            }
            __syncwarp();
        }
    } else if (warp >= OUT2_NORM_WARP0) {
        // ------------------------------------------------------------------ LayerNorm warps: z = (x - mean) * rstd, two threads per row
        const int tn = (warp - OUT2_NORM_WARP0) * 32 + lane;       // 0..255
        for (int it = 0; it < ntile; ++it) {
            const int st = it % XS, bb = it & 1;
            ptx::mbar_wait(&x_full[st], (it / XS) & 1u);
            ptx::mbar_wait(&z_empty[bb], ((it >> 1) & 1u) ^ 1u);
            rows_normalize<C, 256, TILE, SPAN_BYTES>(sX + st * Cf::X_BYTES, sZ + bb * Cf::X_BYTES, tn, a.eps);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&z_full[bb]);
        }
    } else if (warp < OUT2_LN_WARP0) {
        // ------------------------------------------------------------------ softmax warps: thread = (pixel row, two heads)
        const int q = warp & 3;
        const int hf = (warp - OUT2_SM_WARP0) >> 2;
        const int r = q * 32 + lane;
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        constexpr float LOG2E = 1.4426950408889634f;
        const ptx::f32x2 L2E = ptx::dup2(LOG2E);
        for (int it = 0; it < ntile; ++it) {
            const int bb = it & 1;
            const uint32_t ph = (it >> 1) & 1u;
            uint8_t* a2 = sA2 + bb * Cf::A2_BYTES + hf * SPAN_BYTES;      // span = head >> 1 = hf
            ptx::mbar_wait(&q_full[bb], ph);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[32];
                ptx::tmem_ld32(tlane + COL_Q + bb * 128 + (2 * hf + hh) * 32, v);
                ptx::tmem_ld_wait();
                if (hh == 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&q_empty[bb]);       // this warp has read its part of Q: MMA 1 two tiles on may overwrite
                }
                float mx = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
                for (int j = 2; j < 32; j += 2) mx = fmaxf(fmaxf(mx, __uint_as_float(v[j])), __uint_as_float(v[j + 1]));
                // exp(q - max) = exp2(q * log2e - max * log2e)
                const ptx::f32x2 NMX = ptx::dup2(-(mx * LOG2E));
                ptx::f32x2 f[16];
                ptx::f32x2 a0 = ptx::mk2(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const ptx::f32x2 d0 = ptx::fma2(ptx::mk2(v[2 * j], v[2 * j + 1]), L2E, NMX);
                    const ptx::f32x2 d1 = ptx::fma2(ptx::mk2(v[2 * j + 2], v[2 * j + 3]), L2E, NMX);
                    f[j] = ptx::mk2(ptx::ex2(ptx::lo(d0)), ptx::ex2(ptx::hi(d0)));
                    f[j + 1] = ptx::mk2(ptx::ex2(ptx::lo(d1)), ptx::ex2(ptx::hi(d1)));
                    a0 = ptx::add2(a0, f[j]);
                    a1 = ptx::add2(a1, f[j + 1]);
                }
                a0 = ptx::add2(a0, a1);
                const ptx::f32x2 INV = ptx::dup2(ptx::rcp(ptx::lo(a0) + ptx::hi(a0)));   // the sum is in [1, 32]
                if (hh == 0) ptx::mbar_wait(&a2_empty[bb], ph ^ 1u);     // MMA 2 two tiles back has consumed this A2 buffer
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint4 o;
                    o.x = ptx::pack_bf16x2(ptx::mul2(f[jj * 4], INV));
                    o.y = ptx::pack_bf16x2(ptx::mul2(f[jj * 4 + 1], INV));
                    o.z = ptx::pack_bf16x2(ptx::mul2(f[jj * 4 + 2], INV));
                    o.w = ptx::pack_bf16x2(ptx::mul2(f[jj * 4 + 3], INV));
                    *reinterpret_cast<uint4*>(a2 + sw_off(r, hh * 4 + jj)) = o;
                }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&a2_full[bb]);
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps: thread = (pixel row, half of the channels)
        const int q = warp & 3;
        const int hf = (warp - OUT2_LN_WARP0) >> 2;
        const int te = (warp - OUT2_LN_WARP0) * 32 + lane;
        const int r = q * 32 + lane;
        if (te < C) { s_bo[te] = a.bo[te]; s_g2[te] = a.g2[te]; }
        named_bar_sync(1, 256);
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* xc = s_xch;                               // two [2][128] exchange slots: y sum, y centred squares
        float* xd = s_xch + 256;
        const int mine = hf * 128 + r, other = (hf ^ 1) * 128 + r;
        ptx::grid_dep_wait();
        for (int it = 0; it < ntile; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int st = it % XS, bb = it & 1;
            uint8_t* xs = sX + st * Cf::X_BYTES;
            ptx::mbar_wait(&x_full[st], (it / XS) & 1u);             // long complete: makes the TMA's writes visible to this thread
            ptx::mbar_wait(&y_full[bb], (it >> 1) & 1u);
            ptx::tc_fence_after();
            uint32_t v[32];
            ptx::tmem_ld32(tlane + COL_Y + bb * 64 + hf * 32, v);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&y_empty[bb]);
            // LayerNorm over the C output channels of this pixel: the two halves of the row exchange their partial sums through
            // shared memory (mean first, then the centred sum of squares)
            ptx::f32x2 y[16];
            ptx::f32x2 ys0 = ptx::mk2(0.f, 0.f), ys1 = ys0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_bo + hf * 32 + 2 * j);
                y[j] = ptx::add2(ptx::mk2(v[2 * j], v[2 * j + 1]), ptx::mk2(b4.x, b4.y));
                y[j + 1] = ptx::add2(ptx::mk2(v[2 * j + 2], v[2 * j + 3]), ptx::mk2(b4.z, b4.w));
                ys0 = ptx::add2(ys0, y[j]);
                ys1 = ptx::add2(ys1, y[j + 1]);
            }
            ys0 = ptx::add2(ys0, ys1);
            const float sum = ptx::lo(ys0) + ptx::hi(ys0);
            xc[mine] = sum;
            named_bar_sync(2 + q, 64);
            const ptx::f32x2 NYM = ptx::dup2(-((sum + xc[other]) * (1.0f / C)));
            ptx::f32x2 yq0 = ptx::mk2(0.f, 0.f), yq1 = yq0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                y[j] = ptx::add2(y[j], NYM);
                y[j + 1] = ptx::add2(y[j + 1], NYM);
                yq0 = ptx::fma2(y[j], y[j], yq0);
                yq1 = ptx::fma2(y[j + 1], y[j + 1], yq1);
            }
            yq0 = ptx::add2(yq0, yq1);
            const float ss = ptx::lo(yq0) + ptx::hi(yq0);
            xd[mine] = ss;
            named_bar_sync(2 + q, 64);
            const ptx::f32x2 YR = ptx::dup2(rsqrtf((ss + xd[other]) * (1.0f / C) + a.eps));
            // normalise, + x, and write the result IN PLACE over the x tile (same row, same swizzled chunk): the TMA-store staging
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                uint4* px = reinterpret_cast<uint4*>(xs + sw_off(r, hf * 4 + jj));
                const uint4 xu = *px;
                const float4 ga = *reinterpret_cast<const float4*>(s_g2 + hf * 32 + jj * 8);
                const float4 gb = *reinterpret_cast<const float4*>(s_g2 + hf * 32 + jj * 8 + 4);
                const int k = jj * 4;
                uint4 u;                                                  // o = (y * rstd) * g2 + x
                u.x = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k], YR), ptx::mk2(ga.x, ga.y), ptx::bf16x2_to_f32x2(xu.x)));
                u.y = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k + 1], YR), ptx::mk2(ga.z, ga.w), ptx::bf16x2_to_f32x2(xu.y)));
                u.z = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k + 2], YR), ptx::mk2(gb.x, gb.y), ptx::bf16x2_to_f32x2(xu.z)));
                u.w = ptx::pack_bf16x2(ptx::fma2(ptx::mul2(y[k + 3], YR), ptx::mk2(gb.z, gb.w), ptx::bf16x2_to_f32x2(xu.w)));
                *px = u;
            }
            ptx::fence_proxy_async_smem();
            named_bar_sync(2 + q, 64);                   // both halves of this 32-row block are in place
            if (hf == 0 && lane == 0) {
                ptx::tma_store_2d(&tmY, xs + (q * 32) * 128, 0, tile * TILE + q * 32);
                ptx::bulk_commit();
                ptx::bulk_wait_read<0>();                 // the store has read the stage: it may be refilled
                ptx::mbar_arrive(&x_empty[st]);
            }
        }
        if (hf == 0 && lane == 0) ptx::bulk_wait_all();
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
cudaError_t linattn_prep_run(const float* wqkv, const float* g1, bf16* out, float* rowsum, float* kshift,
                             int* max_bound_bits, int C, cudaStream_t s) {
    if (C != 64 && C != 128) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(max_bound_bits, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    linattn_prep_kernel<<<3 * HD, 128, 0, s>>>(wqkv, g1, out, rowsum, kshift, max_bound_bits, C);
    return cudaGetLastError();
}

int linattn_fused_prepare(const LinAttnFusedDesc& d, int num_sms, LinAttnFusedLaunch* out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
    if ((d.C != 64 && d.C != 128) || d.n < TILE || d.n % TILE != 0 || d.B <= 0) {
        snprintf(err, errlen, "linattn_fused: unsupported shape B=%d n=%d C=%d", d.B, d.n, d.C);
        return 1;
    }
    out->d = d;
    out->num_sms = num_sms;
    const int tiles_per_img = d.n / TILE;
    int parts = 1;                      // split images until there are >= ~4 CTAs per SM or one tile per CTA
    while (parts < tiles_per_img && d.B * parts < 4 * num_sms && tiles_per_img % (parts * 2) == 0) parts *= 2;
    if (parts > d.max_parts) parts = d.max_parts;
    out->parts = parts;
    out->tiles_per_unit = tiles_per_img / parts;
    out->num_tiles = d.B * tiles_per_img;
    const int per_sm = d.C == 64 ? 2 : 1;
    out->out_grid = out->num_tiles < per_sm * num_sms ? out->num_tiles : per_sm * num_sms;
    const cuuint64_t M = static_cast<cuuint64_t>(d.B) * d.n;
    {
        cuuint64_t dims[2] = {(cuuint64_t)d.C, M};
        cuuint64_t str[1] = {(cuuint64_t)d.C * 2};
        cuuint32_t box[2] = {64, TILE};
        if (encode_tmap_bf16(&out->tmX, d.x, 2, dims, str, box, err, errlen)) return 1;
        cuuint32_t boxk[2] = {64, KV_PX};
        if (encode_tmap_bf16(&out->tmXk, d.x, 2, dims, str, boxk, err, errlen)) return 1;
        cuuint32_t boxy[2] = {64, 32};
        if (encode_tmap_bf16(&out->tmY, d.y, 2, dims, str, boxy, err, errlen)) return 1;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d.C, 3 * HD};
        cuuint64_t str[1] = {(cuuint64_t)d.C * 2};
        cuuint32_t box[2] = {64, 128};
        if (encode_tmap_bf16(&out->tmW, d.wqkv, 2, dims, str, box, err, errlen)) return 1;
    }
    {
        cuuint64_t dims[2] = {HD, (cuuint64_t)d.B * d.C};
        cuuint64_t str[1] = {HD * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)d.C};
        if (encode_tmap_bf16(&out->tmM, d.mb, 2, dims, str, box, err, errlen)) return 1;
    }
    return 0;
}

template <int C>
static cudaError_t run_c(const LinAttnFusedLaunch& l, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(linattn_kv_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, KvCfg<C>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(linattn_kv2_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Kv2Cfg<C>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(linattn_out_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, OutCfg<C>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(linattn_out2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Out2Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const LinAttnFusedDesc& d = l.d;
    KvArgs ka;
    ka.n = d.n; ka.tiles_per_unit = l.tiles_per_unit * (TILE / KV_PX); ka.parts = l.parts; ka.kshift = d.kshift;
    ka.ctx_part = d.ctx_part; ka.s_part = d.s_part; ka.eps = d.eps;
    static const bool kv_v1 = [] { const char* v = getenv("HD_LA_KV"); return v && v[0] == '1'; }();
    cudaError_t e;
    if (kv_v1) {
        e = launch_pdl(linattn_kv_kernel<C>, dim3(d.B * l.parts), dim3(KV_THREADS), KvCfg<C>::SMEM_BYTES, s, l.tmXk, l.tmW, ka);
    } else {
        const int units = d.B * l.parts;
        const int grid = units < l.num_sms ? units : l.num_sms;     // persistent: one CTA per SM walks its units
        e = launch_pdl(linattn_kv2_kernel<C>, dim3(grid), dim3(KV2_THREADS), Kv2Cfg<C>::SMEM_BYTES, s, l.tmXk, l.tmW, ka, units);
    }
    if (e != cudaSuccess) return e;
    e = launch_pdl(linattn_mix_kernel, dim3(d.B), dim3(MIX_THREADS), 0, s, d.ctx_part, d.s_part, d.wo, d.mb, l.parts, C,
                   0.17677669529663687f / static_cast<float>(d.n));   // 32^-0.5 / n
    if (e != cudaSuccess) return e;
    OutArgs oa;
    oa.n = d.n; oa.tiles_per_img = d.n / TILE; oa.num_tiles = l.num_tiles; oa.rowsum = d.rowsum; oa.bo = d.bo; oa.g2 = d.g2;
    oa.eps = d.eps;
    static const bool out_v1 = [] { const char* v = getenv("HD_LA_OUT"); return v && v[0] == '1'; }();
    if (C == OUT2_C && !out_v1) {
        const int grid = l.num_tiles < l.num_sms ? l.num_tiles : l.num_sms;     // persistent: one CTA per SM
        return launch_pdl(linattn_out2_kernel, dim3(grid), dim3(OUT2_THREADS), Out2Cfg::SMEM_BYTES, s, l.tmX, l.tmW, l.tmM, l.tmY, oa);
    }
    return launch_pdl(linattn_out_kernel<C>, dim3(l.out_grid), dim3(OUT_THREADS), OutCfg<C>::SMEM_BYTES, s, l.tmX, l.tmW, l.tmM, l.tmY, oa);
}

cudaError_t linattn_fused_run(const LinAttnFusedLaunch& l, cudaStream_t s) {
    return l.d.C == 64 ? run_c<64>(l, s) : run_c<128>(l, s);
}

}  // namespace hd
