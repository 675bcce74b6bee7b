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
//   linattn_kv_kernel   (one CTA per (image, pixel range)): for every 128-pixel tile, TMA-load x, tcgen05 GEMMs
//       K^T[d, px] = Wk' x^T and V^T[e, px] = Wv' x^T (weights pre-multiplied by the LayerNorm gain, the LayerNorm's mean /
//       rstd applied per pixel COLUMN in the epilogue: W LN(x) = rstd * (W' x - mu * rowsum(W'))), p = exp(k - shift_d)
//       written bf16 K-major to shared memory, then ctx[d, e] += P V^T as a third tcgen05 GEMM accumulating in TMEM over
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

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
    float2 t;
    t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
    t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
    t = ptx::unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
    t = ptx::unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}

// mean / rstd over the C channels of row r of a [128 x C] swizzled x tile (two passes over shared memory)
template <int C, uint32_t SPAN_STRIDE = SPAN_BYTES>
__device__ __forceinline__ void row_stats(const uint8_t* sx, int r, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int sp = 0; sp < C / 64; ++sp)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float v[8];
            unpack8(*reinterpret_cast<const uint4*>(sx + sp * SPAN_STRIDE + sw_off(r, c)), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j];
        }
    mean = s * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int sp = 0; sp < C / 64; ++sp)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float v[8];
            unpack8(*reinterpret_cast<const uint4*>(sx + sp * SPAN_STRIDE + sw_off(r, c)), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss = fmaf(v[j] - mean, v[j] - mean, ss);
        }
    rstd = rsqrtf(ss * (1.0f / C) + eps);
}

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
    const float* rowsum;     // [384]
    const float* kshift;     // [128]
    float* ctx_part;         // [B * parts][128 (h*32+d)][32 (e)]
    float* s_part;           // [B * parts][128]
    float eps;
};

constexpr int KV_PX = 64;                // pixels per kv tile: K^T / V^T accumulators are 64 columns each, so the CTA needs
                                         // 256 TMEM columns and two CTAs share an SM (their serial MMA <-> epilogue chains
                                         // interleave; one 128-pixel CTA per SM left the tensor pipe and the SFUs idle half the time)
template <int C>
struct KvCfg {
    static constexpr int SPANS = C / 64;
    static constexpr uint32_t W_BYTES = SPANS * SPAN_BYTES;          // one of Wk' / Wv' [128 rows][C]
    static constexpr uint32_t XSPAN_BYTES = KV_PX * 128;             // one 64-channel span of a 64-pixel x tile
    static constexpr uint32_t X_BYTES = SPANS * XSPAN_BYTES;
    static constexpr uint32_t PV_BYTES = SPAN_BYTES;                 // [128 rows][64 px]
    static constexpr uint32_t SMALL_BYTES = 3 * 256 + 5 * 512 + 128;
    static constexpr int SMEM_BYTES = 2 * W_BYTES + 2 * X_BYTES + 2 * PV_BYTES + SMALL_BYTES + 1024;
    static constexpr int CTAS_PER_SM = C == 64 ? 2 : 1;
};
constexpr int KV_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue: (TMEM lane quarter, 32-pixel half)

template <int C>
__global__ void __launch_bounds__(KV_THREADS, KvCfg<C>::CTAS_PER_SM)
linattn_kv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const KvArgs a) {
    using Cf = KvCfg<C>;
    constexpr int SPANS = Cf::SPANS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    uint8_t* sWk = smem;
    uint8_t* sWv = sWk + Cf::W_BYTES;
    uint8_t* sX = sWv + Cf::W_BYTES;                 // 2 stages
    uint8_t* sP = sX + 2 * Cf::X_BYTES;
    uint8_t* sV = sP + Cf::PV_BYTES;
    float* s_mu = reinterpret_cast<float*>(sV + Cf::PV_BYTES);   // [64] rstd * mean per pixel
    float* s_rstd = s_mu + KV_PX;                    // [64]
    float* s_rl = s_rstd + KV_PX;                    // [64] rstd * log2(e): the K path's multiplier, so an exponent is two FMAs
    float* s_sk = s_rl + KV_PX;                      // [128]
    float* s_sv = s_sk + 128;
    float* s_shift = s_sv + 128;
    float* s_S = s_shift + 128;                      // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_S + 256);
    uint64_t* w_bar = bars;
    uint64_t* x_full = bars + 1;                     // [2]
    uint64_t* x_empty = bars + 3;                    // [2]
    uint64_t* d_full = bars + 5;
    uint64_t* pv_ready = bars + 6;
    uint64_t* ctx_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        ptx::mbar_init(w_bar, 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&x_full[i], 1); ptx::mbar_init(&x_empty[i], 1); }
        ptx::mbar_init(d_full, 1);
        ptx::mbar_init(pv_ready, 256);
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
            const int st = t & 1;
            ptx::mbar_wait(&x_empty[st], ((t >> 1) & 1u) ^ 1u);
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
            if (t < T) ptx::mbar_wait(&x_full[t & 1], (t >> 1) & 1u);
            if (t > 0) ptx::mbar_wait(pv_ready, (t - 1) & 1u);     // epilogue t-1: D_K / D_V drained, P / V written
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                if (t > 0) {                                        // ctx[d, e] += P V^T over the 64 pixels of tile t-1
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16(tmem_base + COL_CTX, dP + 2u * k, dV + 2u * k, idesc_ctx, (t > 1 || k != 0) ? 1u : 0u);
                }
                if (t < T) {                                        // K^T = Wk' x^T, V^T = Wv' x^T of tile t
                    const uint64_t xoff = static_cast<uint64_t>(((t & 1) * Cf::X_BYTES) >> 4);
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
        if (te < 128) {
            s_sk[te] = a.rowsum[HD + te];
            s_sv[te] = a.rowsum[2 * HD + te];
            s_shift[te] = a.kshift[te];
        }
        named_bar_sync(1, 256);
        constexpr float LOG2E = 1.4426950408889634f;
        const float sk_r = s_sk[r] * LOG2E, sv_r = s_sv[r], shift_r = s_shift[r] * LOG2E;   // exp(k) = exp2(k * log2 e)
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int px0 = hcol * 32;
        float ssum = 0.f;
        for (int t = 0; t < T; ++t) {
            const int st = t & 1;
            ptx::mbar_wait(&x_full[st], (t >> 1) & 1u);          // x tile visible to this thread
            if (te < KV_PX) {                                     // per-pixel LayerNorm statistics while the MMAs run
                float mean, rstd;
                row_stats<C, Cf::XSPAN_BYTES>(sX + st * Cf::X_BYTES, te, a.eps, mean, rstd);
                s_mu[te] = rstd * mean;          // W LN(x) = rstd * (W'x) - (rstd * mean) * rowsum(W')
                s_rstd[te] = rstd;
                s_rl[te] = rstd * 1.4426950408889634f;
            }
            ptx::mbar_wait(d_full, t & 1u);                       // K^T / V^T of tile t ready (and P / V of tile t-1 consumed)
            ptx::tc_fence_after();
            named_bar_sync(1, 256);
            if (te == 0) ptx::mbar_arrive(&x_empty[st]);          // MMAs done (d_full) and statistics read: stage free
            uint32_t v[32];
            ptx::tmem_ld32(tlane + COL_K + px0, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 ma = *reinterpret_cast<const float4*>(s_mu + px0 + jj * 8);
                const float4 mb = *reinterpret_cast<const float4*>(s_mu + px0 + jj * 8 + 4);
                const float4 ra = *reinterpret_cast<const float4*>(s_rl + px0 + jj * 8);
                const float4 rb = *reinterpret_cast<const float4*>(s_rl + px0 + jj * 8 + 4);
                const float mu[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
                const float rs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};      // rstd * log2(e)
                uint32_t pk[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const int j = e2 * 2;
                    const float p0 = ptx::ex2(fmaf(rs[j], __uint_as_float(v[jj * 8 + j]), -fmaf(mu[j], sk_r, shift_r)));
                    const float p1 = ptx::ex2(fmaf(rs[j + 1], __uint_as_float(v[jj * 8 + j + 1]), -fmaf(mu[j + 1], sk_r, shift_r)));
                    // the denominator sums the unrounded p: the bf16 rounding of the MMA operand is unbiased, over n >= 1024 pixels
                    // the two sums agree to ~1e-5 (the unpack round trip was 7 % of the kernel's instructions)
                    ssum += p0 + p1;
                    pk[e2] = ptx::pack_bf16x2(p0, p1);
                }
                *reinterpret_cast<uint4*>(sP + sw_off(r, hcol * 4 + jj)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            ptx::tmem_ld32(tlane + COL_V + px0, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 ma = *reinterpret_cast<const float4*>(s_mu + px0 + jj * 8);
                const float4 mb = *reinterpret_cast<const float4*>(s_mu + px0 + jj * 8 + 4);
                const float4 ra = *reinterpret_cast<const float4*>(s_rstd + px0 + jj * 8);
                const float4 rb = *reinterpret_cast<const float4*>(s_rstd + px0 + jj * 8 + 4);
                const float mu[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
                const float rs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                uint32_t pk[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const int j = e2 * 2;
                    const float v0 = fmaf(rs[j], __uint_as_float(v[jj * 8 + j]), -mu[j] * sv_r);
                    const float v1 = fmaf(rs[j + 1], __uint_as_float(v[jj * 8 + j + 1]), -mu[j + 1] * sv_r);
                    pk[e2] = ptx::pack_bf16x2(v0, v1);
                }
                *reinterpret_cast<uint4*>(sV + sw_off(r, hcol * 4 + jj)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(pv_ready);
        }
        s_S[hcol * 128 + r] = ssum;
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

// ------------------------------------------------------------------------------------------------ kernel 2: mix
__global__ void __launch_bounds__(128)
linattn_mix_kernel(const float* __restrict__ ctx_part, const float* __restrict__ s_part, const float* __restrict__ wo,
                   bf16* __restrict__ mb, int parts, int C, float inv_n_scale) {
    __shared__ float s_ctx[HD][33];
    __shared__ __align__(16) float s_wo[32 * HD];   // this block's rows of Wo (C / 4 <= 32 output channels)
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    const int b = blockIdx.x;
    const int t = threadIdx.x;          // hd = h*32 + d
    const int h = t >> 5;
    const int per = C / gridDim.y;      // output channels of this block
    {   // coalesced copy of Wo[blockIdx.y * per .. +per][128]; its latency overlaps the partial loads below
        const float4* src = reinterpret_cast<const float4*>(wo + static_cast<size_t>(blockIdx.y) * per * HD);
        for (int i = t; i < per * HD / 4; i += 128) reinterpret_cast<float4*>(s_wo)[i] = __ldg(src + i);
    }
    // all partial loads of this thread (<= 8 parts x (S + eight 16-byte ctx chunks)) are issued before any is consumed
    float S = 0.f;
    float4 acc4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float sp[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) sp[p] = p < parts ? __ldg(s_part + (static_cast<size_t>(b) * parts + p) * HD + t) : 0.f;
#pragma unroll
    for (int p0 = 0; p0 < 8; p0 += 2) {
        float4 v[2][8];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const int p = p0 + pp;
            const float4* src = reinterpret_cast<const float4*>(ctx_part + ((static_cast<size_t>(b) * parts + (p < parts ? p : 0)) * HD + t) * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[pp][j] = p < parts ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc4[j].x += v[pp][j].x; acc4[j].y += v[pp][j].y; acc4[j].z += v[pp][j].z; acc4[j].w += v[pp][j].w;
            }
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) S += sp[p];
    const float norm = inv_n_scale / S;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s_ctx[t][4 * j] = acc4[j].x * norm; s_ctx[t][4 * j + 1] = acc4[j].y * norm;
        s_ctx[t][4 * j + 2] = acc4[j].z * norm; s_ctx[t][4 * j + 3] = acc4[j].w * norm;
    }
    __syncthreads();
    float c[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) c[e] = s_ctx[t][e];
    for (int cl = 0; cl < per; ++cl) {
        const int co = blockIdx.y * per + cl;
        const float4* w = reinterpret_cast<const float4*>(s_wo + cl * HD + h * 32);   // broadcast reads
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float4 w4 = w[e];
            acc = fmaf(w4.x, c[4 * e], acc); acc = fmaf(w4.y, c[4 * e + 1], acc);
            acc = fmaf(w4.z, c[4 * e + 2], acc); acc = fmaf(w4.w, c[4 * e + 3], acc);
        }
        mb[(static_cast<size_t>(b) * C + co) * HD + t] = __float2bfloat16(acc);
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
    static constexpr uint32_t SMALL_BYTES = 512 + 2 * 4 * C + 128 + 1024;    // rowsum, bo, g2, barriers, [2][128] half-row exchange
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
    float* s_xch = s_g2 + C;                         // [2][128]: the two column halves of a row exchange their LayerNorm partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xch + 256);
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
        // of the output channels (LayerNorm); both halves recompute the cheap per-row statistics instead of exchanging them
        const int q = warp & 3;
        const int hf = (warp - 2) >> 2;
        const int te = (warp - 2) * 32 + lane;          // 0..255
        const int r = q * 32 + lane;
        if (te < 128) s_sq[te] = a.rowsum[te];
        if (te < C) { s_bo[te] = a.bo[te]; s_g2[te] = a.g2[te]; }
        named_bar_sync(1, 256);
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int it = 0; it < ntile; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int st = it & 1;
            uint8_t* xs = sX + st * Cf::X_BYTES;
            ptx::mbar_wait(&x_full[st], (it >> 1) & 1u);
            float mean, rstd;
            row_stats<C>(xs, r, a.eps, mean, rstd);
            ptx::mbar_wait(q_full, it & 1u);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                const int h = 2 * hf + hh;
                uint32_t v[32];
                ptx::tmem_ld32(tlane + COL_Q + h * 32, v);
                ptx::tmem_ld_wait();
                float f[32];
                float mx = -INFINITY;
                // q = rstd * (acc - mean * rowsum) and exp(q - max) = exp2(q * log2e - max'): one FMA per element with
                // al = rstd * log2e, cl = mean * al (softmax is invariant to the common shift, the max is taken after scaling)
                const float al = rstd * 1.4426950408889634f, cl = mean * al;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 s4 = *reinterpret_cast<const float4*>(s_sq + h * 32 + j);
                    f[j] = fmaf(__uint_as_float(v[j]), al, -(cl * s4.x));
                    f[j + 1] = fmaf(__uint_as_float(v[j + 1]), al, -(cl * s4.y));
                    f[j + 2] = fmaf(__uint_as_float(v[j + 2]), al, -(cl * s4.z));
                    f[j + 3] = fmaf(__uint_as_float(v[j + 3]), al, -(cl * s4.w));
                    mx = fmaxf(fmaxf(fmaxf(mx, f[j]), fmaxf(f[j + 1], f[j + 2])), f[j + 3]);
                }
                float ssum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) { f[j] = ptx::ex2(f[j] - mx); ssum += f[j]; }
                const float inv = 1.0f / ssum;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint4 o;
                    o.x = ptx::pack_bf16x2(f[jj * 8] * inv, f[jj * 8 + 1] * inv);
                    o.y = ptx::pack_bf16x2(f[jj * 8 + 2] * inv, f[jj * 8 + 3] * inv);
                    o.z = ptx::pack_bf16x2(f[jj * 8 + 4] * inv, f[jj * 8 + 5] * inv);
                    o.w = ptx::pack_bf16x2(f[jj * 8 + 6] * inv, f[jj * 8 + 7] * inv);
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
            constexpr int HALF = C / 2;
            float y[HALF];
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < HALF / 32; ++cc) {
                const int c32 = hf * (HALF / 32) + cc;
                uint32_t v[32];
                ptx::tmem_ld32(tlane + COL_Y + c32 * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    y[cc * 32 + j] = __uint_as_float(v[j]) + s_bo[c32 * 32 + j];
                    sum += y[cc * 32 + j];
                }
            }
            s_xch[hf * 128 + r] = sum;
            named_bar_sync(2 + q, 64);
            const float ymean = (sum + s_xch[(hf ^ 1) * 128 + r]) * (1.0f / C);
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < HALF; ++j) {
                y[j] -= ymean;
                ss = fmaf(y[j], y[j], ss);
            }
            named_bar_sync(2 + q, 64);                   // the partner has read this row's sum: the slot may be reused
            s_xch[hf * 128 + r] = ss;
            named_bar_sync(2 + q, 64);
            const float yrstd = rsqrtf((ss + s_xch[(hf ^ 1) * 128 + r]) * (1.0f / C) + a.eps);
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
                    float xr[8];
                    unpack8(*px, xr);
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c32 * 32 + jj * 8 + j;
                        o[j] = fmaf(y[cc * 32 + jj * 8 + j] * yrstd, s_g2[c], xr[j]);
                    }
                    uint4 u;
                    u.x = ptx::pack_bf16x2(o[0], o[1]);
                    u.y = ptx::pack_bf16x2(o[2], o[3]);
                    u.z = ptx::pack_bf16x2(o[4], o[5]);
                    u.w = ptx::pack_bf16x2(o[6], o[7]);
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
        e = cudaFuncSetAttribute(linattn_out_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, OutCfg<C>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const LinAttnFusedDesc& d = l.d;
    KvArgs ka;
    ka.n = d.n; ka.tiles_per_unit = l.tiles_per_unit * (TILE / KV_PX); ka.parts = l.parts; ka.rowsum = d.rowsum; ka.kshift = d.kshift;
    ka.ctx_part = d.ctx_part; ka.s_part = d.s_part; ka.eps = d.eps;
    cudaError_t e = launch_pdl(linattn_kv_kernel<C>, dim3(d.B * l.parts), dim3(KV_THREADS), KvCfg<C>::SMEM_BYTES, s, l.tmXk, l.tmW, ka);
    if (e != cudaSuccess) return e;
    e = launch_pdl(linattn_mix_kernel, dim3(d.B, 4), dim3(128), 0, s, d.ctx_part, d.s_part, d.wo, d.mb, l.parts, C,
                   0.17677669529663687f / static_cast<float>(d.n));   // 32^-0.5 / n
    if (e != cudaSuccess) return e;
    OutArgs oa;
    oa.n = d.n; oa.tiles_per_img = d.n / TILE; oa.num_tiles = l.num_tiles; oa.rowsum = d.rowsum; oa.bo = d.bo; oa.g2 = d.g2;
    oa.eps = d.eps;
    return launch_pdl(linattn_out_kernel<C>, dim3(l.out_grid), dim3(OUT_THREADS), OutCfg<C>::SMEM_BYTES, s, l.tmX, l.tmW, l.tmM, l.tmY, oa);
}

cudaError_t linattn_fused_run(const LinAttnFusedLaunch& l, cudaStream_t s) {
    return l.d.C == 64 ? run_c<64>(l, s) : run_c<128>(l, s);
}

}  // namespace hd
