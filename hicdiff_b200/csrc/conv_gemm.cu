// Implicit-GEMM convolution for the HiCDiff eps-predictors on sm_100a.
//
// Replaces, on the sampling path, every F.conv2d the reference issues with >= 64 input channels:
//   WeightStandardizedConv2d / Block.proj      /root/reference/src/hicdiff_condition.py:84-97,158
//   to_qkv / to_out / res_conv / final 3x3     :183,205,208,236,237,320,336
//   Downsample (pixel-unshuffle + 1x1)         :78-82
//   Upsample's 3x3 (input already upsampled)   :72-76
//   hicedrn_Diff Block.proj / body_tail / tail /root/reference/src/model/hicedrn_Diff.py:169-180,259,263
//
// GEMM view: D[M = B*H*W pixels, N = Cout] = A[M, K] * W[N, K]^T with K = taps * Cin.
//   * A is never materialised: one K block = (filter tap, 64-channel chunk) and is fetched by ONE TMA box
//     load {64 ch, W, rows, imgs} from the NHWC activation at spatial offset (dy, dx); out-of-bounds rows /
//     columns are zero-filled by the TMA unit, which is exactly the conv's zero padding.
//   * A channel concat (torch.cat((x, skip), 1)) is two tensor maps walked back to back inside each tap.
//   * The pixel-unshuffle of Downsample is a 5-D view [C, p2, W/2, p1, B*H/2] of the same NHWC buffer.
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) accumulates into TMEM; two accumulator
//     stages so the epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, one per SM.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> regs -> global).
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace hd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // 64 bf16 = one 128-byte swizzle span
constexpr uint32_t A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BUDGET = 192 * 1024;     // operand ring; + 32 KiB of epilogue staging stays under the 227 KiB limit
constexpr int EPI_WARPS = 4;
constexpr uint32_t EPI_BUF_BYTES = 32 * 128;                      // 32 rows x 64 bf16, 128B-swizzled TMA-store box
constexpr uint32_t EPI_STAGE_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;   // two buffers per epilogue warp

template <int BN>
struct TileCfg {
    static constexpr uint32_t B_STAGE_BYTES = BN * BLOCK_K * 2;
    static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                          : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct KArgs {
    int M, N, num_m_tiles, num_n_tiles, nkb, chunks0, chunks1, mode, W, P, kw, pad;
    ConvEpilogue epi;
    bf16* out;
    int ldo;
};

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD, const KArgs a) {
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = sB + STAGES * Cfg::B_STAGE_BYTES;      // 1024-aligned: all stage sizes are multiples of 1 KiB
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + EPI_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            ptx::prefetch_tmap(&tmA0);
            ptx::prefetch_tmap(&tmA1);
            ptx::prefetch_tmap(&tmB);
            ptx::prefetch_tmap(&tmD);
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 128);
        }
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = a.num_m_tiles * a.num_n_tiles;
    const int chunks = a.chunks0 + a.chunks1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / a.num_n_tiles;
                const int nt = tile - mt * a.num_n_tiles;
                const int m0 = mt * BLOCK_M;
                const int n0 = nt * BN;
                const int b0 = m0 / a.P;
                const int h0 = (m0 - b0 * a.P) / a.W;
                const int row0 = m0 / a.W;  // merged (b, h) row for the unshuffle view
                int tap = 0, chunk = 0;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    const bool second = chunk >= a.chunks0;
                    const CUtensorMap* tm = second ? &tmA1 : &tmA0;
                    const int c0 = (second ? chunk - a.chunks0 : chunk) * BLOCK_K;
                    uint8_t* dstA = sA + stage * A_STAGE_BYTES;
                    if (a.mode == CONV_TAPS) {
                        const int dy = tap / a.kw - a.pad;
                        const int dx = tap - (tap / a.kw) * a.kw - a.pad;
                        ptx::tma_load_4d(dstA, tm, &full_bar[stage], c0, dx, h0 + dy, b0);
                    } else {
                        // tap = p1 * 2 + p2
                        ptx::tma_load_5d(dstA, tm, &full_bar[stage], c0, tap & 1, 0, tap >> 1, row0);
                    }
                    ptx::tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, n0);
                    if (++chunk == chunks) { chunk = 0; ++tap; }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1u;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < a.nkb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
                    const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle span: +2 in the >>4 address field
                        ptx::umma_bf16(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (kb == a.nkb - 1) ptx::umma_commit(&tfull_bar[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps)
        // Each warp owns 32 accumulator rows (its TMEM lane quarter).  bf16 output goes through a per-warp, 128B-swizzled
        // staging buffer and leaves as TMA tensor stores of 32 x 64 boxes (full-line writes, rows past M are clipped
        // by the TMA unit); the fp32 head path (N padded to 16, one valid column) stores directly.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;            // accumulator row == pixel within the tile
        const ConvEpilogue& e = a.epi;
        uint8_t* my_stage = sEpi + q * (2 * EPI_BUF_BYTES);
        int buf = 0;
        int iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
            const int mt = tile / a.num_n_tiles;
            const int nt = tile - mt * a.num_n_tiles;
            const int m = mt * BLOCK_M + r;
            const int n0 = nt * BN;
            const bool valid = m < a.M;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1u;

            const float* frow = nullptr;
            if (e.film != nullptr) {
                const int b = valid ? m / a.P : 0;
                const int row = e.film_row[b * e.film_row_stride];
                frow = e.film + static_cast<size_t>(row) * e.film_ld + e.film_off;
            }
            const int shift_off = e.film_has_scale ? a.N : 0;

            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

            auto finish = [&](float (&f)[32], int n) {      // bias / FiLM / SiLU / scale / residual on 32 columns
                if (e.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
                        f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
                    }
                }
                if (frow != nullptr) {
                    if (e.film_has_scale) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(frow + n + j));
                            f[j] *= (sc.x + 1.0f); f[j + 1] *= (sc.y + 1.0f);
                            f[j + 2] *= (sc.z + 1.0f); f[j + 3] *= (sc.w + 1.0f);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 sh = __ldg(reinterpret_cast<const float4*>(frow + shift_off + n + j));
                        f[j] += sh.x; f[j + 1] += sh.y; f[j + 2] += sh.z; f[j + 3] += sh.w;
                    }
                }
                if (e.silu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
                }
                if (e.out_scale != 1.0f) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] *= e.out_scale;
                }
                if (e.res != nullptr && valid) {
                    const uint4* rp = reinterpret_cast<const uint4*>(e.res + static_cast<size_t>(m) * e.ldr + n);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 rr = __ldg(rp + j / 8);
                        float2 t;
                        t = ptx::unpack_bf16x2(rr.x); f[j] += t.x; f[j + 1] += t.y;
                        t = ptx::unpack_bf16x2(rr.y); f[j + 2] += t.x; f[j + 3] += t.y;
                        t = ptx::unpack_bf16x2(rr.z); f[j + 4] += t.x; f[j + 5] += t.y;
                        t = ptx::unpack_bf16x2(rr.w); f[j + 6] += t.x; f[j + 7] += t.y;
                    }
                }
            };

            if constexpr (BN < 64) {
                // fp32 head (tail conv of HiCEDRN): 16 accumulator columns, n_valid of them real
                uint32_t v[16];
                ptx::tmem_ld16(taddr, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[as]);
                if (valid && e.out_f32 != nullptr) {
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < e.n_valid) {
                            float val = __uint_as_float(v[j]) + (e.bias != nullptr ? __ldg(e.bias + n0 + j) : 0.f);
                            if (e.res != nullptr) val += __bfloat162float(e.res[static_cast<size_t>(m) * e.ldr + n0 + j]);
                            e.out_f32[static_cast<size_t>(m) * e.n_valid + n0 + j] = val;
                        }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN; c += 64) {
                    uint32_t v0[32], v1[32];
                    ptx::tmem_ld32(taddr + c, v0);
                    ptx::tmem_ld32(taddr + c + 32, v1);
                    ptx::tmem_ld_wait();
                    if (c + 64 == BN) {          // accumulator fully drained: hand the TMEM stage back to the MMA warp
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&tempty_bar[as]);
                    }
                    float f0[32], f1[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { f0[j] = __uint_as_float(v0[j]); f1[j] = __uint_as_float(v1[j]); }
                    finish(f0, n0 + c);
                    finish(f1, n0 + c + 32);
                    if (e.gn_part != nullptr) {
                        // GroupNorm partials of this warp's 32 rows x 64 columns (rows are all valid: M % 32 == 0).
                        const int lg = 31 - __clz(a.N >> 6);            // log2(pieces of 8 columns per group)
                        float ps[8];
#pragma unroll
                        for (int p8 = 0; p8 < 8; ++p8) {
                            float t = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) t += (p8 < 4) ? f0[8 * p8 + j] : f1[8 * (p8 - 4) + j];
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                            ps[p8] = t;
                        }
                        const float inv_cnt = 1.0f / (32.0f * static_cast<float>(a.N >> 3));
                        float gsum[8], pmean[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float t = 0.f;
#pragma unroll
                            for (int p8 = 0; p8 < 8; ++p8) t += ((p8 >> lg) == k) ? ps[p8] : 0.f;
                            gsum[k] = t;
                        }
#pragma unroll
                        for (int p8 = 0; p8 < 8; ++p8) {
                            float t = 0.f;
#pragma unroll
                            for (int k = 0; k < 8; ++k) t = ((p8 >> lg) == k) ? gsum[k] * inv_cnt : t;
                            pmean[p8] = t;
                        }
#pragma unroll
                        for (int p8 = 0; p8 < 8; ++p8) {
                            float t = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float dlt = ((p8 < 4) ? f0[8 * p8 + j] : f1[8 * (p8 - 4) + j]) - pmean[p8];
                                t = fmaf(dlt, dlt, t);
                            }
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                            ps[p8] = t;
                        }
                        if (lane == 0 && valid) {
                            const int ngroups = 8 >> lg;                         // groups inside this 64-column chunk
                            const int g0 = (n0 + c) / (a.N >> 3);                // first global group of the chunk
                            float2* dst = e.gn_part + (static_cast<size_t>(mt) * 4 + q) * 8 + g0;
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                float m2 = 0.f;
#pragma unroll
                                for (int p8 = 0; p8 < 8; ++p8) m2 += ((p8 >> lg) == k) ? ps[p8] : 0.f;
                                if (k < ngroups) dst[k] = make_float2(gsum[k], m2);
                            }
                        }
                    }
                    // staging buffer free?  (at most one older store of this warp may still be reading the OTHER buffer)
                    if (lane == 0) ptx::bulk_wait_read<1>();
                    __syncwarp();
                    uint8_t* stage_row = my_stage + buf * EPI_BUF_BYTES + lane * 128;
                    const int sw = lane & 7;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = ptx::pack_bf16x2(f0[8 * j], f0[8 * j + 1]);
                        o.y = ptx::pack_bf16x2(f0[8 * j + 2], f0[8 * j + 3]);
                        o.z = ptx::pack_bf16x2(f0[8 * j + 4], f0[8 * j + 5]);
                        o.w = ptx::pack_bf16x2(f0[8 * j + 6], f0[8 * j + 7]);
                        *reinterpret_cast<uint4*>(stage_row + ((j ^ sw) << 4)) = o;
                        uint4 p;
                        p.x = ptx::pack_bf16x2(f1[8 * j], f1[8 * j + 1]);
                        p.y = ptx::pack_bf16x2(f1[8 * j + 2], f1[8 * j + 3]);
                        p.z = ptx::pack_bf16x2(f1[8 * j + 4], f1[8 * j + 5]);
                        p.w = ptx::pack_bf16x2(f1[8 * j + 6], f1[8 * j + 7]);
                        *reinterpret_cast<uint4*>(stage_row + (((j + 4) ^ sw) << 4)) = p;
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tmD, my_stage + buf * EPI_BUF_BYTES, n0 + c, mt * BLOCK_M + q * 32);
                        ptx::bulk_commit();
                    }
                    buf ^= 1;
                }
            }
        }
        if (lane == 0) ptx::bulk_wait_all();    // all tensor stores of this warp have landed before the CTA retires
        __syncwarp();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_map(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, char* err, int errlen) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available");
        return 1;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu/%llu/%llu)", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull);
        return 1;
    }
    return 0;
}

int encode_activation_map(CUtensorMap* tm, const ConvSrc& s, const ConvGemmDesc& d, char* err, int errlen) {
    const cuuint64_t C = s.C;
    if (d.mode == CONV_TAPS) {
        const int P = d.H * d.W;
        const int rows = P >= BLOCK_M ? BLOCK_M / d.W : d.H;
        const int imgs = P >= BLOCK_M ? 1 : BLOCK_M / P;
        cuuint64_t dims[4] = {C, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t str[3] = {C * 2, C * 2 * d.W, C * 2 * d.W * d.H};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)d.W, (cuuint32_t)rows, (cuuint32_t)imgs};
        return encode_map(tm, s.ptr, 4, dims, str, box, err, errlen);
    }
    // input is [B, 2H, 2W, C]; view as [C, p2, W, p1, B*H]
    cuuint64_t dims[5] = {C, 2, (cuuint64_t)d.W, 2, (cuuint64_t)d.B * d.H};
    cuuint64_t str[4] = {C * 2, C * 2 * 2, C * 2 * 2 * d.W, C * 2 * 2 * d.W * 2};
    cuuint32_t box[5] = {BLOCK_K, 1, (cuuint32_t)d.W, 1, (cuuint32_t)(BLOCK_M / d.W)};
    return encode_map(tm, s.ptr, 5, dims, str, box, err, errlen);
}

template <int BN>
cudaError_t launch_bn(const ConvGemmLaunch& l, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TileCfg<BN>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    KArgs k;
    k.M = l.M; k.N = l.N; k.num_m_tiles = l.num_m_tiles; k.num_n_tiles = l.num_n_tiles; k.nkb = l.nkb;
    k.chunks0 = l.chunks0; k.chunks1 = l.chunks1; k.mode = l.mode; k.W = l.W; k.P = l.P; k.kw = l.kw; k.pad = l.pad;
    k.epi = l.epi; k.out = l.out; k.ldo = l.ldo;
    conv_gemm_kernel<BN><<<l.grid, NUM_THREADS, TileCfg<BN>::SMEM_BYTES, s>>>(l.tmA0, l.tmA1, l.tmB, l.tmD, k);
    return cudaGetLastError();
}

}  // namespace

int conv_gemm_prepare(const ConvGemmDesc& d, int num_sms, ConvGemmLaunch* out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
    const int P = d.H * d.W;
    if (d.src0.ptr == nullptr || d.src0.C % BLOCK_K != 0 || (d.src1.ptr != nullptr && d.src1.C % BLOCK_K != 0)) {
        snprintf(err, errlen, "conv_gemm: input channels must be multiples of %d (got %d/%d)", BLOCK_K, d.src0.C,
                 d.src1.ptr ? d.src1.C : 0);
        return 1;
    }
    if (d.W > BLOCK_M || BLOCK_M % d.W != 0 || (P < BLOCK_M && BLOCK_M % P != 0) || (P >= BLOCK_M && P % BLOCK_M != 0)) {
        snprintf(err, errlen, "conv_gemm: unsupported spatial size %dx%d for a %d-pixel tile", d.H, d.W, BLOCK_M);
        return 1;
    }
    if (d.mode == CONV_TAPS && d.ksize != 1 && d.ksize != 3) {
        snprintf(err, errlen, "conv_gemm: kernel size %d unsupported", d.ksize);
        return 1;
    }
    int bn;
    if (d.N % 256 == 0) bn = 256;
    else if (d.N % 128 == 0) bn = 128;
    else if (d.N % 64 == 0) bn = 64;
    else if (d.N % 16 == 0 && d.N <= 48) bn = 16;
    else {
        snprintf(err, errlen, "conv_gemm: output channels %d unsupported", d.N);
        return 1;
    }
    const int M = d.B * P;
    out->bn = bn;
    out->M = M;
    out->N = d.N;
    out->num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    out->num_n_tiles = d.N / bn;
    // Few tiles relative to the machine: prefer a narrower N tile so more SMs get work.
    while (bn > 64 && out->num_m_tiles * out->num_n_tiles < num_sms && d.N % (bn / 2) == 0) {
        bn /= 2;
        out->bn = bn;
        out->num_n_tiles = d.N / bn;
    }
    out->chunks0 = d.src0.C / BLOCK_K;
    out->chunks1 = d.src1.ptr ? d.src1.C / BLOCK_K : 0;
    const int taps = d.mode == CONV_TAPS ? d.ksize * d.ksize : 4;
    out->nkb = taps * (out->chunks0 + out->chunks1);
    out->mode = d.mode;
    out->W = d.W;
    out->P = P;
    out->kw = d.mode == CONV_TAPS ? d.ksize : 2;
    out->pad = d.mode == CONV_TAPS ? d.ksize / 2 : 0;
    out->epi = d.epi;
    out->out = d.out;
    out->ldo = d.N;
    const int tiles = out->num_m_tiles * out->num_n_tiles;
    out->grid = tiles < num_sms ? tiles : num_sms;

    if (encode_activation_map(&out->tmA0, d.src0, d, err, errlen)) return 1;
    if (d.src1.ptr) {
        if (encode_activation_map(&out->tmA1, d.src1, d, err, errlen)) return 1;
    } else {
        out->tmA1 = out->tmA0;
    }
    const cuuint64_t Ktot = (cuuint64_t)out->nkb * BLOCK_K;
    cuuint64_t wd[2] = {Ktot, (cuuint64_t)d.N};
    cuuint64_t ws[1] = {Ktot * 2};
    cuuint32_t wb[2] = {BLOCK_K, (cuuint32_t)bn};
    if (encode_map(&out->tmB, d.weight, 2, wd, ws, wb, err, errlen)) return 1;
    if (d.epi.out_f32 == nullptr) {
        if (bn < 64 || d.out == nullptr) {
            snprintf(err, errlen, "conv_gemm: bf16 output needs an N tile >= 64 and an output buffer");
            return 1;
        }
        cuuint64_t od[2] = {(cuuint64_t)d.N, (cuuint64_t)M};
        cuuint64_t os[1] = {(cuuint64_t)d.N * 2};
        cuuint32_t ob[2] = {64, 32};
        if (encode_map(&out->tmD, d.out, 2, od, os, ob, err, errlen)) return 1;
    } else {
        if (bn >= 64) {
            snprintf(err, errlen, "conv_gemm: the fp32 head path is built for N <= 48 (padded to 16-column tiles)");
            return 1;
        }
        out->tmD = out->tmB;   // unused by the fp32 head path
    }
    switch (bn) {
        case 256: out->smem_bytes = TileCfg<256>::SMEM_BYTES; break;
        case 128: out->smem_bytes = TileCfg<128>::SMEM_BYTES; break;
        case 64: out->smem_bytes = TileCfg<64>::SMEM_BYTES; break;
        default: out->smem_bytes = TileCfg<16>::SMEM_BYTES; break;
    }
    return 0;
}

cudaError_t conv_gemm_run(const ConvGemmLaunch& l, cudaStream_t s) {
    switch (l.bn) {
        case 256: return launch_bn<256>(l, s);
        case 128: return launch_bn<128>(l, s);
        case 64: return launch_bn<64>(l, s);
        case 16: return launch_bn<16>(l, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hd
