// Implicit-GEMM convolution for the HiCDiff eps-predictors on sm_100a.
//
// Replaces, on the sampling path, every F.conv2d the reference issues with >= 64 input channels:
//   WeightStandardizedConv2d / Block.proj      /root/reference/src/hicdiff_condition.py:84-97,158
//   to_qkv / to_out / res_conv / final 3x3     :183,205,208,236,237,320,336
//   Downsample (pixel-unshuffle + 1x1)         :78-82
//   Upsample (nearest x2 + 3x3)                :72-76
//   hicedrn_Diff Block.proj / body_tail / tail /root/reference/src/model/hicedrn_Diff.py:169-180,259,263
//
// GEMM view: D[M = B*H*W pixels, N = Cout] = A[M, K] * W[N, K]^T with K = taps * Cin.  A is never materialised:
//   * GENERAL path: one K block = (filter tap, 64-channel chunk), fetched by ONE TMA box load {64 ch, W, rows, imgs}
//     from the NHWC activation at spatial offset (dy, dx); out-of-bounds rows / columns are zero-filled by the TMA
//     unit, which is exactly the conv's zero padding.
//   * SLAB path (3x3, W >= 16, N <= 128): one TMA box {64 ch, W, R + 2 rows} per (chunk, dx) serves the three dy
//     taps -- the A descriptor of tap dy starts (dy + 1) * W pixel rows (a multiple of 1 KiB, so swizzle atoms stay
//     aligned) into the slab.  One pipeline stage then carries 12 MMAs instead of 4, and the L2 -> SM traffic of the
//     A operand drops from 9 to 3 * (R + 2) / R loads per pixel.  When the whole weight matrix fits (N == 64,
//     K <= 1152: 72 / 144 KiB) it is loaded into shared memory ONCE per persistent CTA (SLAB_RES).
//   * DX-STACKED path (3x3, N == 64, resident weights; the default there): the measured issue law of tcgen05 on B200 is
//     N/2 + 43 clk per K = 16 MMA (scripts/ubench_tcgen05.cu), i.e. the fixed part is paid per MMA, not per column.  So the
//     three dx taps of a filter row are stacked along N: ONE un-shifted slab per chunk, B tile of 192 rows (3 x 64 output
//     channels), one MMA per (dy, K = 16) -- 139 clk instead of 3 x 75 -- and the epilogue adds the accumulator's three
//     64-column groups shifted by -1 / 0 / +1 pixel (whole image rows per tile, so the shifts stay inside it).  With the
//     MMAs that short the epilogue chain became the limiter; where shared memory allows the kernel runs 18 warps with two
//     epilogue groups on alternating tiles (K_DX3G), else 10 warps (K_DX3).  profiles/r02_notes.md 11.
//   * A channel concat (torch.cat((x, skip), 1)) is two tensor maps walked back to back inside each tap.
//   * Downsample's pixel-unshuffle is a 5-D view [C, p2, W/2, p1, B*H/2] of the same NHWC buffer.
//   * Upsample's nearest x2 is folded into the conv: output phase (a, b) = (y & 1, x & 1) is a 2x2 conv over the
//     LOW-resolution input with pre-summed weights (prep.cu), K = 4*Cin instead of 9*Cin, and the upsampled tensor
//     never exists; the epilogue scatters each phase through a 5-D tensor map of the output.
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) accumulates into TMEM; two accumulator stages so the
//     epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, one per SM.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> regs -> smem -> TMA
//     store).  Producer and issuer loops are warp-uniform with one elected lane issuing (ncu showed the earlier
//     `if (lane == 0)` form issue-bound: ~650 cycles of scalar code per K block, profiles/r01_conv_gemm_ncu.md).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace hd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                           // 64 bf16 = one 128-byte swizzle span
constexpr uint32_t A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr uint32_t SLAB_CAP_BYTES = 32 * 1024;        // (R + 2) * W pixels * 128 B: 32 / 24 / 20 KiB at W = 64 / 32 / 16
constexpr int NUM_THREADS = 320;                       // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARPS = 8;
constexpr uint32_t EPI_BUF_BYTES = 32 * 64;           // 32 rows x 32 bf16, 64B-swizzled TMA-store box
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = 256;

enum Kind : int { K_GENERAL = 0, K_SLAB = 1, K_SLAB_RES = 2, K_PAD = 3, K_DX3 = 4, K_DX3G = 5 };
constexpr int NUM_THREADS_2G = 576;                    // K_DX3G: warps 2..9 and 10..17 are two epilogue groups on alternating tiles
constexpr int DX3_N = 192;                             // K_DX3: MMA width = 3 dx taps x 64 output channels
constexpr uint32_t DX3_ACC_STRIDE = 256;               // TMEM columns between the two 192-column accumulator stages
constexpr uint32_t DX3_XCHG_BYTES = 1024;              // 4 warp pairs x 2 directions x 32 fp32: boundary rows between lane quarters
constexpr uint32_t PAD_SLAB_CAP_BYTES = 42 * 1024;     // (rows + 2) * (W + 2) pixels * 128 B: 42240 B at W = 64, 30464 B at W = 32

template <int BN, int KIND, bool GNF = false, int CG = 1>
struct Cfg {
    // one (N tile, K block) of weights; a CTA pair (CG == 2, tcgen05 cta_group::2) splits the N rows between its two CTAs
    static constexpr uint32_t B_BLOCK_BYTES = (BN / CG) * BLOCK_K * 2;
    static constexpr bool DX = KIND == K_DX3 || KIND == K_DX3G;      // dx-stacked kinds: MMAs of N = 192 over the un-shifted slab
    static_assert(!DX || (BN == 64 && !GNF && CG == 1), "the dx-stacked kinds are built for N == 64, plain epilogue, one CTA");
    static constexpr int THREADS = KIND == K_DX3G ? NUM_THREADS_2G : NUM_THREADS;
    static constexpr uint32_t A_STAGE_BYTES = KIND == K_GENERAL ? A_TILE_BYTES : (KIND == K_PAD ? PAD_SLAB_CAP_BYTES : SLAB_CAP_BYTES);
    static constexpr uint32_t B_STAGE_BYTES = KIND == K_GENERAL ? B_BLOCK_BYTES : (KIND == K_SLAB ? 3 * B_BLOCK_BYTES : 0);
    static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int EPI_BUFS = BN <= 64 ? 1 : 2;                             // staging buffers per epilogue warp
    // K_DX3G: 16 epilogue warps, each with one 32 rows x 32 columns TMA-store box (2 KiB, 64B swizzle) filled in two column halves
    static constexpr uint32_t EPI_BYTES = KIND == K_DX3G ? 16 * 2 * 1024 : EPI_WARPS * EPI_BUFS * EPI_BUF_BYTES;
    // accumulator ring in TMEM: 2 stages, or 4 when the epilogue also applies GroupNorm (its second phase trails by one
    // tile while it waits for the image's statistics)
    static constexpr int ACC_STAGES = GNF ? 4 : 2;
    static constexpr uint32_t ACC_STRIDE = DX ? DX3_ACC_STRIDE : BN;  // TMEM columns from one accumulator stage to the next
    static constexpr uint32_t ACC_COLS = ACC_STAGES * ACC_STRIDE;
    static constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    static constexpr uint32_t GN_AUX_BYTES = GNF ? (128 + 2 * 3 * BN * 4) : 0;    // (mean, rstd)[2][8] + (mul, add, post)[2][BN]
    static constexpr uint32_t XCHG_BYTES = KIND == K_DX3 ? DX3_XCHG_BYTES : (KIND == K_DX3G ? 2 * DX3_XCHG_BYTES : 0);
    static constexpr int AUX_BYTES = BAR_BYTES + static_cast<int>(GN_AUX_BYTES) + static_cast<int>(XCHG_BYTES);
    static int stages(uint32_t res_b_bytes) {
        const int avail = SMEM_LIMIT - 1024 - AUX_BYTES - static_cast<int>(EPI_BYTES) - static_cast<int>(res_b_bytes);
        int s = avail / static_cast<int>(STAGE_BYTES);
        return s > MAX_STAGES ? MAX_STAGES : s;
    }
    static int smem_bytes(int stages, uint32_t res_b_bytes) {
        return stages * STAGE_BYTES + res_b_bytes + EPI_BYTES + 1024 /*align*/ + AUX_BYTES;
    }
};

struct KArgs {
    int M, N, num_m_tiles, num_n_tiles, num_tiles, nkb, chunks0, chunks1, mode, W, P, kh, kw, pad, stages;
    int passes;                    // passes over the source chunks per tap: 1, or 2 with split weights (hi, then lo)
    int static_weights;            // 1: the weights are constants of the graph (their load may precede the PDL wait)
    int Wl_box, rows_box;          // CONV_UPSAMPLE store box: low-res pixels per row / rows per 32-pixel warp block
    uint32_t slab_bytes;           // bytes of one slab TMA box
    uint32_t slab_dy_bytes;        // W * 128: A-descriptor advance per dy tap
    uint32_t res_b_bytes;          // resident weight bytes (K_SLAB_RES, K_PAD)
    int PW, tiles_per_img, H;      // K_PAD: padded row pitch W + 2, 128-position tiles per image, image height
    int m_tiles_real;              // number of real 128-row M tiles (a CTA pair may own one phantom tile at the end)
    ConvEpilogue epi;
};

__device__ __forceinline__ float silu_f(float v) {      // h + h * tanh(h), h = v / 2: one SFU op (see norm.cu)
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

struct TileCoord { int mt, nt, phase; };

__device__ __forceinline__ TileCoord decode_tile(const KArgs& a, int tile) {
    TileCoord t;
    if (a.mode == CONV_UPSAMPLE) {
        const int units = a.num_n_tiles * 4;
        t.mt = tile / units;
        const int rem = tile - t.mt * units;
        t.phase = rem / a.num_n_tiles;
        t.nt = rem - t.phase * a.num_n_tiles;
    } else {
        t.mt = tile / a.num_n_tiles;
        t.nt = tile - t.mt * a.num_n_tiles;
        t.phase = 0;
    }
    return t;
}

// Warp-wide sums of 16 per-lane values with 16 shuffles: afterwards lane L holds the total of value (L >> 1).
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float keep = up ? v[i + 8] : v[i];
            const float send = up ? v[i] : v[i + 8];
            v[i] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float keep = up ? v[i + 4] : v[i];
            const float send = up ? v[i] : v[i + 4];
            v[i] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
    {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float keep = up ? v[i + 2] : v[i];
            const float send = up ? v[i] : v[i + 2];
            v[i] = keep + __shfl_xor_sync(FULL, send, 4);
        }
    }
    {
        const bool up = (lane & 2) != 0;
        const float keep = up ? v[1] : v[0];
        const float send = up ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    return v[0];
}

// Arrival counter protocol of the GroupNorm-fused epilogue.  Producer: partial stores, __syncwarp, then ONE release
// reduction (orders the warp's prior stores at L2 without the SC fence + L1 invalidate of __threadfence()).  Consumer: a
// relaxed (strong, L2) poll -- the data it guards is only ever read with ld.global.cg, i.e. from L2, the coherence
// point, and only after the poll has returned, so no L1 invalidation is needed on the acquire side either.
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_cg_f2(const float2* p) {     // L2-only load: data written by other SMs in this launch
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op (tanh.approx, rel. error ~2^-11, below the bf16 output rounding)
__device__ __forceinline__ float silu_tanh(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// Warp-wide sums of 8 per-lane values with 9 shuffles: afterwards lane L holds the total of value (L >> 2).
__device__ __forceinline__ float transpose_reduce8(float (&v)[8], int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float keep = up ? v[i + 4] : v[i];
            const float send = up ? v[i] : v[i + 4];
            v[i] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float keep = up ? v[i + 2] : v[i];
            const float send = up ? v[i] : v[i + 2];
            v[i] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
    {
        const bool up = (lane & 4) != 0;
        const float keep = up ? v[1] : v[0];
        const float send = up ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 2);
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    return v[0];
}

template <int BN, int KIND, bool GNF, int CG>
__global__ void __launch_bounds__((Cfg<BN, KIND, GNF, CG>::THREADS), 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                 const __grid_constant__ CUtensorMap tmD31, const __grid_constant__ CUtensorMap tmD30, const KArgs a) {
    using C = Cfg<BN, KIND, GNF, CG>;
    constexpr uint32_t B_BLOCK = C::B_BLOCK_BYTES;
    constexpr int ACC = C::ACC_STAGES;
    // CG == 2: two CTAs (a cluster) compute one 256-row tile pair with tcgen05 cta_group::2 -- each supplies its own 128 rows
    // of A and HALF of the N rows of B, so the per-SM shared-memory operand traffic of a K = 16 step drops from
    // 4096 + 32 N to 4096 + 16 N bytes (the measured limiter of the 1-CTA form).  The leader (rank 0) issues every MMA; TMA
    // completions of both CTAs land on the leader's barriers; commits are multicast to both.
    constexpr bool TWO = CG == 2;
    static_assert(!TWO || (!GNF && KIND != K_PAD && BN >= 64), "cta_group::2 is built for the plain general / slab kinds");
    const uint32_t rank = TWO ? ptx::cluster_ctarank() : 0u;
    const int tile0 = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_stride = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    const int stages = a.stages;
    uint8_t* sA = KIND == K_PAD ? smem + a.res_b_bytes : smem;
    uint8_t* sB = KIND == K_PAD ? smem : smem + stages * C::A_STAGE_BYTES;    // per-stage weights, or the resident matrix
    uint8_t* sEpi = smem + stages * C::A_STAGE_BYTES +
                    ((KIND == K_SLAB_RES || KIND == K_PAD || C::DX) ? a.res_b_bytes : stages * C::B_STAGE_BYTES);   // all sizes are KiB multiples
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + C::EPI_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 4;
    uint64_t* res_bar = tempty_bar + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
    float* s_gn = reinterpret_cast<float*>(sEpi + C::EPI_BYTES + BAR_BYTES);   // GNF: [2][16] (mean[8], rstd[8])
    float* s_ma = s_gn + 32;                                                   // GNF: [2][3 * BN] (mul, add, post)
    float* s_xchg = s_gn;                                                      // K_DX3 (never GNF): [4 pairs][2][32] boundary rows

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
        if constexpr (TWO) { ptx::tmem_alloc_cg2(tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish_cg2(); }
        else { ptx::tmem_alloc(tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish(); }
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], CG);          // CG == 2: the leader's expect_tx arrival + the peer's remote arrival
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < ACC; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 256 * CG);  // the leader's barrier collects the epilogue threads of both CTAs
        }
        ptx::mbar_init(res_bar, CG);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    if constexpr (TWO) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: from here on our successor in the stream may be scheduled (its CTAs take the SMs as ours retire and run their own
    // prologue meanwhile); each role below waits for the PREDECESSOR grid right before it first touches activations
    ptx::grid_dep_launch();

    // K-loop chunks per tap: one pass over the sources' 64-channel chunks, or two with split (hi + lo) weights -- the second
    // pass re-reads the same activations (source chunk = chunk % wrap) against the lo half of the tap's weight columns
    const int src_chunks = a.chunks0 + a.chunks1;
    const int chunks = src_chunks * a.passes;

    // Stage hand-off helpers (called by the elected producer lane).  CG == 2: both CTAs' loads complete on the LEADER's barrier.
    auto arm = [&](uint64_t* bar, uint32_t bytes) -> uint32_t {
        if constexpr (TWO) {
            const uint32_t lb = ptx::mapa(ptx::smem_u32(bar), 0);
            if (rank == 0) ptx::mbar_arrive_expect_tx(bar, 2 * bytes); else ptx::mbar_arrive_cluster(lb);
            return lb;
        } else {
            ptx::mbar_arrive_expect_tx(bar, bytes);
            return 0;
        }
    };
    auto ld2 = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, uint32_t lb, int c0, int c1) {
        if constexpr (TWO) ptx::tma_load_2d_cg2(dst, tm, lb, c0, c1); else ptx::tma_load_2d(dst, tm, bar, c0, c1);
    };
    auto ld4 = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, uint32_t lb, int c0, int c1, int c2, int c3) {
        if constexpr (TWO) ptx::tma_load_4d_cg2(dst, tm, lb, c0, c1, c2, c3); else ptx::tma_load_4d(dst, tm, bar, c0, c1, c2, c3);
    };
    auto ld5 = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, uint32_t lb, int c0, int c1, int c2, int c3, int c4) {
        if constexpr (TWO) ptx::tma_load_5d_cg2(dst, tm, lb, c0, c1, c2, c3, c4); else ptx::tma_load_5d(dst, tm, bar, c0, c1, c2, c3, c4);
    };
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if constexpr (TWO) ptx::umma_bf16_cg2(d, da, db, idesc, acc); else ptx::umma_bf16(d, da, db, idesc, acc);
    };
    auto commit = [&](uint64_t* bar) {
        if constexpr (TWO) ptx::umma_commit_cg2(bar, 3); else ptx::umma_commit(bar);
    };
    auto release_acc = [&](uint64_t* bar) {      // epilogue thread: this accumulator stage is drained
        if constexpr (TWO) {
            if (rank == 0) ptx::mbar_arrive(bar); else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
        } else {
            ptx::mbar_arrive(bar);
        }
    };
    const int brow_off = static_cast<int>(rank) * (BN / CG);    // this CTA's slice of the N rows of every weight block

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (warp-uniform, one lane issues)
        if (!a.static_weights) ptx::grid_dep_wait();     // the trainer's weight layouts are written earlier in the same graph
        if constexpr (KIND == K_SLAB_RES || KIND == K_PAD) {
            if (ptx::elect_one()) {
                const uint32_t lb = arm(res_bar, a.res_b_bytes);
                for (int kb = 0; kb < a.nkb; ++kb)
                    ld2(sB + kb * B_BLOCK, &tmB, res_bar, lb, kb * BLOCK_K, brow_off);
            }
            __syncwarp();
        } else if constexpr (C::DX) {
            // resident weights, re-ordered on the way in: the three dx blocks of one (dy, chunk) sit back to back, so that
            // they read as ONE K-major tile of 192 rows (row = dx * 64 + cout); global K order stays (tap, chunk)
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(res_bar, a.res_b_bytes);
                for (int kb = 0; kb < a.nkb; ++kb) {
                    const int tap = kb / chunks, chunk = kb - tap * chunks;
                    const int dyi = tap / 3, dxi = tap - dyi * 3;
                    ptx::tma_load_2d(sB + ((dyi * chunks + chunk) * 3 + dxi) * B_BLOCK, &tmB, res_bar, kb * BLOCK_K, 0);
                }
            }
            __syncwarp();
        }
        ptx::grid_dep_wait();      // weights above are constants of the plan; the activation loads below are not
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = tile0; tile < a.num_tiles; tile += tile_stride) {
            const TileCoord tc = decode_tile(a, tile);
            const int m0 = (TWO ? 2 * tc.mt + static_cast<int>(rank) : tc.mt) * BLOCK_M;   // a pair covers M tiles 2*mt, 2*mt + 1
            const int b0 = m0 / a.P;
            const int h0 = (m0 - b0 * a.P) / a.W;
            const int nrow = tc.nt * BN + tc.phase * a.N + brow_off;   // weight row of this tile (per-phase matrices stacked)
            if constexpr (KIND == K_PAD) {
                // one slab {64 ch, W + 2, rows + 2} per 64-channel chunk serves all nine taps: the box starts at column -1, so
                // the zero halo columns are part of the shared-memory row pitch and a tap is a pure row offset
                const int b = tc.mt / a.tiles_per_img;
                const int y0 = ((tc.mt - b * a.tiles_per_img) * BLOCK_M) / a.PW;
                for (int chunk = 0; chunk < chunks; ++chunk) {
                    const int cs = chunk >= src_chunks ? chunk - src_chunks : chunk;
                    const bool second = cs >= a.chunks0;
                    const CUtensorMap* tm = second ? &tmA1 : &tmA0;
                    const int c0 = (second ? cs - a.chunks0 : cs) * BLOCK_K;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], a.slab_bytes);
                        ptx::tma_load_4d(sA + stage * C::A_STAGE_BYTES, tm, &full_bar[stage], c0, -1, y0 - 1, b);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else if constexpr (KIND == K_GENERAL) {
                const int row0 = m0 / a.W;                          // merged (b, h) row for the unshuffle view
                const int dy0 = a.mode == CONV_UPSAMPLE ? (tc.phase >> 1) - 1 : -a.pad;
                const int dx0 = a.mode == CONV_UPSAMPLE ? (tc.phase & 1) - 1 : -a.pad;
                int kcol = 0;
                for (int ky = 0; ky < a.kh; ++ky) {
                    for (int kx = 0; kx < a.kw; ++kx) {
                        for (int chunk = 0; chunk < chunks; ++chunk) {
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                            if (ptx::elect_one()) {
                                const uint32_t lb = arm(&full_bar[stage], C::STAGE_BYTES);
                                const int cs = chunk >= src_chunks ? chunk - src_chunks : chunk;
                                const bool second = cs >= a.chunks0;
                                const CUtensorMap* tm = second ? &tmA1 : &tmA0;
                                const int c0 = (second ? cs - a.chunks0 : cs) * BLOCK_K;
                                uint8_t* dstA = sA + stage * C::A_STAGE_BYTES;
                                if (a.mode == CONV_UNSHUFFLE)
                                    ld5(dstA, tm, &full_bar[stage], lb, c0, kx, 0, ky, row0);
                                else
                                    ld4(dstA, tm, &full_bar[stage], lb, c0, dx0 + kx, h0 + dy0 + ky, b0);
                                ld2(sB + stage * C::B_STAGE_BYTES, &tmB, &full_bar[stage], lb, kcol, nrow);
                            }
                            __syncwarp();
                            kcol += BLOCK_K;
                            if (++stage == stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            } else if constexpr (C::DX) {
                // ONE un-shifted slab {64 ch, W, rows + 2} per chunk: the dy taps are descriptor offsets into it, the dx taps
                // are the three 64-column groups of the 192-wide accumulator (shifted by one pixel in the epilogue)
                for (int chunk = 0; chunk < chunks; ++chunk) {
                    const int cs = chunk >= src_chunks ? chunk - src_chunks : chunk;
                    const bool second = cs >= a.chunks0;
                    const CUtensorMap* tm = second ? &tmA1 : &tmA0;
                    const int c0 = (second ? cs - a.chunks0 : cs) * BLOCK_K;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], a.slab_bytes);
                        ptx::tma_load_4d(sA + stage * C::A_STAGE_BYTES, tm, &full_bar[stage], c0, 0, h0 - 1, b0);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else {
                for (int chunk = 0; chunk < chunks; ++chunk) {
                    const int cs = chunk >= src_chunks ? chunk - src_chunks : chunk;
                    const bool second = cs >= a.chunks0;
                    const CUtensorMap* tm = second ? &tmA1 : &tmA0;
                    const int c0 = (second ? cs - a.chunks0 : cs) * BLOCK_K;
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                        if (ptx::elect_one()) {
                            constexpr uint32_t wbytes = KIND == K_SLAB ? 3 * B_BLOCK : 0;
                            const uint32_t lb = arm(&full_bar[stage], a.slab_bytes + wbytes);
                            ld4(sA + stage * C::A_STAGE_BYTES, tm, &full_bar[stage], lb, c0, dxi - 1, h0 - 1, b0);
                            if constexpr (KIND == K_SLAB) {
#pragma unroll
                                for (int dyi = 0; dyi < 3; ++dyi)
                                    ld2(sB + stage * C::B_STAGE_BYTES + dyi * B_BLOCK, &tmB, &full_bar[stage], lb,
                                        ((dyi * 3 + dxi) * chunks + chunk) * BLOCK_K, nrow);
                            }
                        }
                        __syncwarp();
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform, one lane issues)
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M * CG, C::DX ? DX3_N : BN);
        const uint64_t descA0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA));
        const uint64_t descB0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB));
        if constexpr (KIND == K_SLAB_RES || KIND == K_PAD || C::DX) {
            ptx::mbar_wait(res_bar, 0);
            ptx::tc_fence_after();
        }
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int tile = tile0; tile < a.num_tiles; tile += tile_stride, ++iter) {
            const int as = iter % ACC;
            const uint32_t aphase = (iter / ACC) & 1u;
            ptx::mbar_wait(&tempty_bar[as], aphase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * C::ACC_STRIDE;
            if constexpr (C::DX) {
                const uint32_t dy_step = a.slab_dy_bytes >> 4;
                for (int chunk = 0; chunk < chunks; ++chunk) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t da = descA0 + static_cast<uint64_t>((stage * C::A_STAGE_BYTES) >> 4);
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            const uint64_t db = descB0 + static_cast<uint64_t>((((dyi * chunks + chunk) * 3) * B_BLOCK) >> 4);
                            const uint64_t dad = da + static_cast<uint64_t>(dyi * dy_step);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                ptx::umma_bf16(tmem_d, dad + 2u * k, db + 2u * k, idesc, (chunk | dyi | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                        if (chunk == chunks - 1) ptx::umma_commit(&tfull_bar[as]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else if constexpr (KIND == K_PAD) {
                const TileCoord tc = decode_tile(a, tile);
                const int p0 = (tc.mt % a.tiles_per_img) * BLOCK_M;
                const int o0 = p0 - (p0 / a.PW) * a.PW;             // padded column of the tile's first position
                const uint32_t sA_u32 = ptx::smem_u32(sA);
                for (int chunk = 0; chunk < chunks; ++chunk) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            // A row of GEMM row i is slab pixel (1 + dy) * PW + o0 + dx + i: the start is NOT a multiple of the
                            // 8-row swizzle atom.  Measured on B200: the tensor core derives the 128B-swizzle phase from the
                            // absolute shared-memory address bits (like the TMA unit that wrote the slab), so a plain
                            // 128-byte-granular start address is correct and the descriptor's base-offset field must stay 0
                            // (setting it to (addr >> 7) & 7 double-applies the phase: tests/test_ops_gpu.py::test_conv_gemm).
                            const int row = (tap / 3) * a.PW + o0 + (tap % 3) - 1;
                            const uint32_t addr = sA_u32 + stage * C::A_STAGE_BYTES + row * 128;
                            const uint64_t da = ptx::make_kmajor_sw128_desc(addr);
                            const uint64_t db = descB0 + static_cast<uint64_t>(((tap * chunks + chunk) * B_BLOCK) >> 4);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                ptx::umma_bf16(tmem_d, da + 2u * k, db + 2u * k, idesc, (chunk | tap | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                        if (chunk == chunks - 1) ptx::umma_commit(&tfull_bar[as]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else if constexpr (KIND == K_GENERAL) {
                for (int kb = 0; kb < a.nkb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t da = descA0 + static_cast<uint64_t>((stage * C::A_STAGE_BYTES) >> 4);
                        const uint64_t db = descB0 + static_cast<uint64_t>((stage * C::B_STAGE_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k)   // +32 bytes along K inside the swizzle span: +2 in the >>4 field
                            mma(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        commit(&empty_bar[stage]);
                        if (kb == a.nkb - 1) commit(&tfull_bar[as]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else {
                const int nst = chunks * 3;
                const uint32_t dy_step = a.slab_dy_bytes >> 4;
                int chunk = 0, dxi = 0;
                for (int st = 0; st < nst; ++st) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t da = descA0 + static_cast<uint64_t>((stage * C::A_STAGE_BYTES) >> 4);
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            uint64_t db;
                            if constexpr (KIND == K_SLAB)
                                db = descB0 + static_cast<uint64_t>((stage * C::B_STAGE_BYTES + dyi * B_BLOCK) >> 4);
                            else
                                db = descB0 + static_cast<uint64_t>((((dyi * 3 + dxi) * chunks + chunk) * B_BLOCK) >> 4);
                            const uint64_t dad = da + static_cast<uint64_t>(dyi * dy_step);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                mma(tmem_d, dad + 2u * k, db + 2u * k, idesc, (st | dyi | k) != 0 ? 1u : 0u);
                        }
                        commit(&empty_bar[stage]);
                        if (st == nst - 1) commit(&tfull_bar[as]);
                    }
                    __syncwarp();
                    if (++dxi == 3) { dxi = 0; ++chunk; }
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 2) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        // Warp (q, hc): TMEM lane quarter q = warp % 4 (32 accumulator rows), column half hc of every 64-column chunk.
        // Two warps per scheduler keep the long dependent chains of the epilogue overlapped (with 4 warps ncu showed the
        // kernel epilogue-bound: one warp per scheduler issues ~1 instruction per 4-5 cycles).  bf16 output goes through a
        // per-warp, 64B-swizzled staging buffer and leaves as TMA tensor stores of 32 rows x 32 columns (rows past M
        // are clipped by the TMA unit); the fp32 head path (N padded to 16, one valid column) stores directly.
        ptx::grid_dep_wait();       // residual reads and every store must follow the predecessor grid (the arena is reused)
        const int q = warp & 3;
        const int hc = ((warp - 2) >> 2) & 1;
        const int r = q * 32 + lane;            // accumulator row == pixel within the tile
        const ConvEpilogue& e = a.epi;
        uint8_t* my_stage = sEpi + (hc * 4 + q) * (C::EPI_BUFS * EPI_BUF_BYTES);
        const int swz = (lane >> 1) & 3;        // 64B swizzle: 16-byte chunk j of row `lane` lives at chunk j ^ swz
        int buf = 0;

        // pack 32 fp32 -> bf16, stage, TMA-store as the 32 x 32 box at (column ncol, row block of this warp)
        auto stage_and_store = [&](const float (&f)[32], int ncol, const TileCoord& tc, const CUtensorMap* tmOut = nullptr) {
            if (tmOut == nullptr) tmOut = &tmD;
            if (lane == 0) {
                if constexpr (C::EPI_BUFS == 2) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>();
            }
            __syncwarp();
            uint8_t* stage_row = my_stage + buf * EPI_BUF_BYTES + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = ptx::pack_bf16x2(f[8 * j], f[8 * j + 1]);
                o.y = ptx::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                o.z = ptx::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                o.w = ptx::pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                *reinterpret_cast<uint4*>(stage_row + ((j ^ swz) << 4)) = o;
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                const int mrow = tc.mt * BLOCK_M + q * 32;
                if (a.mode == CONV_UPSAMPLE) {
                    // low-res pixel block -> output phase (pa, pb) through the [N, 2, W/2, 2, B*H/2] view
                    const int rowl = mrow / a.Wl_box;              // merged (b, low-res row)
                    const int wl = a.rows_box > 1 ? 0 : mrow - rowl * a.Wl_box;
                    ptx::tma_store_5d(tmOut, my_stage + buf * EPI_BUF_BYTES, ncol, tc.phase & 1, wl, tc.phase >> 1, rowl);
                } else {
                    ptx::tma_store_2d(tmOut, my_stage + buf * EPI_BUF_BYTES, ncol, mrow);
                }
                ptx::bulk_commit();
            }
            if constexpr (C::EPI_BUFS == 2) buf ^= 1;
        };
        // (sum, M2) of the four 8-channel pieces of this warp's 32 columns over its 32 rows -> gn_part
        auto write_partials = [&](const float (&f)[32], int ncol, int mt) {
            float sq[8];
#pragma unroll
            for (int p8 = 0; p8 < 4; ++p8) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s1 += f[8 * p8 + j];
                    s2 = fmaf(f[8 * p8 + j], f[8 * p8 + j], s2);
                }
                sq[2 * p8] = s1;
                sq[2 * p8 + 1] = s2;
            }
            const float mine = transpose_reduce8(sq, lane);                   // lane L: value (L >> 2)
            const float other = __shfl_down_sync(0xffffffffu, mine, 4);
            if ((lane & 7) == 0) {                                            // lane 8p: sum and sum of squares of piece p
                const float m2 = fmaxf(other - mine * mine * (1.0f / 256.0f), 0.f);
                e.gn_part[(static_cast<size_t>(mt) * 4 + q) * (a.N >> 3) + (ncol >> 3) + (lane >> 3)] = make_float2(mine, m2);
            }
        };

        if constexpr (GNF) {
            // ---- GroupNorm-fused epilogue (slab kinds, P >= 128, N <= 128): the conv's output leaves the SM already
            // normalised.  Phase A of tile i (as soon as its accumulator is complete): per-piece (sum, M2) partials from
            // the fp32 accumulators + bias -> global, then one release-increment of the image's arrival counter per warp.
            // Phase B trails by one tile: wait until every warp block of the image has arrived (all CTAs of this persistent
            // grid are co-resident and run phase A before any wait on it, so the wait cannot deadlock), merge the image's
            // partials in a fixed order (Chan), fold bias / gamma / beta / FiLM into one multiply-add per column, re-read
            // the accumulator from TMEM, SiLU, (+ SR3 post-add) (+ residual), TMA store.
            const int te = ((warp - 2) << 5) + lane;        // 0..255
            const int cpg = a.N >> 3;                       // channels per group
            const int ppg = a.N >> 6;                       // 8-channel pieces per group
            const int ppr = a.N >> 3;                       // pieces per partial row
            const int nwb = a.P >> 5;                       // warp blocks per image
            const int target = nwb * a.num_n_tiles * 2;     // two half-width warps per (warp block, N tile)
            const float inv_total = 1.0f / (static_cast<float>(a.P) * static_cast<float>(cpg));
            int prev_tile = -1, prev_iter = 0;
            for (int iter = 0;; ++iter) {
                const int tile = tile0 + iter * tile_stride;
                const bool have = tile < a.num_tiles;
                if (have) {
                    // ------------------------------------------------ phase A: statistics of tile `iter`
                    const TileCoord tc = decode_tile(a, tile);
                    const int n0 = tc.nt * BN;
                    const int as = iter % ACC;
                    ptx::mbar_wait(&tfull_bar[as], (iter / ACC) & 1u);
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + hc * 32;
#pragma unroll 1
                    for (int c = 0; c < BN; c += 64) {
                        uint32_t v[32];
                        ptx::tmem_ld32(taddr + c, v);
                        ptx::tmem_ld_wait();
                        const int ncol = n0 + c + hc * 32;
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + ncol + j));
                            f[j] = __uint_as_float(v[j]) + bb.x; f[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
                            f[j + 2] = __uint_as_float(v[j + 2]) + bb.z; f[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
                        }
                        write_partials(f, ncol, tc.mt);
                    }
                    __syncwarp();
                    if (lane == 0) red_release_add(e.gn_counter + (tc.mt * BLOCK_M + q * 32) / a.P, 1);
                }
                if (prev_tile >= 0) {
                    // ------------------------------------------------ phase B: normalise + store tile `prev_iter`
                    const TileCoord tc = decode_tile(a, prev_tile);
                    const int mt = tc.mt;
                    const int n0 = tc.nt * BN;
                    const int m = mt * BLOCK_M + r;
                    const int b = (mt * BLOCK_M) / a.P;                 // P >= 128: the whole tile lies in one image
                    const int as = prev_iter % ACC;
                    const int gb = prev_iter & 1;
                    // column constants do not depend on the statistics: issue their loads first so the latency hides
                    // behind the arrival wait below
                    float cgam = 0.f, cbet = 0.f, cbias = 0.f, csc = 1.0f, csh = 0.f, cpost = 0.f;
                    if (te < BN) {
                        const int c = n0 + te;
                        cgam = __ldg(e.gn_gamma + c);
                        cbet = __ldg(e.gn_beta + c);
                        cbias = __ldg(e.bias + c);
                        if (e.film != nullptr) {
                            const int row = e.film_row[b * e.film_row_stride];
                            const float* frow = e.film + static_cast<size_t>(row) * e.film_ld + e.film_off;
                            if (e.film_has_scale) {
                                csc = __ldg(frow + c) + 1.0f;
                                csh = __ldg(frow + a.N + c);
                            } else {
                                cpost = __ldg(frow + c);                       // SR3: additive noise embedding after the activation
                            }
                        }
                    }
                    if (lane == 0) {
                        const int* cnt = e.gn_counter + b;
                        unsigned long long t0 = 0;
                        unsigned spins = 0;
                        while (ld_relaxed_gpu(cnt) < target) {
                            if ((++spins & 0xffu) == 0) {
                                unsigned long long t1;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                                if (t0 == 0) t0 = t1;
                                if (t1 - t0 > 2000000000ull) __trap();   // a protocol bug must surface as an error, not a hang
                            }
                        }
                    }
                    __syncwarp();
                    // epilogue warp w merges group w of image b: all its partials in ONE batch of loads (<= 8 per lane),
                    // then fixed-order shuffle reductions -> bit-identical statistics in every CTA
                    {
                        constexpr int MAXE = 8;                              // nent <= 256 (conv_gemm_can_fuse_gn)
                        const int g = warp - 2;
                        const int nent = nwb * ppg;
                        const float2* pp = e.gn_part + (static_cast<size_t>(b) * nwb) * ppr + g * ppg;
                        float2 ent[MAXE];
#pragma unroll
                        for (int i = 0; i < MAXE; ++i) {
                            const int idx = lane + 32 * i;
                            const int blk = idx / ppg;
                            ent[i] = idx < nent ? ld_cg_f2(pp + static_cast<size_t>(blk) * ppr + (idx - blk * ppg)) : make_float2(0.f, 0.f);
                        }
                        float s1 = 0.f;
#pragma unroll
                        for (int i = 0; i < MAXE; ++i) s1 += ent[i].x;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                        const float mean = s1 * inv_total;
                        float m2 = 0.f;
#pragma unroll
                        for (int i = 0; i < MAXE; ++i) {
                            if (lane + 32 * i < nent) {
                                const float dm = ent[i].x * (1.0f / 256.0f) - mean;
                                m2 += ent[i].y + 256.0f * dm * dm;
                            }
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
                        if (lane == 0) {
                            s_gn[gb * 16 + g] = mean;
                            s_gn[gb * 16 + 8 + g] = rsqrtf(m2 * inv_total + e.gn_eps);
                        }
                    }
                    named_bar_sync(2, 256);
                    if (te < BN) {
                        const int g = (n0 + te) / cpg;
                        float mul = cgam * s_gn[gb * 16 + 8 + g];
                        float add = cbet - s_gn[gb * 16 + g] * mul;
                        add = fmaf(cbias, mul, add);                         // (acc + bias) * mul + add
                        if (e.film != nullptr && e.film_has_scale) {
                            mul *= csc;
                            add = fmaf(add, csc, csh);
                        }
                        s_ma[gb * 3 * BN + te] = mul;
                        s_ma[gb * 3 * BN + BN + te] = add;
                        s_ma[gb * 3 * BN + 2 * BN + te] = cpost;
                    }
                    named_bar_sync(2, 256);
                    const float* smul = s_ma + gb * 3 * BN + hc * 32;
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + hc * 32;
#pragma unroll 1
                    for (int c = 0; c < BN; c += 64) {
                        uint32_t v[32];
                        ptx::tmem_ld32(taddr + c, v);
                        ptx::tmem_ld_wait();
                        if (c + 64 == BN) {          // accumulator fully drained: hand the TMEM stage back to the MMA warp
                            ptx::tc_fence_before();
                            ptx::mbar_arrive(&tempty_bar[as]);
                        }
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 m0 = *reinterpret_cast<const float4*>(smul + c + j);
                            const float4 a0 = *reinterpret_cast<const float4*>(smul + BN + c + j);
                            const float4 p0 = *reinterpret_cast<const float4*>(smul + 2 * BN + c + j);
                            f[j] = silu_tanh(fmaf(__uint_as_float(v[j]), m0.x, a0.x)) + p0.x;
                            f[j + 1] = silu_tanh(fmaf(__uint_as_float(v[j + 1]), m0.y, a0.y)) + p0.y;
                            f[j + 2] = silu_tanh(fmaf(__uint_as_float(v[j + 2]), m0.z, a0.z)) + p0.z;
                            f[j + 3] = silu_tanh(fmaf(__uint_as_float(v[j + 3]), m0.w, a0.w)) + p0.w;
                        }
                        const int ncol = n0 + c + hc * 32;
                        if (e.res != nullptr) {
                            const uint4* rp = reinterpret_cast<const uint4*>(e.res + static_cast<size_t>(m) * e.ldr + ncol);
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 ra = __ldg(rp + j / 8);
                                float2 t;
                                t = ptx::unpack_bf16x2(ra.x); f[j] += t.x; f[j + 1] += t.y;
                                t = ptx::unpack_bf16x2(ra.y); f[j + 2] += t.x; f[j + 3] += t.y;
                                t = ptx::unpack_bf16x2(ra.z); f[j + 4] += t.x; f[j + 5] += t.y;
                                t = ptx::unpack_bf16x2(ra.w); f[j + 6] += t.x; f[j + 7] += t.y;
                            }
                        }
                        stage_and_store(f, ncol, tc);
                    }
                }
                if (!have) break;
                prev_tile = tile;
                prev_iter = iter;
            }
        } else if constexpr (KIND == K_DX3G) {
            // ---- dx-stacked conv, TWO epilogue groups of 8 warps on alternating tiles (group g owns accumulator stage g).
            // Measured on the one-group form (profiles/r02_notes.md 10): at K = 576 the MMAs of a tile take ~2050 clk but one
            // warp's epilogue chain ~3600 clk (659 instructions issued at 1 per ~5 clk: in-order dependent chains, two warps per
            // scheduler) -- the kernel was epilogue-latency bound.  Two tiles in flight hide that chain.  To fit 576 threads
            // (<= 112 registers) a warp walks its 32 columns in two halves of 16 that fill one 2 KiB staging box, stored by TMA
            // once per tile (16-byte stores straight from registers were measured 1.7x SLOWER: 32 lines per instruction); the
            // shifted sums are FMAs against 0 / 1 lane masks, the boundary rows between lane quarters use one barrier per half.
            const int grp = (warp - 2) >> 3;
            const int xw = r & (a.W - 1);
            const bool cross = a.W > 32;                                  // warp-uniform: quarter boundaries inside an image row
            const float m_l = (xw != 0 && lane != 0) ? 1.f : 0.f;         // take D0 of the previous lane
            const float m_r = (xw != a.W - 1 && lane != 31) ? 1.f : 0.f;  // take D2 of the next lane
            const float m_e = (cross && ((lane == 0 && xw != 0) || (lane == 31 && xw != a.W - 1))) ? 1.f : 0.f;   // take the exchanged row
            const int pair = hc * 2 + (q >> 1);
            float* xs0 = s_xchg + grp * 256 + pair * 32;                  // [2 halves][4 pairs][D0 row: 16 | D2 row: 16] per group
            const int bar_id = 3 + grp * 4 + pair;
            uint8_t* my_box = sEpi + (warp - 2) * 2048;                   // 32 rows x 32 columns bf16, 64-byte rows, 64B swizzle (as tmD expects)
            for (int iter = grp;; iter += 2) {
                const int tile = tile0 + iter * tile_stride;
                if (tile >= a.num_tiles) break;
                const TileCoord tc = decode_tile(a, tile);
                const int mt = tc.mt;
                const int m = mt * BLOCK_M + r;
                ptx::mbar_wait(&tfull_bar[grp], (iter >> 1) & 1u);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * C::ACC_STRIDE + hc * 32;
                float sq[8];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int ncol = hc * 32 + half * 16;                 // BN == N == 64: one N tile
                    uint32_t v0[16], v1[16], v2[16];
                    ptx::tmem_ld16(taddr + half * 16, v0);
                    ptx::tmem_ld16(taddr + 64 + half * 16, v1);
                    ptx::tmem_ld16(taddr + 128 + half * 16, v2);
                    ptx::tmem_ld_wait();
                    if (half == 1) {                                      // accumulator drained: the MMA warp may start tile iter + 2
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&tempty_bar[grp]);
                    }
                    float* xs = xs0 + half * 128;
                    if (cross) {
                        if ((q & 1) == 0 && lane == 31) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(xs + j) = make_float4(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1]),
                                                                                   __uint_as_float(v0[j + 2]), __uint_as_float(v0[j + 3]));
                        }
                        if ((q & 1) == 1 && lane == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(xs + 16 + j) = make_float4(__uint_as_float(v2[j]), __uint_as_float(v2[j + 1]),
                                                                                        __uint_as_float(v2[j + 2]), __uint_as_float(v2[j + 3]));
                        }
                        named_bar_sync(bar_id, 64);
                        // (the buffer of this half is rewritten one tile later, after both warps passed the other half's barrier)
                    }
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bb = e.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(e.bias + ncol + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        f[j] = __uint_as_float(v1[j]) + bb.x; f[j + 1] = __uint_as_float(v1[j + 1]) + bb.y;
                        f[j + 2] = __uint_as_float(v1[j + 2]) + bb.z; f[j + 3] = __uint_as_float(v1[j + 3]) + bb.w;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);
                        f[j] = fmaf(up, m_l, f[j]);
                        f[j] = fmaf(dn, m_r, f[j]);
                    }
                    if (cross) {
                        const float* ep = xs + ((q & 1) ? 0 : 16);       // upper warp: the lower one's D0 row, and vice versa
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 ev = *reinterpret_cast<const float4*>(ep + j);
                            f[j] = fmaf(ev.x, m_e, f[j]); f[j + 1] = fmaf(ev.y, m_e, f[j + 1]);
                            f[j + 2] = fmaf(ev.z, m_e, f[j + 2]); f[j + 3] = fmaf(ev.w, m_e, f[j + 3]);
                        }
                    }
                    if (e.out_scale != 1.0f) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] *= e.out_scale;
                    }
                    if (e.res != nullptr) {
                        const uint4* rp = reinterpret_cast<const uint4*>(e.res + static_cast<size_t>(m) * e.ldr + ncol);
#pragma unroll
                        for (int j = 0; j < 16; j += 8) {
                            const uint4 rr = __ldg(rp + j / 8);
                            float2 t;
                            t = ptx::unpack_bf16x2(rr.x); f[j] += t.x; f[j + 1] += t.y;
                            t = ptx::unpack_bf16x2(rr.y); f[j + 2] += t.x; f[j + 3] += t.y;
                            t = ptx::unpack_bf16x2(rr.z); f[j + 4] += t.x; f[j + 5] += t.y;
                            t = ptx::unpack_bf16x2(rr.w); f[j + 6] += t.x; f[j + 7] += t.y;
                        }
                    }
                    // this half's two 8-channel pieces: (sum, sum of squares) per lane, reduced over the warp after both halves
#pragma unroll
                    for (int p8 = 0; p8 < 2; ++p8) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            s1 += f[8 * p8 + j];
                            s2 = fmaf(f[8 * p8 + j], f[8 * p8 + j], s2);
                        }
                        sq[half * 4 + 2 * p8] = s1;
                        sq[half * 4 + 2 * p8 + 1] = s2;
                    }
                    // both halves go into ONE 32 x 32 box (this half = 16-byte chunks 2 * half, 2 * half + 1 of the 64-byte row) and
                    // leave with one fence + one TMA store per tile; the box was handed to the TMA unit a whole tile pair ago
                    if (half == 0) {
                        if (lane == 0) ptx::bulk_wait_read<0>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 o;
                        o.x = ptx::pack_bf16x2(f[8 * j], f[8 * j + 1]);
                        o.y = ptx::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                        o.z = ptx::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                        o.w = ptx::pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                        *reinterpret_cast<uint4*>(my_box + lane * 64 + (((2 * half + j) ^ swz) << 4)) = o;
                    }
                    if (half == 1) {
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&tmD, my_box, hc * 32, mt * BLOCK_M + q * 32);
                            ptx::bulk_commit();
                        }
                    }
                }
                if (e.gn_part != nullptr) {
                    const float mine = transpose_reduce8(sq, lane);                   // lane L: value (L >> 2)
                    const float other = __shfl_down_sync(0xffffffffu, mine, 4);
                    if ((lane & 7) == 0) {                                            // lane 8p: sum and sum of squares of piece p
                        const float m2 = fmaxf(other - mine * mine * (1.0f / 256.0f), 0.f);
                        e.gn_part[(static_cast<size_t>(mt) * 4 + q) * (a.N >> 3) + (hc * 4) + (lane >> 3)] = make_float2(mine, m2);
                    }
                }
            }
        } else if constexpr (KIND == K_PAD) {
            // ---- padded-slab epilogue: GEMM row i of tile mt is PADDED position p0 + i of image b (row pitch W + 2); the
            // positions that fall on a halo column (or past the image) carry garbage and are masked out of the statistics
            // and clipped from the output by the TMA unit (negative / >= W column coordinates are simply not written).
            int iter = 0;
            for (int tile = tile0; tile < a.num_tiles; tile += tile_stride, ++iter) {
                const TileCoord tc = decode_tile(a, tile);
                const int mt = tc.mt;
                const int b = mt / a.tiles_per_img;
                const int p0 = (mt - b * a.tiles_per_img) * BLOCK_M;
                const int y0 = p0 / a.PW;
                const int o0 = p0 - y0 * a.PW;
                const int pos = o0 + r;
                const int yy = pos / a.PW;
                const int xp = pos - yy * a.PW;
                const bool valid = xp >= 1 && xp <= a.W && (y0 + yy) < a.H;
                const int as = iter % ACC;
                ptx::mbar_wait(&tfull_bar[as], (iter / ACC) & 1u);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + hc * 32;
                const int ncol = hc * 32;                       // BN == N == 64: one N tile
                uint32_t v[32];
                ptx::tmem_ld32(taddr, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[as]);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + ncol + j));
                    f[j] = __uint_as_float(v[j]) + bb.x; f[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
                    f[j + 2] = __uint_as_float(v[j + 2]) + bb.z; f[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
                }
                if (e.gn_part != nullptr) {
                    // (sum, M2) of each 8-channel piece over the VALID rows of this warp block; the consumer recomputes the
                    // valid-row count from the geometry (groupnorm_apply_kernel, padded partial layout)
                    const int nv = __popc(__ballot_sync(0xffffffffu, valid));
                    float sq[8];
#pragma unroll
                    for (int p8 = 0; p8 < 4; ++p8) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float x = valid ? f[8 * p8 + j] : 0.f;
                            s1 += x;
                            s2 = fmaf(x, x, s2);
                        }
                        sq[2 * p8] = s1;
                        sq[2 * p8 + 1] = s2;
                    }
                    const float mine = transpose_reduce8(sq, lane);
                    const float other = __shfl_down_sync(0xffffffffu, mine, 4);
                    if ((lane & 7) == 0) {
                        const float cnt = 8.0f * static_cast<float>(nv);
                        const float m2 = nv > 0 ? fmaxf(other - mine * mine / cnt, 0.f) : 0.f;
                        e.gn_part[(static_cast<size_t>(mt) * 4 + q) * (a.N >> 3) + (ncol >> 3) + (lane >> 3)] = make_float2(mine, m2);
                    }
                }
                // stage (64B swizzle) and store.  TMA stores reject negative coordinates, so the valid rows are COMPACTED in the
                // staging buffer: they are consecutive pixels of the dense [B*H*W, N] output (a halo pair sits exactly where
                // one image row ends and the next begins), and the 32 - nv halo rows are dropped by using the tensor map whose
                // box has nv rows (nv is 32, 31 or 30; 0 for a block past the end of the image).
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                const int nvr = __popc(vmask);
                const int srow = __popc(vmask & ((1u << lane) - 1u));           // compacted staging row of this lane
                if (lane == 0) ptx::bulk_wait_read<0>();
                __syncwarp();
                if (valid) {
                    uint8_t* stage_row = my_stage + srow * 64;
                    const int sw2 = (srow >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = ptx::pack_bf16x2(f[8 * j], f[8 * j + 1]);
                        o.y = ptx::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                        o.z = ptx::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                        o.w = ptx::pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                        *reinterpret_cast<uint4*>(stage_row + ((j ^ sw2) << 4)) = o;
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                // dense pixel index of the first valid row (broadcast from the first valid lane)
                const int mpix = (b * a.H + y0 + yy) * a.W + (xp - 1);
                const int first = nvr > 0 ? __ffs(vmask) - 1 : 0;
                const int m_first = __shfl_sync(0xffffffffu, mpix, first);
                if (lane == 0 && nvr > 0) {
                    const CUtensorMap* tm = nvr == 32 ? &tmD : (nvr == 31 ? &tmD31 : &tmD30);
                    ptx::tma_store_2d(tm, my_stage, ncol, m_first);
                    ptx::bulk_commit();
                }
            }
        } else {
            int iter = 0;
            for (int tile = tile0; tile < a.num_tiles; tile += tile_stride, ++iter) {
                TileCoord tc = decode_tile(a, tile);
                if constexpr (TWO) tc.mt = 2 * tc.mt + static_cast<int>(rank);      // this CTA's M tile of the pair
                const int mt = tc.mt;
                const int m = mt * BLOCK_M + r;
                const int n0 = tc.nt * BN;
                const bool valid = m < a.M;
                const int as = iter % ACC;
                const uint32_t aphase = (iter / ACC) & 1u;

                const float* frow = nullptr;
                if (e.film != nullptr) {
                    const int b = valid ? m / a.P : 0;
                    const int row = e.film_row[b * e.film_row_stride];
                    frow = e.film + static_cast<size_t>(row) * e.film_ld + e.film_off;
                }
                const int shift_off = e.film_has_scale ? a.N : 0;

                ptx::mbar_wait(&tfull_bar[as], aphase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * C::ACC_STRIDE;

                if constexpr (BN < 64) {
                    // fp32 head (tail conv of HiCEDRN): 16 accumulator columns, n_valid of them real; column half 0 does the work
                    uint32_t v[16];
                    if (hc == 0) {
                        ptx::tmem_ld16(taddr, v);
                        ptx::tmem_ld_wait();
                    }
                    ptx::tc_fence_before();
                    release_acc(&tempty_bar[as]);
                    if (hc == 0 && valid && e.out_f32 != nullptr) {
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
                        const int ncol = n0 + c + hc * 32;
                        float f[32];
                        if constexpr (KIND == K_DX3) {
                            // accumulator column groups: [0, 64) = dx -1, [64, 128) = dx 0, [128, 192) = dx +1, all computed on the
                            // UN-shifted pixel rows: out[r] = D0[r - 1] + D1[r] + D2[r + 1], the neighbours dropped where the tap
                            // falls on the conv's zero padding (x == 0 / x == W - 1).  A tile is whole image rows, so r - 1 / r + 1
                            // never leave it; across the 32-lane TMEM quarters the rows travel through shared memory (only at
                            // W = 64: with W <= 32 every quarter boundary is a row boundary).
                            uint32_t v0[32], v1[32], v2[32];
                            ptx::tmem_ld32(taddr + hc * 32, v0);
                            ptx::tmem_ld32(taddr + 64 + hc * 32, v1);
                            ptx::tmem_ld32(taddr + 128 + hc * 32, v2);
                            ptx::tmem_ld_wait();
                            ptx::tc_fence_before();
                            ptx::mbar_arrive(&tempty_bar[as]);
                            const int xw = r & (a.W - 1);
                            const bool cross = a.W > 32;                      // warp-uniform
                            // 0 / 1 lane masks: the shifted sums are FMAs (a select per term made the one-group epilogue 659
                            // instructions per tile, profiles/r02_notes.md 10); lanes 0 / 31 take the exchanged row instead
                            const float m_l = (xw != 0 && lane != 0) ? 1.f : 0.f;
                            const float m_r = (xw != a.W - 1 && lane != 31) ? 1.f : 0.f;
                            const float m_e = (cross && ((lane == 0 && xw != 0) || (lane == 31 && xw != a.W - 1))) ? 1.f : 0.f;
                            float4 edge[8];
                            if (cross) {
                                const int pair = hc * 2 + (q >> 1);
                                float* xs = s_xchg + pair * 64;            // [0, 32): D0 of the lower warp's last row; [32, 64): D2 of the upper warp's first row
                                if ((q & 1) == 0 && lane == 31) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 4)
                                        *reinterpret_cast<float4*>(xs + j) = make_float4(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1]),
                                                                                           __uint_as_float(v0[j + 2]), __uint_as_float(v0[j + 3]));
                                }
                                if ((q & 1) == 1 && lane == 0) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 4)
                                        *reinterpret_cast<float4*>(xs + 32 + j) = make_float4(__uint_as_float(v2[j]), __uint_as_float(v2[j + 1]),
                                                                                                __uint_as_float(v2[j + 2]), __uint_as_float(v2[j + 3]));
                                }
                                named_bar_sync(3 + pair, 64);
                                const float* ep = xs + ((q & 1) ? 0 : 32);    // upper warp reads the lower one's D0, and vice versa
#pragma unroll
                                for (int j = 0; j < 8; ++j) edge[j] = *reinterpret_cast<const float4*>(ep + 4 * j);
                                named_bar_sync(3 + pair, 64);                 // the rows are consumed: the next tile may overwrite them
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);
                                const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);
                                f[j] = fmaf(dn, m_r, fmaf(up, m_l, __uint_as_float(v1[j])));
                            }
                            if (cross) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = fmaf(reinterpret_cast<const float*>(edge)[j], m_e, f[j]);
                            }
                        } else {
                            uint32_t v[32];
                            ptx::tmem_ld32(taddr + c + hc * 32, v);
                            ptx::tmem_ld_wait();
                            if (c + 64 == BN) {          // accumulator fully drained: hand the TMEM stage back to the MMA warp
                                ptx::tc_fence_before();
                                release_acc(&tempty_bar[as]);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        }
                        // bias / FiLM / SiLU / scale / residual on this warp's 32 columns
                        if (e.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + ncol + j));
                                f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
                            }
                        }
                        if (frow != nullptr) {
                            if (e.film_has_scale) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 sc = __ldg(reinterpret_cast<const float4*>(frow + ncol + j));
                                    f[j] *= (sc.x + 1.0f); f[j + 1] *= (sc.y + 1.0f);
                                    f[j + 2] *= (sc.z + 1.0f); f[j + 3] *= (sc.w + 1.0f);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 sh = __ldg(reinterpret_cast<const float4*>(frow + shift_off + ncol + j));
                                f[j] += sh.x; f[j + 1] += sh.y; f[j + 2] += sh.z; f[j + 3] += sh.w;
                            }
                        }
                        if (e.silu == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
                        } else if (e.silu == 2) {          // exact form (ex2 + rcp): "bf16w2" precision
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
                        }
                        if (e.out_scale != 1.0f) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] *= e.out_scale;
                        }
                        if (e.res != nullptr && valid && e.gnres != nullptr) {
                            // ResnetBlock tail: res is block2's RAW conv output; add SiLU(GroupNorm(res)) with the folded affine
                            const uint4* rp = reinterpret_cast<const uint4*>(e.res + static_cast<size_t>(m) * e.ldr + ncol);
                            const float4* mp = reinterpret_cast<const float4*>(e.gnres + static_cast<size_t>(m / a.P) * a.N + ncol);
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 rr = __ldg(rp + j / 8);
                                const float4 q0 = __ldg(mp + j / 2), q1 = __ldg(mp + j / 2 + 1), q2 = __ldg(mp + j / 2 + 2), q3 = __ldg(mp + j / 2 + 3);
                                float2 t;
                                t = ptx::unpack_bf16x2(rr.x); f[j] += silu_f(fmaf(t.x, q0.x, q0.y)); f[j + 1] += silu_f(fmaf(t.y, q0.z, q0.w));
                                t = ptx::unpack_bf16x2(rr.y); f[j + 2] += silu_f(fmaf(t.x, q1.x, q1.y)); f[j + 3] += silu_f(fmaf(t.y, q1.z, q1.w));
                                t = ptx::unpack_bf16x2(rr.z); f[j + 4] += silu_f(fmaf(t.x, q2.x, q2.y)); f[j + 5] += silu_f(fmaf(t.y, q2.z, q2.w));
                                t = ptx::unpack_bf16x2(rr.w); f[j + 6] += silu_f(fmaf(t.x, q3.x, q3.y)); f[j + 7] += silu_f(fmaf(t.y, q3.z, q3.w));
                            }
                        } else if (e.res != nullptr && valid) {
                            const uint4* rp = reinterpret_cast<const uint4*>(e.res + static_cast<size_t>(m) * e.ldr + ncol);
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
                        if (e.gn_part != nullptr && mt < a.m_tiles_real) write_partials(f, ncol, mt);   // rows are all valid: M % 32 == 0
                        stage_and_store(f, ncol, tc);
                        if (e.out_lo != nullptr) {       // low half through the second output map (tmD31 doubles as it)
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] -= __bfloat162float(__float2bfloat16(f[j]));
                            stage_and_store(f, ncol, tc, &tmD31);
                        }
                    }
                }
            }
        }   // !GNF
        if (lane == 0) ptx::bulk_wait_all();    // all tensor stores of this warp have landed before the CTA retires
        __syncwarp();
    }

    ptx::tc_fence_before();
    if constexpr (TWO) ptx::cluster_sync_all(); else __syncthreads();   // the peer's smem / TMEM stay alive until the leader's MMAs retired
    if (warp == 0) {
        if constexpr (TWO) ptx::tmem_dealloc_cg2(tmem_base, C::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
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

}  // namespace

int encode_tmap_bf16(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, char* err, int errlen, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available");
        return 1;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(err, errlen, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu/%llu/%llu)", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull);
        return 1;
    }
    return 0;
}

namespace {

inline int encode_map(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, char* err, int errlen, int swizzle_bytes = 128) {
    return encode_tmap_bf16(tm, ptr, rank, dims, strides_bytes, box, err, errlen, swizzle_bytes);
}

// Hs x Ws: spatial size of the SOURCE tensor the taps walk over (== output size except for CONV_UPSAMPLE: low-res).
int encode_activation_map(CUtensorMap* tm, const ConvSrc& s, const ConvGemmDesc& d, int Hs, int Ws, int slab_rows,
                          char* err, int errlen, int box_w = 0) {
    const cuuint64_t C = s.C;
    if (d.mode != CONV_UNSHUFFLE) {
        const int P = Hs * Ws;
        int rows = P >= BLOCK_M ? BLOCK_M / Ws : Hs;
        const int imgs = P >= BLOCK_M ? 1 : BLOCK_M / P;
        if (slab_rows > 0) rows = slab_rows;
        cuuint64_t dims[4] = {C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)d.B};
        cuuint64_t str[3] = {C * 2, C * 2 * Ws, C * 2 * Ws * Hs};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)(box_w > 0 ? box_w : Ws), (cuuint32_t)rows, (cuuint32_t)imgs};
        return encode_map(tm, s.ptr, 4, dims, str, box, err, errlen);
    }
    // input is [B, 2H, 2W, C]; view as [C, p2, W, p1, B*H]
    cuuint64_t dims[5] = {C, 2, (cuuint64_t)d.W, 2, (cuuint64_t)d.B * d.H};
    cuuint64_t str[4] = {C * 2, C * 2 * 2, C * 2 * 2 * d.W, C * 2 * 2 * d.W * 2};
    cuuint32_t box[5] = {BLOCK_K, 1, (cuuint32_t)d.W, 1, (cuuint32_t)(BLOCK_M / d.W)};
    return encode_map(tm, s.ptr, 5, dims, str, box, err, errlen);
}

template <int BN, int KIND, bool GNF = false, int CG = 1>
cudaError_t launch_cfg(const ConvGemmLaunch& l, cudaStream_t s) {
    static int attr_bytes = 0;
    if (l.smem_bytes > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, KIND, GNF, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             l.smem_bytes);
        if (e != cudaSuccess) return e;
        attr_bytes = l.smem_bytes;
    }
    KArgs k;
    k.M = l.M; k.N = l.N; k.num_m_tiles = l.num_m_tiles; k.num_n_tiles = l.num_n_tiles; k.num_tiles = l.num_tiles;
    k.nkb = l.nkb; k.chunks0 = l.chunks0; k.chunks1 = l.chunks1; k.mode = l.mode; k.W = l.W; k.P = l.P;
    k.kh = l.kh; k.kw = l.kw; k.pad = l.pad; k.stages = l.stages; k.Wl_box = l.Wl_box; k.rows_box = l.rows_box;
    k.passes = l.passes;
    k.static_weights = l.static_weights;
    k.slab_bytes = l.slab_bytes; k.slab_dy_bytes = l.slab_dy_bytes; k.res_b_bytes = l.res_b_bytes;
    k.PW = l.PW; k.tiles_per_img = l.tiles_per_img; k.H = l.Hh; k.m_tiles_real = l.m_tiles_real;
    k.epi = l.epi;
    if constexpr (CG == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(l.grid);
        cfg.blockDim = dim3(Cfg<BN, KIND, GNF, CG>::THREADS);
        cfg.dynamicSmemBytes = l.smem_bytes;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, KIND, GNF, CG>, l.tmA0, l.tmA1, l.tmB, l.tmD, l.tmD31, l.tmD30, k);
    } else {
        return launch_pdl(conv_gemm_kernel<BN, KIND, GNF, CG>, dim3(l.grid), dim3(Cfg<BN, KIND, GNF, CG>::THREADS), l.smem_bytes, s, l.tmA0, l.tmA1, l.tmB, l.tmD,
                          l.tmD31, l.tmD30, k);
    }
}

template <int BN, int KIND, bool GNF, int CG = 1>
void size_one(ConvGemmLaunch* l) {
    l->stages = Cfg<BN, KIND, GNF, CG>::stages(l->res_b_bytes);
    l->smem_bytes = Cfg<BN, KIND, GNF, CG>::smem_bytes(l->stages, l->res_b_bytes);
}

template <int BN>
void size_cfg(ConvGemmLaunch* l) {
    if constexpr (BN >= 64) {
        if (l->cg == 2) {
            switch (l->kind) {
                case K_SLAB: if constexpr (BN <= 128) size_one<BN, K_SLAB, false, 2>(l); break;
                case K_SLAB_RES: if constexpr (BN == 64) size_one<BN, K_SLAB_RES, false, 2>(l); break;
                default: size_one<BN, K_GENERAL, false, 2>(l); break;
            }
            return;
        }
    }
    switch (l->kind) {
        case K_SLAB: if (l->gnf) size_one<BN, K_SLAB, true>(l); else size_one<BN, K_SLAB, false>(l); break;
        case K_SLAB_RES: if (l->gnf) size_one<BN, K_SLAB_RES, true>(l); else size_one<BN, K_SLAB_RES, false>(l); break;
        case K_PAD: size_one<BN, K_PAD, false>(l); break;
        case K_DX3: if constexpr (BN == 64) size_one<BN, K_DX3, false>(l); break;
        case K_DX3G: if constexpr (BN == 64) size_one<BN, K_DX3G, false>(l); break;
        default: size_one<BN, K_GENERAL, false>(l); break;
    }
}

}  // namespace

bool conv_gemm_can_fuse_gn(int B, int H, int W, int N, int ksize, ConvMode mode) {
    const int P = H * W;
    return B > 0 && mode == CONV_TAPS && ksize == 3 && P >= BLOCK_M && P % BLOCK_M == 0 && W >= 16 && BLOCK_M % W == 0 &&
           (N == 64 || N == 128) && (P / 32) * (N / 64) <= 256;   // <= 256 partials per (image, group): one batch per lane
}

int conv_gemm_prepare(const ConvGemmDesc& d, int num_sms, ConvGemmLaunch* out, char* err, int errlen) {
    memset(out, 0, sizeof(*out));
    const bool up = d.mode == CONV_UPSAMPLE;
    // the spatial grid the GEMM rows enumerate: output pixels, or (CONV_UPSAMPLE) low-resolution input pixels per phase
    const int Hs = up ? d.H / 2 : d.H, Ws = up ? d.W / 2 : d.W;
    const int P = Hs * Ws;
    if (d.src0.ptr == nullptr || d.src0.C % BLOCK_K != 0 || (d.src1.ptr != nullptr && d.src1.C % BLOCK_K != 0)) {
        snprintf(err, errlen, "conv_gemm: input channels must be multiples of %d (got %d/%d)", BLOCK_K, d.src0.C,
                 d.src1.ptr ? d.src1.C : 0);
        return 1;
    }
    if (up && ((d.H | d.W) & 1)) {
        snprintf(err, errlen, "conv_gemm: upsample output size %dx%d must be even", d.H, d.W);
        return 1;
    }
    if (Ws > BLOCK_M || BLOCK_M % Ws != 0 || (P < BLOCK_M && BLOCK_M % P != 0) || (P >= BLOCK_M && P % BLOCK_M != 0)) {
        snprintf(err, errlen, "conv_gemm: unsupported spatial size %dx%d for a %d-pixel tile", Hs, Ws, BLOCK_M);
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
    out->M = M;
    out->N = d.N;
    out->num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int phases = up ? 4 : 1;
    // Few tiles relative to the machine: prefer a narrower N tile so more SMs get work.
    while (bn > 64 && out->num_m_tiles * (d.N / bn) * phases < num_sms && d.N % (bn / 2) == 0) bn /= 2;
    out->chunks0 = d.src0.C / BLOCK_K;
    out->chunks1 = d.src1.ptr ? d.src1.C / BLOCK_K : 0;
    out->passes = d.wsplit ? 2 : 1;
    out->static_weights = d.static_weights ? 1 : 0;
    const int chunks = (out->chunks0 + out->chunks1) * out->passes;
    int pad_slab_rows = 0;
    uint32_t pad_slab_bytes = 0;
    const int taps = d.mode == CONV_TAPS ? d.ksize * d.ksize : 4;
    out->nkb = taps * chunks;
    // slab path: 3x3 taps over whole image rows, N <= 128 (a 256-wide weight stage would leave one pipeline stage)
    out->kind = K_GENERAL;
    if (d.mode == CONV_TAPS && d.ksize == 3 && P >= BLOCK_M && Ws >= 16 && d.N <= 128) {
        if (bn > 128) bn = 128;
        out->kind = K_SLAB;
        if (bn == 64 && d.N == 64 && out->nkb * 64 * BLOCK_K * 2 <= 144 * 1024) {
            out->kind = K_SLAB_RES;
            out->res_b_bytes = static_cast<uint32_t>(out->nkb) * 64 * BLOCK_K * 2;
            // padded-slab form: ONE box {64 ch, W + 2, rows + 2} per chunk serves all nine taps (3x less L2 -> SM traffic);
            // needs the bias(+statistics)-only epilogue and W in {32, 64}
            // Measured (profiles/r01_notes.md): these N = 64 convs are bound by the tensor core's shared-memory OPERAND
            // bandwidth (~64 B/clk: 64 clk for the 128-row A slice + N/2 for B per K = 16 step), not by L2 -> SM traffic, so
            // the padded form buys no time on B200; it stays opt-in (ConvGemmDesc::pad_mode, HD_CONV_PAD overrides).
            static const int pad_env = [] { const char* v = getenv("HD_CONV_PAD"); return v ? atoi(v) : -1; }();
            const int pad_mode = pad_env >= 0 ? pad_env : d.pad_mode;
            const bool plain = d.epi.res == nullptr && d.epi.film == nullptr && !d.epi.silu && d.epi.out_scale == 1.0f &&
                               d.epi.out_f32 == nullptr && d.epi.gn_gamma == nullptr && d.epi.bias != nullptr;
            if (pad_mode > 0 && plain && (Ws == 64 || Ws == 32) && (Hs * (Ws + 2)) % 32 == 0) {
                const int PW = Ws + 2;
                const int span_rows = (PW - 1 + BLOCK_M - 1) / PW + 1;            // image rows a 128-position tile can touch
                const uint32_t sbytes = static_cast<uint32_t>(span_rows + 2) * PW * BLOCK_K * 2;
                const int avail = SMEM_LIMIT - 1024 - BAR_BYTES - 16384 - static_cast<int>(out->res_b_bytes);
                const int nst = avail / static_cast<int>(PAD_SLAB_CAP_BYTES);
                if (sbytes <= PAD_SLAB_CAP_BYTES && (nst >= 2 || (pad_mode > 1 && nst >= 1))) {
                    out->kind = K_PAD;
                    out->PW = PW;
                    out->Hh = Hs;
                    out->tiles_per_img = (Hs * PW + BLOCK_M - 1) / BLOCK_M;
                    out->num_m_tiles = d.B * out->tiles_per_img;
                    pad_slab_rows = span_rows + 2;
                    pad_slab_bytes = sbytes;
                }
            }
        }
    }
    out->bn = bn;
    out->num_n_tiles = d.N / bn;
    out->m_tiles_real = out->num_m_tiles;
    // CTA pairs (tcgen05 cta_group::2): two CTAs share one weight tile, each fetches half of its rows.  Opt-in: correct
    // (test_conv_gemm_cta_pairs) but measured 1.4-2x SLOWER than the 1-CTA form on this pool's B200s -- the peer's half of B
    // reaches the tensor core at only ~19 B/clk (profiles/r01_notes.md); to be resolved before it becomes the default.
    out->cg = 1;
    {
        static const int cg_env = [] { const char* v = getenv("HD_CONV_CG2"); return v ? atoi(v) : -1; }();
        const int want = cg_env >= 0 ? cg_env : d.cg2_mode;
        const bool plain_kind = out->kind != K_PAD && d.epi.gn_gamma == nullptr && d.epi.out_f32 == nullptr && bn >= 64;
        if (want > 0 && plain_kind && out->num_m_tiles >= 2 && num_sms >= 2) {
            out->cg = 2;
            if (out->kind == K_SLAB_RES) out->res_b_bytes /= 2;
            out->num_m_tiles = (out->num_m_tiles + 1) / 2;          // M-tile PAIRS from here on
        }
    }
    // dx-stacked form of the resident-weight kind (default where it applies): one un-shifted slab per chunk, MMAs of N = 192
    // = 3 dx taps x 64 channels, the one-pixel shifts in the epilogue.  By the measured issue law (N/2 + 43 clk per K = 16 step,
    // profiles/r02_ubench_tcgen05.md) three taps cost 139 clk instead of 3 x 75.  HD_CONV_DX3=0 / ConvGemmDesc::dx3_mode = 0
    // fall back to one MMA per tap.
    {
        static const int dx3_env = [] { const char* v = getenv("HD_CONV_DX3"); return v ? atoi(v) : -1; }();
        const int want = dx3_env >= 0 ? dx3_env : d.dx3_mode;
        if (want > 0 && out->kind == K_SLAB_RES && out->cg == 1 && d.epi.gn_gamma == nullptr && Ws <= 64) {
            // 2 (default): two epilogue groups on alternating tiles, register-direct stores; 1: one group, TMA-store staging
            const bool simple_epi = d.epi.film == nullptr && !d.epi.silu && d.epi.out_lo == nullptr && d.epi.out_f32 == nullptr && d.out != nullptr &&
                                    d.epi.gnres == nullptr;
            const bool fits = Cfg<64, K_DX3G>::stages(out->res_b_bytes) >= 2;    // K = 1152: no room for the second group's boxes
            out->kind = (want >= 2 && simple_epi && fits) ? K_DX3G : K_DX3;
        }
    }
    out->num_tiles = out->num_m_tiles * out->num_n_tiles * phases;
    out->mode = d.mode;
    out->W = Ws;
    out->P = P;
    out->kh = d.mode == CONV_TAPS ? d.ksize : 2;
    out->kw = out->kh;
    out->pad = d.mode == CONV_TAPS ? d.ksize / 2 : 0;
    out->epi = d.epi;
    out->out = d.out;
    out->ldo = d.N;
    out->grid = out->num_tiles < num_sms ? out->num_tiles : num_sms;
    if (out->cg == 2) {
        const int pairs = num_sms / 2;
        out->grid = 2 * (out->num_tiles < pairs ? out->num_tiles : pairs);
    }
    int slab_rows = 0;
    if (out->kind == K_PAD) {
        slab_rows = pad_slab_rows;
        out->slab_bytes = pad_slab_bytes;
    } else if (out->kind != K_GENERAL) {
        slab_rows = BLOCK_M / Ws + 2;
        out->slab_bytes = static_cast<uint32_t>(slab_rows) * Ws * BLOCK_K * 2;
        out->slab_dy_bytes = static_cast<uint32_t>(Ws) * BLOCK_K * 2;
        if (out->slab_bytes > SLAB_CAP_BYTES) {
            snprintf(err, errlen, "conv_gemm: slab of %u bytes exceeds the stage capacity", out->slab_bytes);
            return 1;
        }
    }
    if (up && (d.epi.res != nullptr || d.epi.film != nullptr || d.epi.gn_part != nullptr || d.epi.out_f32 != nullptr)) {
        snprintf(err, errlen, "conv_gemm: the upsample conv supports the bias epilogue only");
        return 1;
    }
    if (d.epi.gnres != nullptr && (d.epi.res == nullptr || d.epi.gn_gamma != nullptr || d.epi.out_f32 != nullptr || up || out->kind == K_PAD)) {
        snprintf(err, errlen, "conv_gemm: the GroupNorm'd residual operand needs the dense bf16 epilogue");
        return 1;
    }
    out->gnf = d.epi.gn_gamma != nullptr ? 1 : 0;
    if (out->gnf && (out->kind == K_GENERAL || !conv_gemm_can_fuse_gn(d.B, d.H, d.W, d.N, d.ksize, d.mode) || d.epi.gn_part == nullptr ||
                     d.epi.gn_counter == nullptr || d.epi.gn_beta == nullptr || d.epi.bias == nullptr)) {
        snprintf(err, errlen, "conv_gemm: this conv cannot run the GroupNorm-fused epilogue");
        return 1;
    }
    if (d.epi.gn_part != nullptr && out->kind != K_PAD && (M % 32 != 0 || bn < 64)) {
        snprintf(err, errlen, "conv_gemm: GroupNorm partials need M %% 32 == 0 and a bf16 output tile");
        return 1;
    }

    const int box_w = out->kind == K_PAD ? out->PW : 0;
    if (encode_activation_map(&out->tmA0, d.src0, d, Hs, Ws, slab_rows, err, errlen, box_w)) return 1;
    if (d.src1.ptr) {
        if (encode_activation_map(&out->tmA1, d.src1, d, Hs, Ws, slab_rows, err, errlen, box_w)) return 1;
    } else {
        out->tmA1 = out->tmA0;
    }
    const cuuint64_t Ktot = (cuuint64_t)out->nkb * BLOCK_K;
    cuuint64_t wd[2] = {Ktot, (cuuint64_t)d.N * phases};
    cuuint64_t ws[1] = {Ktot * 2};
    cuuint32_t wb[2] = {BLOCK_K, (cuuint32_t)(bn / out->cg)};
    if (encode_map(&out->tmB, d.weight, 2, wd, ws, wb, err, errlen)) return 1;
    if (d.epi.out_f32 == nullptr) {
        if (bn < 64 || d.out == nullptr) {
            snprintf(err, errlen, "conv_gemm: bf16 output needs an N tile >= 64 and an output buffer");
            return 1;
        }
        if (out->kind == K_PAD) {
            // dense [N, M] output; a warp block stores its 32, 31 or 30 valid (non-halo) rows: one tensor map per box height
            cuuint64_t od[2] = {(cuuint64_t)d.N, (cuuint64_t)M};
            cuuint64_t os[1] = {(cuuint64_t)d.N * 2};
            cuuint32_t ob[2] = {32, 32};
            if (encode_map(&out->tmD, d.out, 2, od, os, ob, err, errlen, 64)) return 1;
            ob[1] = 31;
            if (encode_map(&out->tmD31, d.out, 2, od, os, ob, err, errlen, 64)) return 1;
            ob[1] = 30;
            if (encode_map(&out->tmD30, d.out, 2, od, os, ob, err, errlen, 64)) return 1;
        } else if (up) {
            // output [B, H, W, N] viewed as [N, pb, W/2, pa, B*H/2]; a warp stores 32 consecutive low-res pixels of one phase
            const cuuint64_t N2 = (cuuint64_t)d.N * 2;
            out->Wl_box = Ws;
            out->rows_box = Ws >= 32 ? 1 : 32 / Ws;
            cuuint64_t od[5] = {(cuuint64_t)d.N, 2, (cuuint64_t)Ws, 2, (cuuint64_t)d.B * Hs};
            cuuint64_t os[4] = {N2, 2 * N2, (cuuint64_t)d.W * N2, 2 * (cuuint64_t)d.W * N2};
            cuuint32_t ob[5] = {32, 1, (cuuint32_t)(Ws >= 32 ? 32 : Ws), 1, (cuuint32_t)out->rows_box};
            if (encode_map(&out->tmD, d.out, 5, od, os, ob, err, errlen, 64)) return 1;
        } else {
            cuuint64_t od[2] = {(cuuint64_t)d.N, (cuuint64_t)M};
            cuuint64_t os[1] = {(cuuint64_t)d.N * 2};
            cuuint32_t ob[2] = {32, 32};     // one epilogue warp: 32 rows x 32 columns, 64-byte rows, 64B swizzle
            if (encode_map(&out->tmD, d.out, 2, od, os, ob, err, errlen, 64)) return 1;
            if (d.epi.out_lo != nullptr && encode_map(&out->tmD31, d.epi.out_lo, 2, od, os, ob, err, errlen, 64)) return 1;

        }
        if (d.epi.out_lo != nullptr && (out->kind == K_PAD || up || d.epi.gn_gamma != nullptr)) {
            snprintf(err, errlen, "conv_gemm: the hi + lo output is built for the plain dense epilogue");
            return 1;
        }
    } else {
        if (bn >= 64 || up) {
            snprintf(err, errlen, "conv_gemm: the fp32 head path is built for N <= 48 (padded to 16-column tiles), no upsample");
            return 1;
        }
        out->tmD = out->tmB;   // unused by the fp32 head path
    }
    switch (bn) {
        case 256: size_cfg<256>(out); break;
        case 128: size_cfg<128>(out); break;
        case 64: size_cfg<64>(out); break;
        default: size_cfg<16>(out); break;
    }
    if (out->stages < (out->kind == K_PAD ? 1 : 2)) {
        snprintf(err, errlen, "conv_gemm: only %d pipeline stage(s) fit in shared memory", out->stages);
        return 1;
    }
    return 0;
}

cudaError_t conv_gemm_run(const ConvGemmLaunch& l, cudaStream_t s) {
    if (l.cg == 2) {
        switch (l.kind) {
            case K_GENERAL:
                switch (l.bn) {
                    case 256: return launch_cfg<256, K_GENERAL, false, 2>(l, s);
                    case 128: return launch_cfg<128, K_GENERAL, false, 2>(l, s);
                    case 64: return launch_cfg<64, K_GENERAL, false, 2>(l, s);
                    default: return cudaErrorInvalidValue;
                }
            case K_SLAB:
                switch (l.bn) {
                    case 128: return launch_cfg<128, K_SLAB, false, 2>(l, s);
                    case 64: return launch_cfg<64, K_SLAB, false, 2>(l, s);
                    default: return cudaErrorInvalidValue;
                }
            case K_SLAB_RES:
                return l.bn == 64 ? launch_cfg<64, K_SLAB_RES, false, 2>(l, s) : cudaErrorInvalidValue;
            default: return cudaErrorInvalidValue;
        }
    }
    switch (l.kind) {
        case K_GENERAL:
            switch (l.bn) {
                case 256: return l.gnf ? cudaErrorInvalidValue : launch_cfg<256, K_GENERAL>(l, s);
                case 128: return launch_cfg<128, K_GENERAL>(l, s);
                case 64: return launch_cfg<64, K_GENERAL>(l, s);
                case 16: return launch_cfg<16, K_GENERAL>(l, s);
                default: return cudaErrorInvalidValue;
            }
        case K_SLAB:
            switch (l.bn) {
                case 128: return l.gnf ? launch_cfg<128, K_SLAB, true>(l, s) : launch_cfg<128, K_SLAB>(l, s);
                case 64: return l.gnf ? launch_cfg<64, K_SLAB, true>(l, s) : launch_cfg<64, K_SLAB>(l, s);
                case 16: return l.gnf ? cudaErrorInvalidValue : launch_cfg<16, K_SLAB>(l, s);
                default: return cudaErrorInvalidValue;
            }
        case K_PAD:
            return (l.bn == 64 && !l.gnf) ? launch_cfg<64, K_PAD>(l, s) : cudaErrorInvalidValue;
        case K_SLAB_RES:
            if (l.bn != 64) return cudaErrorInvalidValue;
            return l.gnf ? launch_cfg<64, K_SLAB_RES, true>(l, s) : launch_cfg<64, K_SLAB_RES>(l, s);
        case K_DX3:
            return (l.bn == 64 && !l.gnf) ? launch_cfg<64, K_DX3>(l, s) : cudaErrorInvalidValue;
        case K_DX3G:
            return (l.bn == 64 && !l.gnf) ? launch_cfg<64, K_DX3G>(l, s) : cudaErrorInvalidValue;
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hd
