// Weight gradient of a 256 -> 256 3x3 "same" convolution on tcgen05 (HiCEDRN training):
//
//     dW[co, ci, ky, kx] = sum_{b, y, x} dY[b, y, x, co] * X[b, y + ky - 1, x + kx - 1, ci]
//
// i.e. nine GEMMs (one per tap) with M = Cout, N = Cin and K = B * H * W pixels.  Both operands are read straight from the
// NHWC bf16 activations as MN-MAJOR tiles: a K block is one image row of 64 pixels, and one 5-D TMA box
// {64 ch, 64 px, 1 row, 1 img, 4 chunks} lands in shared memory as [chunk][px][64 ch] -- K rows (pixels) of 128 B, 8-pixel
// swizzle atoms 1 KiB apart (SBO), 64-channel groups 8 KiB apart (LBO): the canonical MN-major SWIZZLE_128B operand of
// tcgen05.mma (a_major = b_major = 1 in the instruction descriptor).  The tap's (dy, dx) shift is a TMA coordinate offset
// of the H / W dimensions whose out-of-range pixels the TMA unit zero-fills (the conv's zero padding).  (A first version
// read planar [B, C, H, W] copies as K-major tiles; there dx falls on the innermost dimension, where a TMA box must start
// on a 16-byte boundary -- an illegal-instruction fault -- and the transposes cost more than the GEMM.)
//
// One CTA = one (tap, K split): the whole dY row tile (both 128-channel halves) and the whole X row tile per stage (64 KiB,
// 3 stages), two M = 128 x N = 256 fp32 accumulators = all 512 TMEM columns, 8 MMAs per stage.  Warp 0 produces (TMA), warp 1
// issues, warps 2-5 drain the accumulators into a [split][tap][co][ci] fp32 workspace that wgrad_reduce_kernel sums in a
// fixed order (deterministic; no atomics) into the reference's [Cout, Cin, 3, 3] layout.
#include <cstdlib>
#include "kernels.h"
#include "ptx.cuh"

#include <cstdio>

namespace hd {
namespace {

constexpr int WG_C = 256;
constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 3;
constexpr uint32_t WG_A_BYTES = WG_C * 64 * 2;   // dY row tile: [4 chunks][64 px][64 co]
constexpr uint32_t WG_B_BYTES = WG_C * 64 * 2;   // X  row tile: [4 chunks][64 px][64 ci]
constexpr uint32_t WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

struct WgradKArgs {
    int kb_total;     // B * H K blocks (image rows)
    int nsplit;
    int H;
    float* part;      // [nsplit][9][256][256]
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradKArgs a) {
    extern __shared__ uint8_t wg_smem_raw[];
    uint8_t* smem = wg_smem_raw + ((1024u - (ptx::smem_u32(wg_smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + WG_STAGES;
    uint64_t* tfull_bar = empty_bar + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap = blockIdx.x % 9, split = blockIdx.x / 9;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int kb0 = static_cast<int>(static_cast<long long>(a.kb_total) * split / a.nsplit);
    const int kb1 = static_cast<int>(static_cast<long long>(a.kb_total) * (split + 1) / a.nsplit);

    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmG); ptx::prefetch_tmap(&tmX); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int b = kb / a.H, y = kb - b * a.H;
            if (y + dy < 0 || y + dy >= a.H) continue;        // the shifted row lies in the zero padding
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (ptx::elect_one()) {
                uint8_t* sA = smem + stage * WG_STAGE_BYTES;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
                ptx::tma_load_5d(sA, &tmG, &full_bar[stage], 0, 0, y, b, 0);
                ptx::tma_load_5d(sA + WG_A_BYTES, &tmX, &full_bar[stage], 0, dx, y + dy, b, 0);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = ptx::make_idesc_bf16_mn(128, WG_C);
        constexpr uint32_t K16_STEP = 2048u >> 4;   // 16 pixels = two 1 KiB swizzle atoms
        const uint64_t desc0 = ptx::make_mnmajor_sw128_desc(ptx::smem_u32(smem), 64 * 128, 1024);
        int stage = 0;
        uint32_t phase = 0;
        bool first = true;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int y = kb % a.H;
            if (y + dy < 0 || y + dy >= a.H) continue;
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t da = desc0 + static_cast<uint64_t>((stage * WG_STAGE_BYTES) >> 4);
                const uint64_t db = da + static_cast<uint64_t>(WG_A_BYTES >> 4);
#pragma unroll
                for (int mh = 0; mh < 2; ++mh) {
                    const uint64_t dam = da + static_cast<uint64_t>((mh * (WG_A_BYTES / 2)) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16(tmem_base + mh * WG_C, dam + K16_STEP * k, db + K16_STEP * k, idesc, (first && k == 0) ? 0u : 1u);
                }
                ptx::umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            first = false;
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(tfull_bar);
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue: TMEM -> fp32 workspace
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        bool any = false;
        for (int kb = kb0; kb < kb1; ++kb) { const int y = kb % a.H; if (y + dy >= 0 && y + dy < a.H) { any = true; break; } }
        ptx::mbar_wait(tfull_bar, 0);
        ptx::tc_fence_after();
        float* out = a.part + (static_cast<size_t>(split) * 9 + tap) * WG_C * WG_C;
#pragma unroll 1
        for (int mh = 0; mh < 2; ++mh) {
            float* orow = out + static_cast<size_t>(mh * 128 + q * 32 + lane) * WG_C;
#pragma unroll 1
            for (int c = 0; c < WG_C; c += 32) {
                uint32_t v[32];
                if (any) {
                    ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mh * WG_C + c, v);
                    ptx::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(orow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// dW[(co * 256 + ci) * 9 + tap] (+)= scale * sum_split part[split][tap][co][ci]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, float scale, int accumulate, float* __restrict__ dw) {
    __shared__ float s_t[9][257];
    const int i0 = blockIdx.x * 256;                 // 256 consecutive (co, ci) pairs
    for (int tap = 0; tap < 9; ++tap) {
        float t = 0.f;
        for (int sp = 0; sp < nsplit; ++sp)
            t += __ldg(part + (static_cast<size_t>(sp) * 9 + tap) * WG_C * WG_C + i0 + threadIdx.x);
        s_t[tap][threadIdx.x] = t * scale;
    }
    __syncthreads();
    float* o = dw + static_cast<size_t>(i0) * 9;
    for (int j = threadIdx.x; j < 256 * 9; j += 256) {
        const float v = s_t[j % 9][j / 9];
        o[j] = accumulate ? o[j] + v : v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// General form (the Unet's convs): any Cout / Cin that are multiples of 64, 1x1 or 3x3, image width 8 .. 64.  One CTA =
// one (tap, 128-channel Cout tile, <= 256-channel Cin tile, K split) with ONE accumulator; a K block is 64 pixels = 64 / W
// image rows (box {64 ch, W, 64 / W rows, 1 img, chunks}); Cout = 64 rides as half of a 128-row tile whose upper chunk the
// TMA unit zero-fills (the chunk coordinate is out of range).
// ------------------------------------------------------------------------------------------------------------------
constexpr int WGG_STAGES = 4;
constexpr uint32_t WGG_CHUNK = 64 * 128;            // one 64-channel chunk of a 64-pixel K block
constexpr uint32_t WGG_A_BYTES = 2 * WGG_CHUNK;
constexpr uint32_t WGG_STAGE_MAX = WGG_A_BYTES + 4 * WGG_CHUNK;
constexpr int WGG_SMEM = WGG_STAGES * WGG_STAGE_MAX + 1024 + 256;

struct WgradGenKArgs {
    int B, H, W, rows_kb;        // rows_kb = 64 / W image rows per K block
    int Cout, Cin, Nt;           // Nt = Cin tile (64, 128 or 256)
    int taps, mtiles, ntiles, nsplit;
    int kb_total;                // B * H / rows_kb
    float* part;                 // [nsplit][taps][Cout][Cin]
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_general_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradGenKArgs a) {
    extern __shared__ uint8_t wg_smem_raw[];
    uint8_t* smem = wg_smem_raw + ((1024u - (ptx::smem_u32(wg_smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WGG_STAGES * WGG_STAGE_MAX);
    uint64_t* empty_bar = full_bar + WGG_STAGES;
    uint64_t* tfull_bar = empty_bar + WGG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int id = blockIdx.x;
    const int tap = id % a.taps; id /= a.taps;
    const int mt = id % a.mtiles; id /= a.mtiles;
    const int nt = id % a.ntiles; id /= a.ntiles;
    const int split = id;
    const int dy = a.taps == 9 ? tap / 3 - 1 : 0, dx = a.taps == 9 ? tap % 3 - 1 : 0;
    const int kb0 = static_cast<int>(static_cast<long long>(a.kb_total) * split / a.nsplit);
    const int kb1 = static_cast<int>(static_cast<long long>(a.kb_total) * (split + 1) / a.nsplit);
    const uint32_t stage_bytes = WGG_A_BYTES + (a.Nt / 64) * WGG_CHUNK;
    const int kb_per_img = a.H / a.rows_kb;
    const uint32_t tmem_cols = a.Nt < 32 ? 32u : static_cast<uint32_t>(a.Nt);

    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmG); ptx::prefetch_tmap(&tmX); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, tmem_cols);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < WGG_STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int b = kb / kb_per_img, y0 = (kb - b * kb_per_img) * a.rows_kb;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (ptx::elect_one()) {
                uint8_t* sA = smem + stage * WGG_STAGE_MAX;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                ptx::tma_load_5d(sA, &tmG, &full_bar[stage], 0, 0, y0, b, mt * 2);
                ptx::tma_load_5d(sA + WGG_A_BYTES, &tmX, &full_bar[stage], 0, dx, y0 + dy, b, nt * (a.Nt / 64));
            }
            __syncwarp();
            if (++stage == WGG_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::make_idesc_bf16_mn(128, static_cast<uint32_t>(a.Nt));
        constexpr uint32_t K16_STEP = 2048u >> 4;
        const uint64_t desc0 = ptx::make_mnmajor_sw128_desc(ptx::smem_u32(smem), WGG_CHUNK, 1024);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t da = desc0 + static_cast<uint64_t>((stage * WGG_STAGE_MAX) >> 4);
                const uint64_t db = da + static_cast<uint64_t>(WGG_A_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_bf16(tmem_base, da + K16_STEP * k, db + K16_STEP * k, idesc, (kb == kb0 && k == 0) ? 0u : 1u);
                ptx::umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == WGG_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(tfull_bar);
        __syncwarp();
    } else {
        const int q = warp & 3;
        ptx::mbar_wait(tfull_bar, 0);
        ptx::tc_fence_after();
        const int co = mt * 128 + q * 32 + lane;
        float* orow = a.part + ((static_cast<size_t>(split) * a.taps + tap) * a.Cout + co) * a.Cin + nt * a.Nt;
#pragma unroll 1
        for (int c = 0; c < a.Nt; c += 32) {
            uint32_t v[32];
            ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
            ptx::tmem_ld_wait();
            if (co < a.Cout) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(orow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Filter-row form of the general kernel for 3x3 convs on 64-pixel-wide images with a Cin tile <= 128: one CTA = one
// (filter row dy, Cout tile, Cin tile, K split) with THREE accumulators (dx = -1, 0, +1).  A K block is one image row: the G
// tile is loaded once and X arrives as ONE halo'd slab {64 ch, 66 pixels starting at column -1} per 64-channel chunk; the
// three taps are the same slab read from K row 0, 1, 2 (descriptor start + 128 bytes per pixel -- the tensor core derives the
// swizzle phase from absolute address bits, see conv_gemm.cu).  L2 -> SM operand traffic per tap drops from (G + X) to
// (G + 1.03 X) / 3: the per-tap kernel is operand-fetch bound at N = 64 (250 TFLOP/s).
// ------------------------------------------------------------------------------------------------------------------
// (W = 32 works the same way with a 32-pixel K block: slab of 34 pixel rows, two K = 16 steps.)

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_rows_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmXs, const WgradGenKArgs a) {
    extern __shared__ uint8_t wg_smem_raw[];
    uint8_t* smem = wg_smem_raw + ((1024u - (ptx::smem_u32(wg_smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WGG_STAGES * WGG_STAGE_MAX);
    uint64_t* empty_bar = full_bar + WGG_STAGES;
    uint64_t* tfull_bar = empty_bar + WGG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int id = blockIdx.x;
    const int trow = id % 3; id /= 3;
    const int mt = id % a.mtiles; id /= a.mtiles;
    const int nt = id % a.ntiles; id /= a.ntiles;
    const int split = id;
    const int dy = trow - 1;
    const int kb0 = static_cast<int>(static_cast<long long>(a.kb_total) * split / a.nsplit);
    const int kb1 = static_cast<int>(static_cast<long long>(a.kb_total) * (split + 1) / a.nsplit);
    const int nchunks = a.Nt / 64;
    const uint32_t a_chunk = static_cast<uint32_t>(a.W) * 128u;               // one 64-channel chunk of the G tile (W pixel rows)
    const uint32_t slab = ((static_cast<uint32_t>(a.W) + 2u + 7u) & ~7u) * 128u;   // W + 2 pixel rows padded to whole 8-row swizzle atoms
    const uint32_t stage_bytes = 2u * a_chunk + nchunks * (static_cast<uint32_t>(a.W) + 2u) * 128u;
    const uint32_t tmem_cols = 3 * a.Nt <= 256 ? 256u : 512u;

    if (warp == 0) {
        if (lane == 0) { ptx::prefetch_tmap(&tmG); ptx::prefetch_tmap(&tmXs); }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, tmem_cols);
        ptx::tmem_relinquish();
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < WGG_STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int b = kb / a.H, y0 = kb - b * a.H;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (ptx::elect_one()) {
                uint8_t* sA = smem + stage * WGG_STAGE_MAX;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                ptx::tma_load_5d(sA, &tmG, &full_bar[stage], 0, 0, y0, b, mt * 2);
                for (int c = 0; c < nchunks; ++c)
                    ptx::tma_load_5d(sA + WGG_A_BYTES + c * slab, &tmXs, &full_bar[stage], 0, -1, y0 + dy, b, nt * nchunks + c);
            }
            __syncwarp();
            if (++stage == WGG_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::make_idesc_bf16_mn(128, static_cast<uint32_t>(a.Nt));
        constexpr uint32_t K16_STEP = 2048u >> 4;
        const uint64_t descA0 = ptx::make_mnmajor_sw128_desc(ptx::smem_u32(smem), a_chunk, 1024);
        const uint64_t descB0 = ptx::make_mnmajor_sw128_desc(ptx::smem_u32(smem) + WGG_A_BYTES, slab, 1024);
        const int ksteps = a.W / 16;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t da = descA0 + static_cast<uint64_t>((stage * WGG_STAGE_MAX) >> 4);
                const uint64_t db = descB0 + static_cast<uint64_t>((stage * WGG_STAGE_MAX) >> 4);
#pragma unroll
                for (int t = 0; t < 3; ++t)            // x + dx is slab K row (x + dx + 1): tap t = dx + 1 starts t pixels (128 B each) in
                    for (int k = 0; k < ksteps; ++k)
                        ptx::umma_bf16(tmem_base + static_cast<uint32_t>(t * a.Nt), da + K16_STEP * k, db + (128u >> 4) * t + K16_STEP * k,
                                       idesc, (kb == kb0 && k == 0) ? 0u : 1u);
                ptx::umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == WGG_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(tfull_bar);
        __syncwarp();
    } else {
        const int q = warp & 3;
        ptx::mbar_wait(tfull_bar, 0);
        ptx::tc_fence_after();
        const int co = mt * 128 + q * 32 + lane;
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
            float* orow = a.part + ((static_cast<size_t>(split) * 9 + trow * 3 + t) * a.Cout + co) * a.Cin + nt * a.Nt;
#pragma unroll 1
            for (int c = 0; c < a.Nt; c += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * a.Nt + c, v);
                ptx::tmem_ld_wait();
                if (co < a.Cout) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(orow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// dw[(co * cin_total + ci0 + ci) * taps + tap] (+)= scale * sum_split part[split][tap][co][ci]; one thread per (co, ci)
// (1x1 convs: reads and writes are both contiguous in ci)
__global__ void __launch_bounds__(256)
wgrad_general_reduce_kernel(const float* __restrict__ part, int nsplit, int taps, int Cout, int Cin, int cin_total, int ci0,
                            float scale, int accumulate, float* __restrict__ dw) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Cout * Cin) return;
    const int co = i / Cin, ci = i - co * Cin;
    const size_t plane = static_cast<size_t>(Cout) * Cin;
    float* o = dw + (static_cast<size_t>(co) * cin_total + ci0 + ci) * taps;
    for (int tap = 0; tap < taps; ++tap) {
        float t = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) t += __ldg(part + (static_cast<size_t>(sp) * taps + tap) * plane + i);
        o[tap] = accumulate ? fmaf(scale, t, o[tap]) : scale * t;
    }
}

// 3x3 form: one CTA per (co, 32 input channels), thread (tap, ci): nine times the parallelism of the kernel above (the split
// loop is a chain of dependent loads per thread), 128-byte reads per warp, and the (ci, tap) transpose goes through shared
// memory so the 288 outputs of the CTA -- contiguous in dw -- are written in order.  Same summation order: same bits.
__global__ void __launch_bounds__(288)
wgrad_general_reduce9_kernel(const float* __restrict__ part, int nsplit, int Cout, int Cin, int cin_total, int ci0, float scale,
                             int accumulate, float* __restrict__ dw) {
    __shared__ float s_t[32][9 + 1];
    const int nci = Cin / 32;
    const int co = blockIdx.x / nci, c0 = (blockIdx.x - co * nci) * 32;
    const int tap = threadIdx.x >> 5, cl = threadIdx.x & 31;
    const size_t plane = static_cast<size_t>(Cout) * Cin;
    const float* src = part + static_cast<size_t>(tap) * plane + static_cast<size_t>(co) * Cin + c0 + cl;
    float t = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) t += __ldg(src + static_cast<size_t>(sp) * 9 * plane);
    s_t[cl][tap] = t;
    __syncthreads();
    const int j = threadIdx.x;                  // output (ci = j / 9, tap = j % 9) of this CTA's contiguous 288-float span
    float* o = dw + (static_cast<size_t>(co) * cin_total + ci0 + c0) * 9 + j;
    const float v = s_t[j / 9][j % 9];
    *o = accumulate ? fmaf(scale, v, *o) : scale * v;
}

}  // namespace

int wgrad_prepare(const bf16* g, const bf16* x, int B, int H, int W, int C, int nsplit, float* part, WgradLaunch* out, char* err,
                       int errlen) {
    if (C != WG_C || W != 64) {
        snprintf(err, errlen, "wgrad: built for 256 channels and 64-pixel rows (got C = %d, W = %d)", C, W);
        return 1;
    }
    if (nsplit < 1 || static_cast<long long>(B) * H / nsplit < 4) {
        snprintf(err, errlen, "wgrad: %d K splits leave fewer than 4 image rows per CTA", nsplit);
        return 1;
    }
    // [B, H, W, C] viewed as {64 ch, W, H, B, C / 64 chunks}
    cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)C / 64};
    cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, 128};
    cuuint32_t box[5] = {64, 64, 1, 1, (cuuint32_t)C / 64};
    if (encode_tmap_bf16(&out->tmG, g, 5, dims, str, box, err, errlen)) return 1;
    if (encode_tmap_bf16(&out->tmX, x, 5, dims, str, box, err, errlen)) return 1;
    out->kb_total = B * H;
    out->H = H;
    out->nsplit = nsplit;
    out->part = part;
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) { snprintf(err, errlen, "wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

cudaError_t wgrad_run(const WgradLaunch& l, cudaStream_t s) {
    WgradKArgs a;
    a.kb_total = l.kb_total; a.nsplit = l.nsplit; a.H = l.H; a.part = l.part;

    wgrad_kernel<<<9 * l.nsplit, WG_THREADS, WG_SMEM, s>>>(l.tmG, l.tmX, a);
    return cudaGetLastError();
}

int wgrad_general_prepare(const bf16* g, const bf16* x, int B, int H, int W, int Cout, int Cin, int ksize, int num_sms,
                          WgradGenLaunch* out, char* err, int errlen) {
    if ((ksize != 1 && ksize != 3) || Cout % 64 != 0 || Cin % 64 != 0 || Cout < 64 || Cin < 64 ||
        (W != 8 && W != 16 && W != 32 && W != 64) || H % (64 / W) != 0) {
        snprintf(err, errlen, "wgrad: unsupported shape (B %d, H %d, W %d, Cout %d, Cin %d, k %d)", B, H, W, Cout, Cin, ksize);
        return 1;
    }
    const int rows = 64 / W;
    const int Nt = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
    cuuint64_t gd[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)Cout / 64};
    cuuint64_t gs[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2, 128};
    cuuint32_t gb[5] = {64, (cuuint32_t)W, (cuuint32_t)rows, 1, 2};
    if (encode_tmap_bf16(&out->tmG, g, 5, gd, gs, gb, err, errlen)) return 1;
    cuuint64_t xd[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)Cin / 64};
    cuuint64_t xs[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2, 128};
    cuuint32_t xb[5] = {64, (cuuint32_t)W, (cuuint32_t)rows, 1, (cuuint32_t)Nt / 64};
    if (encode_tmap_bf16(&out->tmX, x, 5, xd, xs, xb, err, errlen)) return 1;
    out->B = B; out->H = H; out->W = W; out->rows_kb = rows; out->Cout = Cout; out->Cin = Cin; out->Nt = Nt;
    out->taps = ksize * ksize;
    out->mtiles = (Cout + 127) / 128;
    out->ntiles = Cin / Nt;
    out->kb_total = B * H / rows;
    static const int rows_min_w = [] {        // HD_WGRAD_ROWS=0: per-tap form everywhere; =64: filter-row form on 64-pixel-wide images only
        const char* v = getenv("HD_WGRAD_ROWS");
        if (!v) return 32;
        const int x = atoi(v);
        return x == 0 ? 1 << 30 : (x == 64 ? 64 : 32);
    }();
    out->rowmode = (ksize == 3 && (W == 64 || W == 32) && W >= rows_min_w && Nt <= 128) ? 1 : 0;
    int ns;
    if (out->rowmode) {
        // filter-row CTAs (three taps each): one full wave
        cuuint32_t sb[5] = {64, (cuuint32_t)W + 2, 1, 1, 1};
        if (encode_tmap_bf16(&out->tmXs, x, 5, xd, xs, sb, err, errlen)) return 1;
        cuuint32_t gr[5] = {64, (cuuint32_t)W, 1, 1, 2};              // a K block is ONE image row here
        if (encode_tmap_bf16(&out->tmG, g, 5, gd, gs, gr, err, errlen)) return 1;
        out->rows_kb = 1;
        out->kb_total = B * H;
        const int units = 3 * out->mtiles * out->ntiles;
        ns = num_sms / units;
        if (ns > out->kb_total / 32) ns = out->kb_total / 32;
        if (ns < 1) ns = 1;
        if (ns > 64) ns = 64;
    } else {
        static const int min_kb = [] { const char* v = getenv("HD_WGRAD_MINKB"); const int x = v ? atoi(v) : 0; return x > 0 ? x : 16; }();
        const int units = out->taps * out->mtiles * out->ntiles;
        ns = (2 * num_sms + units - 1) / units;             // about two waves of CTAs
        if (ns > out->kb_total / min_kb) ns = out->kb_total / min_kb;   // >= 16 K blocks per CTA: below that the fp32 partials (written, then re-read
        if (ns < 1) ns = 1;                                      // by the reduction) cost more than the parallelism buys on the low-resolution levels
        if (ns > 32) ns = 32;
    }
    out->nsplit = ns;
    out->part = nullptr;
    cudaError_t e = cudaFuncSetAttribute(wgrad_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGG_SMEM);
    if (e != cudaSuccess) { snprintf(err, errlen, "wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

size_t wgrad_general_part_bytes(const WgradGenLaunch& l) {
    return static_cast<size_t>(l.nsplit) * l.taps * l.Cout * l.Cin * sizeof(float);
}

cudaError_t wgrad_general_run(const WgradGenLaunch& l, cudaStream_t s) {
    WgradGenKArgs a;
    a.B = l.B; a.H = l.H; a.W = l.W; a.rows_kb = l.rows_kb; a.Cout = l.Cout; a.Cin = l.Cin; a.Nt = l.Nt; a.taps = l.taps;
    a.mtiles = l.mtiles; a.ntiles = l.ntiles; a.nsplit = l.nsplit; a.kb_total = l.kb_total; a.part = l.part;
    if (l.rowmode) wgrad_rows_kernel<<<3 * l.mtiles * l.ntiles * l.nsplit, WG_THREADS, WGG_SMEM, s>>>(l.tmG, l.tmXs, a);
    else wgrad_general_kernel<<<l.taps * l.mtiles * l.ntiles * l.nsplit, WG_THREADS, WGG_SMEM, s>>>(l.tmG, l.tmX, a);
    return cudaGetLastError();
}

cudaError_t wgrad_general_reduce_run(const WgradGenLaunch& l, int cin_total, int ci0, float scale, int accumulate, float* dw,
                                     cudaStream_t s) {
    const int n = l.Cout * l.Cin;
    if (l.taps == 9 && l.Cin % 32 == 0)
        wgrad_general_reduce9_kernel<<<n / 32, 288, 0, s>>>(l.part, l.nsplit, l.Cout, l.Cin, cin_total, ci0, scale, accumulate, dw);
    else
        wgrad_general_reduce_kernel<<<(n + 255) / 256, 256, 0, s>>>(l.part, l.nsplit, l.taps, l.Cout, l.Cin, cin_total, ci0, scale, accumulate, dw);
    return cudaGetLastError();
}

cudaError_t wgrad_reduce_run(const float* part, int nsplit, float scale, int accumulate, float* dw, cudaStream_t s) {
    wgrad_reduce_kernel<<<WG_C * WG_C / 256, 256, 0, s>>>(part, nsplit, scale, accumulate, dw);
    return cudaGetLastError();
}

size_t wgrad_part_bytes(int nsplit) { return static_cast<size_t>(nsplit) * 9 * WG_C * WG_C * sizeof(float); }

}  // namespace hd
