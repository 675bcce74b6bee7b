// Fused normalisation kernels (HBM-bandwidth bound; one read + one write of the activation).
//
// groupnorm_film_silu: nn.GroupNorm(8, C, eps=1e-5) -> x*(scale+1)+shift -> SiLU [-> + postadd] [-> + residual]
//   reference: Block.forward   /root/reference/src/hicdiff_condition.py:162-171
//              ResnetBlock     :185-197 (the "+ res_conv(x)" is the optional residual operand here)
//              SR3 ResnetBlock /root/reference/src/hicdiff_sr3.py:246-251 (additive noise embedding = postadd)
//   The statistics come for free from the producing conv: its epilogue (conv_gemm.cu) reduces the fp32 accumulators of
//   every 32-row warp block to a (sum, M2-about-the-block-mean) pair per 8-channel piece with warp shuffles and writes
//   them to a small [M/32][C/8] buffer (no atomics).  This kernel merges an image's partials with Chan's formula (stable, fixed
//   order, so deterministic) and then streams the activation once: read, normalise, modulate, SiLU, (+residual), write.
//   (A first version owned one image per thread-block cluster with DSMEM reductions; ncu showed it latency-bound on
//   cluster syncs and co-scheduling -- see profiles/r01_notes.md.)
//
// channel_layernorm: LayerNorm over channels with gain g (eps 1e-5), optional residual add and optional
//   nearest-neighbour 2x upsample on the way out.
//   reference: LayerNorm :99-108, PreNorm :110-118, Residual :64-70, Upsample's nn.Upsample :74
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_GROUPS = 8;
// 16-byte chunks in flight per thread and stream (template parameter CH of the apply kernel): 4 with a residual stream, 8 without --
// the 2-stream form then keeps as many loads in flight as the 3-stream form (which ran at 91 % of the copy bandwidth against 76 %:
// at 512 threads per SM, 4 chunks are 32 KiB in flight per SM, below what the loaded HBM latency asks for at 6.5 TB/s)
constexpr int GN_CHUNKS_PER_THREAD = 4;
constexpr int GN_SLAB_CHUNKS = GN_THREADS * GN_CHUNKS_PER_THREAD;        // 16 KiB of the image per CTA and iteration (CH = 4)

// x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE SFU op (tanh.approx, rel. error 2^-11) instead of ex2 + rcp.  ncu showed the
// apply kernel at ~50% issue utilisation with 2 SFU ops per element (134 M per 128 MiB tensor = 29 us of SFU time alone); the
// absolute error |h| * 2^-11 stays below the bf16 output rounding for the |x| <~ 8 that follow a GroupNorm.
__device__ __forceinline__ float silu_f(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// Streaming apply, persistent form: each CTA owns a contiguous run of 16 KiB slabs (so it crosses at most a few image
// boundaries), merges an image's per-warp-block partials when it enters the image (Chan's parallel variance formula, fixed
// order -> every CTA derives bit-identical statistics), and keeps the next slab's loads in flight while it normalises /
// modulates / activates the current one.  (The one-CTA-per-64-KiB version paid the statistics prologue and a partial
// last wave on every launch: 3.9 TB/s at batch 256.)
// exact form for the "bf16w2" precision mode: ex2 + rcp (each ~2^-22).  tanh.approx's 2^-11 is a FIXED function error -- like a
// rounded weight it perturbs every step of a chain the same way (measured: it alone kept the T = 1000 PSNR 1.4e-2 dB off).
__device__ __forceinline__ float silu_exact(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

template <bool HAS_RES, bool HAS_POST, bool EXACT = false, int CH = GN_CHUNKS_PER_THREAD>   // compile out the residual stream and the SR3 post-add where a launch has none (issue-bound kernel)
__global__ void __launch_bounds__(GN_THREADS, 2)
groupnorm_apply_kernel(const GroupNormArgs a, const int slabs_per_img, const int total_slabs, const int slabs_per_cta) {
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    const int s0 = blockIdx.x * slabs_per_cta;
    const int s1 = min(s0 + slabs_per_cta, total_slabs);
    if (s0 >= s1) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int chunks_per_pixel = a.C / 8;
    const int cpg = a.C / GN_GROUPS;
    const uint4* xin = reinterpret_cast<const uint4*>(a.x);
    const uint4* rin = HAS_RES ? reinterpret_cast<const uint4*>(a.res) : nullptr;
    uint4* yout = reinterpret_cast<uint4*>(a.y);
    const int my_cp = tid % chunks_per_pixel;                // fixed per thread: GN_THREADS % chunks_per_pixel == 0
    const int my_c = my_cp * 8;
    const int my_g = my_c / cpg;

    uint4 u[CH], r[CH];
    auto load_slab = [&](int s, uint4 (&uu)[CH], uint4 (&rr)[CH]) {
        const size_t base = static_cast<size_t>(s) * (GN_THREADS * CH) + tid;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            uu[k] = __ldg(xin + base + k * GN_THREADS);
            if constexpr (HAS_RES) rr[k] = __ldg(rin + base + k * GN_THREADS);
        }
    };
    load_slab(s0, u, r);                                     // in flight before the first statistics merge

    float mul[8], add[8], post[8];
    int cur_img = -1;
    for (int s = s0; s < s1; ++s) {
        const int b = s / slabs_per_img;
        if (b != cur_img) {
            cur_img = b;
            __syncthreads();                                 // everyone is done with the previous image's statistics
            {   // warp g merges group g: nwb warp blocks x (C / 64) eight-channel pieces; 256 elements stand behind each
                // partial of the dense layout, 8 * (valid rows of the block) behind one written by a padded-slab conv
                const int nwb = a.part_tpi > 0 ? a.part_tpi * 4 : a.P / 32;   // warp blocks of this image
                const int ppg = a.C / 64;                    // pieces per group
                const int ppr = a.C / 8;                     // pieces per partial row
                const int nent = nwb * ppg;
                const float2* pp = a.part + (static_cast<size_t>(b) * nwb) * ppr + warp * ppg;
                const float total = static_cast<float>(a.P) * static_cast<float>(cpg);
                const int PW = a.part_W + 2;
                const int HPW = a.part_W > 0 ? (a.P / a.part_W) * PW : 0;
                auto block_count = [&](int blk) -> float {
                    if (a.part_tpi == 0) return 256.0f;
                    auto valid_below = [&](int n) {          // valid (non-halo, in-image) padded positions below n
                        n = n < HPW ? n : HPW;
                        const int rows = n / PW;
                        int rem = n - rows * PW - 1;
                        rem = rem < 0 ? 0 : (rem > a.part_W ? a.part_W : rem);
                        return rows * a.part_W + rem;
                    };
                    return 8.0f * static_cast<float>(valid_below(32 * blk + 32) - valid_below(32 * blk));
                };
                float mean, m2 = 0.f;
                if (nent <= 160) {                           // one batch of <= 5 independent loads per lane
                    float2 mine[5];
                    float sm = 0.f;
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const int idx = lane + 32 * i;
                        const int blk = idx / ppg;
                        mine[i] = idx < nent ? __ldg(pp + static_cast<size_t>(blk) * ppr + (idx - blk * ppg)) : make_float2(0.f, 0.f);
                        sm += mine[i].x;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
                    mean = sm / total;
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const int idx = lane + 32 * i;
                        if (idx < nent) {
                            const float cnt = block_count(idx / ppg);
                            if (cnt > 0.f) {
                                const float dm = mine[i].x / cnt - mean;
                                m2 += mine[i].y + cnt * dm * dm;
                            }
                        }
                    }
                } else {
                    float sm = 0.f;
                    for (int idx = lane; idx < nent; idx += 32) {
                        const int blk = idx / ppg;
                        sm += __ldg(pp + static_cast<size_t>(blk) * ppr + (idx - blk * ppg)).x;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
                    mean = sm / total;
                    for (int idx = lane; idx < nent; idx += 32) {   // second sweep hits L1/L2
                        const int blk = idx / ppg;
                        const float2 e = __ldg(pp + static_cast<size_t>(blk) * ppr + (idx - blk * ppg));
                        const float cnt = block_count(blk);
                        if (cnt > 0.f) {
                            const float dm = e.x / cnt - mean;
                            m2 += e.y + cnt * dm * dm;
                        }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
                if (lane == 0) {
                    const float rstd = rsqrtf(m2 / total + a.eps);
                    s_mean[warp] = mean;
                    s_rstd[warp] = rstd;
                    if (a.stats_out != nullptr && s == b * slabs_per_img)      // the CTA that owns the image's first slab records them
                        a.stats_out[b * GN_GROUPS + warp] = make_float2(mean, rstd);
                }
            }
            __syncthreads();
            const float mean = s_mean[my_g], rstd = s_rstd[my_g];
            const float* frow = nullptr;
            const float* prow = nullptr;
            if (a.film != nullptr || a.postadd != nullptr) {
                const int row = a.film_row[b * a.film_row_stride];
                if (a.film != nullptr) frow = a.film + static_cast<size_t>(row) * a.film_ld + a.film_off;
                if (a.postadd != nullptr) prow = a.postadd + static_cast<size_t>(row) * a.film_ld + a.postadd_off;
            }
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + my_c));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + my_c + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + my_c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + my_c + 4));
            const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = my_c + j;
                float gm = gam[j] * rstd;
                float bt = bet[j] - mean * gm;
                if (frow != nullptr) {
                    const float sc = __ldg(frow + c) + 1.0f;
                    const float sh = __ldg(frow + a.C + c);
                    gm *= sc;
                    bt = bt * sc + sh;
                }
                mul[j] = gm;
                add[j] = bt;
                post[j] = prow != nullptr ? __ldg(prow + c) : 0.f;
            }
        }
        uint4 un[CH], rn[CH];
        if (s + 1 < s1) load_slab(s + 1, un, rn);            // next slab in flight while this one is processed
        const size_t obase = static_cast<size_t>(s) * (GN_THREADS * CH) + tid;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            float v[8];
            float2 t;
            t = ptx::unpack_bf16x2(u[k].x); v[0] = t.x; v[1] = t.y;
            t = ptx::unpack_bf16x2(u[k].y); v[2] = t.x; v[3] = t.y;
            t = ptx::unpack_bf16x2(u[k].z); v[4] = t.x; v[5] = t.y;
            t = ptx::unpack_bf16x2(u[k].w); v[6] = t.x; v[7] = t.y;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = EXACT ? silu_exact(fmaf(v[j], mul[j], add[j])) : silu_f(fmaf(v[j], mul[j], add[j]));
                if constexpr (HAS_POST) v[j] += post[j];
            }
            if constexpr (HAS_RES) {
                t = ptx::unpack_bf16x2(r[k].x); v[0] += t.x; v[1] += t.y;
                t = ptx::unpack_bf16x2(r[k].y); v[2] += t.x; v[3] += t.y;
                t = ptx::unpack_bf16x2(r[k].z); v[4] += t.x; v[5] += t.y;
                t = ptx::unpack_bf16x2(r[k].w); v[6] += t.x; v[7] += t.y;
            }
            uint4 o;
            o.x = ptx::pack_bf16x2(v[0], v[1]);
            o.y = ptx::pack_bf16x2(v[2], v[3]);
            o.z = ptx::pack_bf16x2(v[4], v[5]);
            o.w = ptx::pack_bf16x2(v[6], v[7]);
            yout[obase + k * GN_THREADS] = o;
        }
        if (s + 1 < s1) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                u[k] = un[k];
                if constexpr (HAS_RES) r[k] = rn[k];
            }
        }
    }
}

// Statistics-only twin of the merge above (dense partial layout): one CTA per image, warp g merges group g in the same order
// with the same arithmetic, then the group's channels get their folded affine.  Consumer: the conv epilogue's `gnres` operand.
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_finalize_kernel(const float2* __restrict__ part, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float eps, float2* __restrict__ ma, const int P, const int C) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cpg = C / GN_GROUPS;
    const int nwb = P / 32;
    const int ppg = C / 64;
    const int ppr = C / 8;
    const int nent = nwb * ppg;
    const float2* pp = part + (static_cast<size_t>(b) * nwb) * ppr + warp * ppg;
    const float total = static_cast<float>(P) * static_cast<float>(cpg);
    float sm = 0.f;
    for (int idx = lane; idx < nent; idx += 32) {
        const int blk = idx / ppg;
        sm += __ldg(pp + static_cast<size_t>(blk) * ppr + (idx - blk * ppg)).x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
    const float mean = sm / total;
    float m2 = 0.f;
    for (int idx = lane; idx < nent; idx += 32) {
        const int blk = idx / ppg;
        const float2 e = __ldg(pp + static_cast<size_t>(blk) * ppr + (idx - blk * ppg));
        const float dm = e.x / 256.0f - mean;
        m2 += e.y + 256.0f * dm * dm;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
    const float rstd = rsqrtf(m2 / total + eps);
    for (int j = lane; j < cpg; j += 32) {
        const int c = warp * cpg + j;
        const float gm = __ldg(gamma + c) * rstd;
        ma[static_cast<size_t>(b) * C + c] = make_float2(gm, __ldg(beta + c) - mean * gm);
    }
}

// Stand-alone partial producer (same layout as the conv epilogue's): one warp per 32-pixel block, lanes = pixels,
// one (sum, M2) pair per 8-channel piece.
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const bf16* __restrict__ x, float2* __restrict__ part, int nblocks, int C) {
    const int wb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wb >= nblocks) return;
    const int pieces = C / 8;
    const bf16* row = x + (static_cast<size_t>(wb) * 32 + lane) * C;
    for (int p = 0; p < pieces; ++p) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + p * 8));
        float v[8];
        float2 t;
        t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
        t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
        t = ptx::unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
        t = ptx::unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float mean = s * (1.0f / 256.0f);
        float m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) m2 = fmaf(v[j] - mean, v[j] - mean, m2);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
        if (lane == 0) part[static_cast<size_t>(wb) * pieces + p] = make_float2(s, m2);
    }
}

// ----------------------------------------------------------------------------------------------- channel LN
// Lanes own 16-byte chunks (8 channels); C/8 chunks make a pixel, so a warp covers 32 / (C/8) pixels per pass (or half
// a pixel per lane pair for C = 512) and keeps LN_UNROLL passes in flight to cover HBM latency.
constexpr int LN_UNROLL = 4;

template <int C>
__global__ void __launch_bounds__(256)
channel_layernorm_kernel(const LayerNormArgs a) {
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    constexpr int CH = C / 8;                       // 16-byte chunks per pixel: 8, 16, 32, 64
    constexpr int LPP = CH < 32 ? CH : 32;          // lanes per pixel
    constexpr int CPL = CH / LPP;                   // chunks per lane (1, or 2 for C = 512)
    constexpr int PPW = 32 / LPP;                   // pixels per warp per pass
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / LPP;                     // which pixel of the pass
    const int cl = lane % LPP;                      // chunk slot inside the pixel
    const long long m0 = static_cast<long long>(warp_global) * (PPW * LN_UNROLL) + sub;

    float gain[CPL][8];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.g + (cl + q * LPP) * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.g + (cl + q * LPP) * 8 + 4));
        gain[q][0] = g0.x; gain[q][1] = g0.y; gain[q][2] = g0.z; gain[q][3] = g0.w;
        gain[q][4] = g1.x; gain[q][5] = g1.y; gain[q][6] = g1.z; gain[q][7] = g1.w;
    }

    uint4 raw[LN_UNROLL][CPL];
    uint4 rlo[LN_UNROLL][CPL];
    uint4 rres[LN_UNROLL][CPL];
#pragma unroll
    for (int u = 0; u < LN_UNROLL; ++u) {
        const long long m = m0 + u * PPW;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            raw[u][q] = make_uint4(0, 0, 0, 0);
            rlo[u][q] = make_uint4(0, 0, 0, 0);        // bf16 zeros: x + 0 when there is no low half
            rres[u][q] = make_uint4(0, 0, 0, 0);
            if (m < a.M) {
                raw[u][q] = __ldg(reinterpret_cast<const uint4*>(a.x + m * C) + cl + q * LPP);
                if (a.x_lo != nullptr) rlo[u][q] = __ldg(reinterpret_cast<const uint4*>(a.x_lo + m * C) + cl + q * LPP);
                if (a.res != nullptr) rres[u][q] = __ldg(reinterpret_cast<const uint4*>(a.res + m * C) + cl + q * LPP);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < LN_UNROLL; ++u) {
        const long long m = m0 + u * PPW;
        float v[CPL][8];
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            float2 t;
            t = ptx::unpack_bf16x2(raw[u][q].x); v[q][0] = t.x; v[q][1] = t.y;
            t = ptx::unpack_bf16x2(raw[u][q].y); v[q][2] = t.x; v[q][3] = t.y;
            t = ptx::unpack_bf16x2(raw[u][q].z); v[q][4] = t.x; v[q][5] = t.y;
            t = ptx::unpack_bf16x2(raw[u][q].w); v[q][6] = t.x; v[q][7] = t.y;
            t = ptx::unpack_bf16x2(rlo[u][q].x); v[q][0] += t.x; v[q][1] += t.y;
            t = ptx::unpack_bf16x2(rlo[u][q].y); v[q][2] += t.x; v[q][3] += t.y;
            t = ptx::unpack_bf16x2(rlo[u][q].z); v[q][4] += t.x; v[q][5] += t.y;
            t = ptx::unpack_bf16x2(rlo[u][q].w); v[q][6] += t.x; v[q][7] += t.y;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[q][j];
        }
#pragma unroll
        for (int off = LPP >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float mean = s * (1.0f / C);
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < CPL; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += (v[q][j] - mean) * (v[q][j] - mean);
#pragma unroll
        for (int off = LPP >> 1; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        const float rstd = rsqrtf(ss * (1.0f / C) + a.eps);
        if (m >= a.M) continue;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[q][j] = (v[q][j] - mean) * rstd * gain[q][j];
            if (a.res != nullptr) {
                float2 t;
                t = ptx::unpack_bf16x2(rres[u][q].x); v[q][0] += t.x; v[q][1] += t.y;
                t = ptx::unpack_bf16x2(rres[u][q].y); v[q][2] += t.x; v[q][3] += t.y;
                t = ptx::unpack_bf16x2(rres[u][q].z); v[q][4] += t.x; v[q][5] += t.y;
                t = ptx::unpack_bf16x2(rres[u][q].w); v[q][6] += t.x; v[q][7] += t.y;
            }
            uint4 o;
            o.x = ptx::pack_bf16x2(v[q][0], v[q][1]);
            o.y = ptx::pack_bf16x2(v[q][2], v[q][3]);
            o.z = ptx::pack_bf16x2(v[q][4], v[q][5]);
            o.w = ptx::pack_bf16x2(v[q][6], v[q][7]);
            const int chunk = cl + q * LPP;
            if (!a.upsample2x) {
                reinterpret_cast<uint4*>(a.y + m * C)[chunk] = o;
            } else {
                const int P = a.H * a.W;
                const int b = static_cast<int>(m / P);
                const int hw = static_cast<int>(m - static_cast<long long>(b) * P);
                const int h = hw / a.W;
                const int w = hw - h * a.W;
                const long long W2 = 2 * a.W;
                const long long base = (static_cast<long long>(b) * 2 * a.H + 2 * h) * W2 + 2 * w;
                reinterpret_cast<uint4*>(a.y + base * C)[chunk] = o;
                reinterpret_cast<uint4*>(a.y + (base + 1) * C)[chunk] = o;
                reinterpret_cast<uint4*>(a.y + (base + W2) * C)[chunk] = o;
                reinterpret_cast<uint4*>(a.y + (base + W2 + 1) * C)[chunk] = o;
            }
        }
    }
}

}  // namespace

cudaError_t groupnorm_film_silu_run(const GroupNormArgs& a, cudaStream_t s) {
    if (a.C % 64 != 0 || a.C > 512 || GN_THREADS % (a.C / 8) != 0 || a.P % 32 != 0 || a.part == nullptr)
        return cudaErrorInvalidValue;
    const size_t img_chunks = static_cast<size_t>(a.P) * a.C / 8;
    if (img_chunks % GN_SLAB_CHUNKS != 0) return cudaErrorInvalidValue;   // images are 32 KiB .. 2 MiB here
    const bool res = a.res != nullptr, post = a.postadd != nullptr;
    // 8 chunks per thread for the 2-stream launches (HD_GN_CH8=0: 4 everywhere, A/B switch)
    static const int ch8_env = [] { const char* v = getenv("HD_GN_CH8"); return v ? atoi(v) : 1; }();
    const bool wide = ch8_env != 0 && !res && !a.exact_act && img_chunks % (2 * GN_SLAB_CHUNKS) == 0;
    const int slab_chunks = wide ? 2 * GN_SLAB_CHUNKS : GN_SLAB_CHUNKS;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            num_sms = 148;
    }
    const int slabs_per_img = static_cast<int>(img_chunks / slab_chunks);
    const int total = a.B * slabs_per_img;
    int grid = 2 * num_sms;                              // two resident CTAs per SM, one contiguous run of slabs each
    if (grid > total) grid = total;
    const int per_cta = (total + grid - 1) / grid;
    grid = (total + per_cta - 1) / per_cta;
    if (wide) {
        if (post) return launch_pdl(groupnorm_apply_kernel<false, true, false, 8>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
        return launch_pdl(groupnorm_apply_kernel<false, false, false, 8>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
    }
    if (a.exact_act) {
        if (res && post) return launch_pdl(groupnorm_apply_kernel<true, true, true>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
        else if (res) return launch_pdl(groupnorm_apply_kernel<true, false, true>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
        else if (post) return launch_pdl(groupnorm_apply_kernel<false, true, true>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
        else return launch_pdl(groupnorm_apply_kernel<false, false, true>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
        return cudaGetLastError();
    }
    if (res && post) return launch_pdl(groupnorm_apply_kernel<true, true, false>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
    else if (res) return launch_pdl(groupnorm_apply_kernel<true, false, false>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
    else if (post) return launch_pdl(groupnorm_apply_kernel<false, true, false>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
    else return launch_pdl(groupnorm_apply_kernel<false, false, false>, dim3(grid), dim3(GN_THREADS), 0, s, a, slabs_per_img, total, per_cta);
    return cudaGetLastError();
}

cudaError_t groupnorm_finalize_run(const float2* part, const float* gamma, const float* beta, float eps, float2* ma, int B, int P, int C,
                                   cudaStream_t s) {
    if (B <= 0 || C % 64 != 0 || C > 512 || P % 32 != 0) return cudaErrorInvalidValue;
    return launch_pdl(groupnorm_finalize_kernel, dim3(B), dim3(GN_THREADS), 0, s, part, gamma, beta, eps, ma, P, C);
}

cudaError_t groupnorm_stats_run(const bf16* x, float2* part, int B, int P, int C, cudaStream_t s) {
    if (P % 32 != 0 || C % 64 != 0) return cudaErrorInvalidValue;
    const int nblocks = B * (P / 32);
    groupnorm_stats_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(x, part, nblocks, C);
    return cudaGetLastError();
}

cudaError_t channel_layernorm_run(const LayerNormArgs& a, cudaStream_t s) {
    const int threads = 256;
    const int lpp = (a.C / 8) < 32 ? (a.C / 8) : 32;
    const int rows_per_block = (threads / 32) * (32 / lpp) * LN_UNROLL;
    const int grid = (a.M + rows_per_block - 1) / rows_per_block;
    switch (a.C) {
        case 64: return launch_pdl(channel_layernorm_kernel<64>, dim3(grid), dim3(threads), 0, s, a);
        case 128: return launch_pdl(channel_layernorm_kernel<128>, dim3(grid), dim3(threads), 0, s, a);
        case 256: return launch_pdl(channel_layernorm_kernel<256>, dim3(grid), dim3(threads), 0, s, a);
        case 512: return launch_pdl(channel_layernorm_kernel<512>, dim3(grid), dim3(threads), 0, s, a);
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace hd
