// Attention cores of the HiCDiff UNet (heads = 4, dim_head = 32), operating on the NHWC qkv tensor
// produced by the to_qkv 1x1 conv GEMM:  qkv[b, n, 384] = [ q(4x32) | k(4x32) | v(4x32) ].
//
// linear_attention   reference: LinearAttention.forward /root/reference/src/hicdiff_condition.py:212-227
//     q = softmax_d(q) * 32^-0.5 ; k = softmax_n(k) ; v = v / n
//     ctx[d, e] = sum_n k[d, n] v[e, n] ; out[e, n] = sum_d ctx[d, e] q[d, n]
//   Both contractions run on the warp-level tensor-core path (mma.sync m16n8k16, bf16 in / fp32 accumulate): they are
//   32x32 per head, far too small for a tcgen05 tile, and the kernels are bound by streaming qkv once or twice.
//
// full_attention     reference: Attention.forward :239-251  (only at 8x8, n = 64)
//     sim = (q * 32^-0.5)^T k ; attn = softmax_j(sim) ; out[i, d] = sum_j attn[i, j] v[d, j]
#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

constexpr int HEADS = 4;
constexpr int DH = 32;
constexpr int QKV_LD = 3 * HEADS * DH;   // 384
constexpr int OUT_LD = HEADS * DH;       // 128

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    float2 t;
    t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
    t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
    t = ptx::unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
    t = ptx::unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 1: per (b, head)  ctxT[e][d] = (sum_n p[n,d] v[n,e]) / (sum_n p[n,d]) / n * 32^-0.5,  p = exp(k - max_n k)
//   pass 1: exact column max of k over n (fp32, shuffles + smem);
//   pass 2: 128-pixel tiles: p (bf16) and v are staged in shared memory (80-byte row pitch: conflict-free ldmatrix),
//           each warp owns 16 pixels per tile and issues 8 mma.sync.m16n8k16 (A = p^T via ldmatrix.trans, B = v via
//           ldmatrix.trans); the softmax denominator sums the SAME bf16-rounded p the MMAs consume.
//   Deterministic: fixed-order cross-warp reduction, no atomics.
// ------------------------------------------------------------------------------------------------------------
constexpr int CTX_THREADS = 256;
constexpr int CTX_WARPS = CTX_THREADS / 32;
constexpr int CTX_TILE = 128;           // pixels staged per iteration (16 per warp)
constexpr int CTX_PITCH = 40;           // bf16 elements per staged row (32 + 8 pad) = 80 bytes

__global__ void __launch_bounds__(CTX_THREADS)
linattn_context_kernel(const LinAttnArgs a) {
    __shared__ __align__(16) unsigned char s_buf[CTX_WARPS * DH * DH * 4];   // staging (20 KB) / reduction (32 KB)
    __shared__ float s_max[DH];
    __shared__ float s_wsum[CTX_WARPS][DH];
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    bf16* sP = reinterpret_cast<bf16*>(s_buf);
    bf16* sV = sP + CTX_TILE * CTX_PITCH;
    float* s_red = reinterpret_cast<float*>(s_buf);

    const int b = blockIdx.x / HEADS;
    const int h = blockIdx.x - b * HEADS;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const bf16* base = a.qkv + static_cast<size_t>(b) * a.n * QKV_LD;
    const bf16* kbase = base + HEADS * DH + h * DH;
    const bf16* vbase = base + 2 * HEADS * DH + h * DH;

    // ---- pass 1: max over n of k[:, d]; thread = (row tid/4 of a 64-row step, 8-wide d chunk tid%4)
    {
        const int ck = tid & 3;
        float mx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
        for (int n = tid >> 2; n < a.n; n += CTX_THREADS / 4) {
            float kv[8];
            load8(kbase + static_cast<size_t>(n) * QKV_LD + ck * 8, kv);
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], kv[j]);
        }
#pragma unroll
        for (int off = 16; off >= 4; off >>= 1)
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], off));
        if (lane < 4)
#pragma unroll
            for (int j = 0; j < 8; ++j) s_wsum[warp][lane * 8 + j] = mx[j];
        __syncthreads();
        if (tid < DH) {
            float m = s_wsum[0][tid];
            for (int w = 1; w < CTX_WARPS; ++w) m = fmaxf(m, s_wsum[w][tid]);
            s_max[tid] = m;
        }
        __syncthreads();
    }

    // ---- pass 2
    const int srow = tid >> 1;            // staged row (pixel within the tile) this thread fills
    const int shalf = tid & 1;            // which 16 of the 32 channels
    float kmax[16], psum[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { kmax[j] = s_max[shalf * 16 + j]; psum[j] = 0.f; }
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;

    // ldmatrix row addresses of this lane (constant across tiles)
    const int lm = lane >> 3, lj = lane & 7;
    const int prow = warp * 16 + lj + (lm >> 1) * 8;     // A: matrices (0,1) pixels 0-7, (2,3) pixels 8-15
    const int pcol = (lm & 1) * 8;                       //    matrices (0,2) d 0-7,    (1,3) d 8-15
    const int vrow = warp * 16 + lj + (lm & 1) * 8;      // B: matrices (0,2) pixels 0-7, (1,3) pixels 8-15
    const int vcol = (lm >> 1) * 8;                      //    matrices (0,1) e 0-7,    (2,3) e 8-15

    for (int n0 = 0; n0 < a.n; n0 += CTX_TILE) {
        {
            uint32_t pk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
            if (n0 + srow < a.n) {   // rows past the image (n = 64 < one tile) contribute p = 0, v = 0
                const size_t g = static_cast<size_t>(n0 + srow) * QKV_LD + shalf * 16;
                float k0[8], k1[8];
                load8(kbase + g, k0);
                load8(kbase + g + 8, k1);
                v0 = __ldg(reinterpret_cast<const uint4*>(vbase + g));
                v1 = __ldg(reinterpret_cast<const uint4*>(vbase + g + 8));
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const __nv_bfloat162 pa = __floats2bfloat162_rn(__expf(k0[j] - kmax[j]), __expf(k0[j + 1] - kmax[j + 1]));
                    const __nv_bfloat162 pb = __floats2bfloat162_rn(__expf(k1[j] - kmax[8 + j]), __expf(k1[j + 1] - kmax[8 + j + 1]));
                    const float2 fa = __bfloat1622float2(pa), fb = __bfloat1622float2(pb);
                    psum[j] += fa.x; psum[j + 1] += fa.y;
                    psum[8 + j] += fb.x; psum[8 + j + 1] += fb.y;
                    pk[j / 2] = *reinterpret_cast<const uint32_t*>(&pa);
                    pk[4 + j / 2] = *reinterpret_cast<const uint32_t*>(&pb);
                }
            }
            uint4* dp = reinterpret_cast<uint4*>(sP + srow * CTX_PITCH + shalf * 16);
            dp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            uint4* dv = reinterpret_cast<uint4*>(sV + srow * CTX_PITCH + shalf * 16);
            dv[0] = v0;
            dv[1] = v1;
        }
        __syncthreads();
        uint32_t af[2][4], bfr[2][4];
        ptx::ldmatrix_x4_trans(af[0], sP + prow * CTX_PITCH + pcol);            // d 0-15
        ptx::ldmatrix_x4_trans(af[1], sP + prow * CTX_PITCH + 16 + pcol);       // d 16-31
        ptx::ldmatrix_x4_trans(bfr[0], sV + vrow * CTX_PITCH + vcol);           // e 0-15  (two n-tiles)
        ptx::ldmatrix_x4_trans(bfr[1], sV + vrow * CTX_PITCH + 16 + vcol);      // e 16-31
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                ptx::mma_bf16_16816(acc[mt][nt], af[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
        __syncthreads();
    }

    // ---- softmax denominators: lanes with equal parity share the channel half
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1)
#pragma unroll
        for (int j = 0; j < 16; ++j) psum[j] += __shfl_xor_sync(0xffffffffu, psum[j], off);
    if (lane < 2)
#pragma unroll
        for (int j = 0; j < 16; ++j) s_wsum[warp][lane * 16 + j] = psum[j];

    // ---- cross-warp reduction of the 32x32 context (fixed order), then normalise and emit ctx^T in bf16
    {
        const int g = lane >> 2, t = lane & 3;
        float* mine = s_red + warp * DH * DH;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int d = mt * 16 + g, e = nt * 8 + 2 * t;
                mine[d * DH + e] = acc[mt][nt][0];
                mine[d * DH + e + 1] = acc[mt][nt][1];
                mine[(d + 8) * DH + e] = acc[mt][nt][2];
                mine[(d + 8) * DH + e + 1] = acc[mt][nt][3];
            }
    }
    __syncthreads();
    bf16* ctx_out = reinterpret_cast<bf16*>(a.ctx) + (static_cast<size_t>(b) * HEADS + h) * DH * DH;
    const float qscale = rsqrtf(static_cast<float>(DH));
    const float inv_n = 1.0f / static_cast<float>(a.n);
    for (int idx = tid; idx < DH * DH; idx += CTX_THREADS) {
        const int d = idx >> 5, e = idx & 31;
        float tot = 0.f, ks = 0.f;
#pragma unroll
        for (int w = 0; w < CTX_WARPS; ++w) {
            tot += s_red[w * DH * DH + idx];
            ks += s_wsum[w][d];
        }
        ctx_out[e * DH + d] = __float2bfloat16(tot / ks * inv_n * qscale);     // transposed: [e][d]
    }
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 2: out[n, h*32 + e] = sum_d softmax_d(q[n, h, :])[d] * ctxT[h][e][d]
//   CTA = 4 warps x 16 pixels; q rows (all 4 heads, 256 B) are staged with coalesced 16-byte loads, the softmax runs on
//   the MMA A-fragment layout (row spread over a lane quad -> 2 shuffles), 8 mma.sync per head, and the result goes
//   back through the same staging buffer so global stores are full 16-byte, row-contiguous.
// ------------------------------------------------------------------------------------------------------------
constexpr int OUT_THREADS = 128;
constexpr int OUT_PIX = 64;              // pixels per CTA (16 per warp)
constexpr int OUT_PITCH = 136;           // bf16 per staged row: 128 + 8 pad = 272 bytes (conflict-free fragment access)
constexpr int CTXT_PITCH = 40;           // bf16 per staged ctx^T row: 32 + 8 pad = 80 bytes

__global__ void __launch_bounds__(OUT_THREADS)
linattn_output_kernel(const LinAttnArgs a) {
    __shared__ __align__(16) bf16 s_q[OUT_PIX * OUT_PITCH];
    __shared__ __align__(16) bf16 s_ctx[HEADS * DH * CTXT_PITCH];
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    const int blocks_per_img = a.n / OUT_PIX;
    const int b = blockIdx.x / blocks_per_img;
    const int pix0 = (blockIdx.x - b * blocks_per_img) * OUT_PIX;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.ctx) + static_cast<size_t>(b) * HEADS * DH * DH);
        for (int i = tid; i < HEADS * DH * DH / 8; i += OUT_THREADS)     // 4 chunks of 16 B per 32-wide row
            *reinterpret_cast<uint4*>(s_ctx + (i >> 2) * CTXT_PITCH + (i & 3) * 8) = __ldg(src + i);
    }
    // q: 64 rows x 256 B, 16 chunks of 16 B per row
    const bf16* qbase = a.qkv + (static_cast<size_t>(b) * a.n + pix0) * QKV_LD;
    for (int i = tid; i < OUT_PIX * 16; i += OUT_THREADS) {
        const int r = i >> 4, c = i & 15;
        *reinterpret_cast<uint4*>(s_q + r * OUT_PITCH + c * 8) = __ldg(reinterpret_cast<const uint4*>(qbase + static_cast<size_t>(r) * QKV_LD) + c);
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    bf16* rows = s_q + (warp * 16) * OUT_PITCH;
#pragma unroll 1
    for (int h = 0; h < HEADS; ++h) {
        // A fragments of raw q: rows g / g+8, columns 2t(+1) + {0, 8, 16, 24}
        float q[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float2 v = ptx::unpack_bf16x2(*reinterpret_cast<const uint32_t*>(rows + (g + 8 * r) * OUT_PITCH + h * DH + c * 8 + 2 * t));
                q[r][2 * c] = v.x;
                q[r][2 * c + 1] = v.y;
            }
        uint32_t af[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float m = q[r][0];
#pragma unroll
            for (int j = 1; j < 8; ++j) m = fmaxf(m, q[r][j]);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { q[r][j] = __expf(q[r][j] - m); s += q[r][j]; }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const float inv = 1.0f / s;
            // k-step 0 covers d 0-15 (columns c = 0, 1), k-step 1 covers d 16-31 (c = 2, 3)
            af[0][r] = ptx::pack_bf16x2(q[r][0] * inv, q[r][1] * inv);
            af[0][r + 2] = ptx::pack_bf16x2(q[r][2] * inv, q[r][3] * inv);
            af[1][r] = ptx::pack_bf16x2(q[r][4] * inv, q[r][5] * inv);
            af[1][r + 2] = ptx::pack_bf16x2(q[r][6] * inv, q[r][7] * inv);
        }
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int r = 0; r < 4; ++r) o[nt][r] = 0.f;
            const bf16* cx = s_ctx + (h * DH + nt * 8 + g) * CTXT_PITCH;   // ctxT[e = nt*8 + g][d]
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(cx + ks * 16 + 2 * t);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(cx + ks * 16 + 8 + 2 * t);
                ptx::mma_bf16_16816(o[nt], af[ks], b0, b1);
            }
        }
        __syncwarp();   // every lane has consumed this head's q before it is overwritten with the result
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<uint32_t*>(rows + g * OUT_PITCH + h * DH + nt * 8 + 2 * t) = ptx::pack_bf16x2(o[nt][0], o[nt][1]);
            *reinterpret_cast<uint32_t*>(rows + (g + 8) * OUT_PITCH + h * DH + nt * 8 + 2 * t) = ptx::pack_bf16x2(o[nt][2], o[nt][3]);
        }
    }
    __syncwarp();
    // coalesced write-back of this warp's 16 rows x 256 B
    bf16* obase = a.out + (static_cast<size_t>(b) * a.n + pix0 + warp * 16) * OUT_LD;
#pragma unroll
    for (int i = lane; i < 16 * 16; i += 32) {
        const int r = i >> 4, c = i & 15;
        reinterpret_cast<uint4*>(obase + static_cast<size_t>(r) * OUT_LD)[c] = *reinterpret_cast<const uint4*>(rows + r * OUT_PITCH + c * 8);
    }
}

constexpr int FA_MAXN = 64;

__global__ void __launch_bounds__(FA_MAXN)
full_attention_kernel(const FullAttnArgs a) {
    __shared__ float s_k[FA_MAXN][DH];
    __shared__ float s_v[FA_MAXN][DH];
    ptx::grid_dep_launch();     // PDL (ptx.cuh): the successor may be scheduled; then wait for the predecessor grid and its memory
    ptx::grid_dep_wait();
    const int b = blockIdx.x / HEADS;
    const int h = blockIdx.x - b * HEADS;
    const int i = threadIdx.x;
    const bf16* base = a.qkv + static_cast<size_t>(b) * a.n * QKV_LD;
    float q[DH];
    if (i < a.n) {
        const bf16* row = base + static_cast<size_t>(i) * QKV_LD + h * DH;
        const float qscale = rsqrtf(static_cast<float>(DH));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float t[8];
            load8(row + c * 8, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) q[c * 8 + j] = t[j] * qscale;
            load8(row + HEADS * DH + c * 8, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) s_k[i][c * 8 + j] = t[j];
            load8(row + 2 * HEADS * DH + c * 8, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) s_v[i][c * 8 + j] = t[j];
        }
    }
    __syncthreads();
    if (i >= a.n) return;
    float sim[FA_MAXN];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < FA_MAXN; ++j) {
        float acc = 0.f;
        if (j < a.n) {
#pragma unroll
            for (int d = 0; d < DH; ++d) acc = fmaf(q[d], s_k[j][d], acc);
            m = fmaxf(m, acc);
        }
        sim[j] = acc;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < FA_MAXN; ++j) {
        sim[j] = j < a.n ? __expf(sim[j] - m) : 0.f;
        s += sim[j];
    }
    const float inv = 1.0f / s;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < FA_MAXN; ++j) {
        const float p = sim[j] * inv;
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = fmaf(p, s_v[j][d], o[d]);
    }
    uint4* op = reinterpret_cast<uint4*>(a.out + (static_cast<size_t>(b) * a.n + i) * OUT_LD + h * DH);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = ptx::pack_bf16x2(o[c * 8 + 0], o[c * 8 + 1]);
        u.y = ptx::pack_bf16x2(o[c * 8 + 2], o[c * 8 + 3]);
        u.z = ptx::pack_bf16x2(o[c * 8 + 4], o[c * 8 + 5]);
        u.w = ptx::pack_bf16x2(o[c * 8 + 6], o[c * 8 + 7]);
        op[c] = u;
    }
}

}  // namespace

cudaError_t linear_attention_run(const LinAttnArgs& a, cudaStream_t s) {
    if (a.n % OUT_PIX != 0) return cudaErrorInvalidValue;
    cudaError_t e = launch_pdl(linattn_context_kernel, dim3(a.B * HEADS), dim3(CTX_THREADS), 0, s, a);
    if (e != cudaSuccess) return e;
    return launch_pdl(linattn_output_kernel, dim3(a.B * (a.n / OUT_PIX)), dim3(OUT_THREADS), 0, s, a);
}

cudaError_t full_attention_run(const FullAttnArgs& a, cudaStream_t s) {
    if (a.n > FA_MAXN) return cudaErrorInvalidValue;
    return launch_pdl(full_attention_kernel, dim3(a.B * HEADS), dim3(FA_MAXN), 0, s, a);
}

}  // namespace hd
