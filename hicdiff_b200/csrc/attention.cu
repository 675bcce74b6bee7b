// Attention cores of the HiCDiff UNet (heads = 4, dim_head = 32), operating on the NHWC qkv tensor
// produced by the to_qkv 1x1 conv GEMM:  qkv[b, n, 384] = [ q(4x32) | k(4x32) | v(4x32) ].
//
// linear_attention   reference: LinearAttention.forward /root/reference/src/hicdiff_condition.py:212-227
//     q = softmax_d(q) * 32^-0.5 ; k = softmax_n(k) ; v = v / n
//     ctx[d, e] = sum_n k[d, n] v[e, n] ; out[e, n] = sum_d ctx[d, e] q[d, n]
//   Kernel 1 (one CTA per (b, head)): exact column max of k over n, then ctx and the softmax denominators in one
//   sweep with a 4x8 register tile per lane; the 1/n, 1/sum and q-scale factors are folded into ctx.
//   Kernel 2 (one thread per (pixel, head)): q softmax in registers, 32x32 ctx from padded shared memory.
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

constexpr int CTX_THREADS = 256;
constexpr int CTX_TILE = 64;   // pixels staged per sweep step

__global__ void __launch_bounds__(CTX_THREADS)
linattn_context_kernel(const LinAttnArgs a) {
    __shared__ float s_p[CTX_TILE][DH];
    __shared__ float s_v[CTX_TILE][DH];
    __shared__ float s_red[CTX_THREADS / 32][DH * DH / 4];   // reused: max reduce, then ctx cross-warp reduce
    __shared__ float s_max[DH];
    __shared__ float s_sum[CTX_THREADS / 32][DH];

    const int b = blockIdx.x / HEADS;
    const int h = blockIdx.x - b * HEADS;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const bf16* base = a.qkv + static_cast<size_t>(b) * a.n * QKV_LD;
    const bf16* kbase = base + HEADS * DH + h * DH;
    const bf16* vbase = base + 2 * HEADS * DH + h * DH;

    // ---- pass 1: max over n of k[:, d]
    const int ck = tid & 3;        // 8-wide d chunk
    const int r0 = tid >> 2;       // row within a 64-row step
    float mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
    for (int n = r0; n < a.n; n += CTX_THREADS / 4) {
        float kv[8];
        load8(kbase + static_cast<size_t>(n) * QKV_LD + ck * 8, kv);
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], kv[j]);
    }
    // lanes with equal (lane & 3) share d columns
#pragma unroll
    for (int off = 16; off >= 4; off >>= 1)
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], off));
    float* s_wmax = &s_red[0][0];   // [8 warps][32]
    if (lane < 4)
#pragma unroll
        for (int j = 0; j < 8; ++j) s_wmax[warp * DH + lane * 8 + j] = mx[j];
    __syncthreads();
    if (tid < DH) {
        float m = s_wmax[tid];
        for (int w = 1; w < CTX_THREADS / 32; ++w) m = fmaxf(m, s_wmax[w * DH + tid]);
        s_max[tid] = m;
    }
    __syncthreads();
    float kmax[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) kmax[j] = s_max[ck * 8 + j];

    // ---- pass 2: ctx[d, e] += exp(k[n,d]-max[d]) * v[n,e]; lane tile = 4 d x 8 e
    const int d0 = (lane >> 2) * 4;
    const int e0 = (lane & 3) * 8;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float psum[4] = {0.f, 0.f, 0.f, 0.f};

    for (int n0 = 0; n0 < a.n; n0 += CTX_TILE) {
        {
            const int n = n0 + r0;
            float kv[8], vv[8];
            load8(kbase + static_cast<size_t>(n) * QKV_LD + ck * 8, kv);
            load8(vbase + static_cast<size_t>(n) * QKV_LD + ck * 8, vv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s_p[r0][ck * 8 + j] = __expf(kv[j] - kmax[j]);
                s_v[r0][ck * 8 + j] = vv[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < CTX_TILE / (CTX_THREADS / 32); ++rr) {
            const int row = warp * (CTX_TILE / (CTX_THREADS / 32)) + rr;
            const float4 p4 = *reinterpret_cast<const float4*>(&s_p[row][d0]);
            const float4 va = *reinterpret_cast<const float4*>(&s_v[row][e0]);
            const float4 vb = *reinterpret_cast<const float4*>(&s_v[row][e0 + 4]);
            const float p[4] = {p4.x, p4.y, p4.z, p4.w};
            const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                psum[i] += p[i];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(p[i], v[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    // ---- cross-warp reduction, quarter of the 32x32 matrix at a time (8 d-rows per round)
    if ((lane & 3) == 0)
#pragma unroll
        for (int i = 0; i < 4; ++i) s_sum[warp][d0 + i] = psum[i];
    float* ctx_out = a.ctx + (static_cast<size_t>(b) * HEADS + h) * DH * DH;
    const float qscale = rsqrtf(static_cast<float>(DH));
    const float inv_n = 1.0f / static_cast<float>(a.n);
    for (int quarter = 0; quarter < 4; ++quarter) {
        __syncthreads();
        // lanes whose d0 falls into this quarter (d in [8q, 8q+8)) publish their 4x8 tile
        if ((d0 >> 3) == quarter) {
            const int dl = d0 & 7;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) s_red[warp][(dl + i) * DH + e0 + j] = acc[i][j];
        }
        __syncthreads();
        {
            const int idx = tid;                 // 256 entries = 8 d-rows x 32 e
            const int d = quarter * 8 + (idx >> 5);
            float tot = 0.f, ks = 0.f;
#pragma unroll
            for (int w = 0; w < CTX_THREADS / 32; ++w) {
                tot += s_red[w][idx];
                ks += s_sum[w][d];
            }
            ctx_out[d * DH + (idx & 31)] = tot / ks * inv_n * qscale;
        }
    }
}

constexpr int OUT_THREADS = 256;
constexpr int OUT_PIX = OUT_THREADS / HEADS;   // 64 pixels per CTA
constexpr int CTX_PAD = DH * DH + 4;           // stagger heads across banks

__global__ void __launch_bounds__(OUT_THREADS)
linattn_output_kernel(const LinAttnArgs a) {
    __shared__ __align__(16) float s_ctx[HEADS * CTX_PAD];
    const int blocks_per_img = a.n / OUT_PIX;
    const int b = blockIdx.x / blocks_per_img;
    const int pix0 = (blockIdx.x - b * blocks_per_img) * OUT_PIX;
    const int tid = threadIdx.x;
    const float* ctx = a.ctx + static_cast<size_t>(b) * HEADS * DH * DH;
    for (int i = tid; i < HEADS * DH * DH; i += OUT_THREADS) s_ctx[(i >> 10) * CTX_PAD + (i & 1023)] = ctx[i];
    __syncthreads();

    const int h = tid & 3;
    const int pix = pix0 + (tid >> 2);
    const bf16* qp = a.qkv + (static_cast<size_t>(b) * a.n + pix) * QKV_LD + h * DH;
    float q[DH];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float t[8];
        load8(qp + c * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) q[c * 8 + j] = t[j];
    }
    float m = q[0];
#pragma unroll
    for (int d = 1; d < DH; ++d) m = fmaxf(m, q[d]);
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) { q[d] = __expf(q[d] - m); s += q[d]; }
    const float inv = 1.0f / s;
    float o[DH];
#pragma unroll
    for (int e = 0; e < DH; ++e) o[e] = 0.f;
    const float* cx = s_ctx + h * CTX_PAD;
#pragma unroll 4
    for (int d = 0; d < DH; ++d) {
        const float qd = q[d] * inv;
#pragma unroll
        for (int e = 0; e < DH; e += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cx + d * DH + e);
            o[e] = fmaf(c4.x, qd, o[e]);
            o[e + 1] = fmaf(c4.y, qd, o[e + 1]);
            o[e + 2] = fmaf(c4.z, qd, o[e + 2]);
            o[e + 3] = fmaf(c4.w, qd, o[e + 3]);
        }
    }
    uint4* op = reinterpret_cast<uint4*>(a.out + (static_cast<size_t>(b) * a.n + pix) * OUT_LD + h * DH);
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

constexpr int FA_MAXN = 64;

__global__ void __launch_bounds__(FA_MAXN)
full_attention_kernel(const FullAttnArgs a) {
    __shared__ float s_k[FA_MAXN][DH];
    __shared__ float s_v[FA_MAXN][DH];
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
    if (a.n % CTX_TILE != 0 || a.n % OUT_PIX != 0) return cudaErrorInvalidValue;
    linattn_context_kernel<<<a.B * HEADS, CTX_THREADS, 0, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    linattn_output_kernel<<<a.B * (a.n / OUT_PIX), OUT_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t full_attention_run(const FullAttnArgs& a, cudaStream_t s) {
    if (a.n > FA_MAXN) return cudaErrorInvalidValue;
    full_attention_kernel<<<a.B * HEADS, FA_MAXN, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace hd
