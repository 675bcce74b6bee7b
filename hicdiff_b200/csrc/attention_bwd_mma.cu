// Tensor-core (mma.sync m16n8k16, bf16 x bf16 -> fp32) form of the linear-attention backward of the Unet training step
// (LinearAttention.forward, /root/reference/src/hicdiff_condition.py:212-227; formulas in attention_bwd.cu).  Two launches
// replace the six of the CUDA-core form:
//   la_ctx_mma_kernel   one CTA per (image, head): column max / sum of k, then ctx[d,e] = sum_n ks vs and dctx[d,e] = sum_n qs dout
//                       as 128-pixel tiles on the ldmatrix / mma.sync pipeline of the forward's linattn_context_kernel
//                       (was la_kstats + la_ctx (8 pixel chunks) + la_reduce: fp32 FMAs, 0.42 ms at the 64x64 level)
//   la_grad_mma_kernel  a warp owns 16 pixels; dqs = dout ctx^T, dks = vs dctx^T, dv = ks dctx are [16 x 32] x [32 x 32] tiles
//                       whose A and C fragments hold the same elements per thread, so the softmax epilogues stay thread-local
//                       (was la_grad_kernel: 32 x 32 mat-vecs per pixel on the FMA pipe, 1.05 ms).  t[d] = sum_n dks ks equals
//                       sum_e dctx[d,e] ctx[d,e] (ctx IS sum_n ks vs), so the former la_t pass over the pixels is a warp
//                       reduction while the two matrices are staged.
// Both were validated stand-alone on a B200 against a CPU evaluation of the same formulas before they moved here
// (rel-RMS: ctx 1.4e-3, dctx 1.7e-3, dq 2.4e-3, dk 1.6e-3, dv 2.9e-3 -- the bf16 operand rounding; profiles/r02_notes.md).
#include "kernels.h"

namespace hd {
namespace la_mma {

constexpr int HEADS = 4, DH = 32, QKV_LD = 3 * HEADS * DH, OUT_LD = HEADS * DH;
constexpr int CTX_THREADS = 256, CTX_WARPS = CTX_THREADS / 32;
constexpr int CTX_TILE = 128;           // pixels staged per iteration (16 per warp)
constexpr int CTX_PITCH = 40;           // bf16 elements per staged row (32 + 8 pad) = 80 bytes: conflict-free ldmatrix

namespace ctxk {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_row)) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// grid B * HEADS, 256 threads.  cd: [bh][2][32][32] fp32 (ctx, dctx) as [d][e]; kmax / ksum: [bh][32]
__global__ void __launch_bounds__(CTX_THREADS)
la_ctx_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, int n, float scale, float* __restrict__ kmax_out,
                  float* __restrict__ ksum_out, float* __restrict__ cd) {
    __shared__ __align__(16) unsigned char s_buf[4 * CTX_TILE * CTX_PITCH * 2];   // staging (40 KB) / reduction (32 KB)
    __shared__ float s_max[DH];
    __shared__ float s_wsum[CTX_WARPS][DH];
    bf16* sP = reinterpret_cast<bf16*>(s_buf);          // exp(k - kmax)
    bf16* sV = sP + CTX_TILE * CTX_PITCH;               // v
    bf16* sQ = sV + CTX_TILE * CTX_PITCH;               // softmax_d(q) * scale
    bf16* sG = sQ + CTX_TILE * CTX_PITCH;               // dout
    float* s_red = reinterpret_cast<float*>(s_buf);

    const int bh = blockIdx.x, b = bh / HEADS, h = bh - b * HEADS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bf16* base = qkv + static_cast<size_t>(b) * n * QKV_LD;
    const bf16* qbase = base + h * DH;
    const bf16* kbase = base + HEADS * DH + h * DH;
    const bf16* vbase = base + 2 * HEADS * DH + h * DH;
    const bf16* gbase = dout + static_cast<size_t>(b) * n * OUT_LD + h * DH;

    // ---- pass 1: max over n of k[:, d]; thread = (row tid/4 of a 64-row step, 8-wide d chunk tid%4)
    {
        const int ck = tid & 3;
        float mx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
        for (int p = tid >> 2; p < n; p += CTX_THREADS / 4) {
            float kv[8];
            load8(kbase + static_cast<size_t>(p) * QKV_LD + ck * 8, kv);
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
    float acc_c[2][4][4], acc_d[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) { acc_c[mt][nt][r] = 0.f; acc_d[mt][nt][r] = 0.f; }

    // ldmatrix row addresses of this lane (constant across tiles)
    const int lm = lane >> 3, lj = lane & 7;
    const int prow = warp * 16 + lj + (lm >> 1) * 8;     // A: matrices (0,1) pixels 0-7, (2,3) pixels 8-15
    const int pcol = (lm & 1) * 8;                       //    matrices (0,2) d 0-7,    (1,3) d 8-15
    const int vrow = warp * 16 + lj + (lm & 1) * 8;      // B: matrices (0,2) pixels 0-7, (1,3) pixels 8-15
    const int vcol = (lm >> 1) * 8;                      //    matrices (0,1) e 0-7,    (2,3) e 8-15

    for (int n0 = 0; n0 < n; n0 += CTX_TILE) {
        {
            const bool valid = n0 + srow < n;          // rows past the image contribute p = qs = 0, v = dout = 0
            const size_t pix = static_cast<size_t>(valid ? n0 + srow : 0);
            const size_t g = pix * QKV_LD + shalf * 16;
            float k0[8], k1[8], q0[8], q1[8];
            load8(kbase + g, k0);
            load8(kbase + g + 8, k1);
            load8(qbase + g, q0);
            load8(qbase + g + 8, q1);
            uint4 v0 = __ldg(reinterpret_cast<const uint4*>(vbase + g)), v1 = __ldg(reinterpret_cast<const uint4*>(vbase + g + 8));
            uint4 g0 = __ldg(reinterpret_cast<const uint4*>(gbase + pix * OUT_LD + shalf * 16));
            uint4 g1 = __ldg(reinterpret_cast<const uint4*>(gbase + pix * OUT_LD + shalf * 16 + 8));
            // softmax over the 32 channels of the pixel: this thread holds 16, lane ^ 1 the other 16 (same pixel, same validity)
            float m = q0[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) m = fmaxf(m, q0[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) m = fmaxf(m, q1[j]);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { q0[j] = __expf(q0[j] - m); q1[j] = __expf(q1[j] - m); s += q0[j] + q1[j]; }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            const float qn = valid ? scale / s : 0.f;
            uint32_t pk[8], pq[8];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const __nv_bfloat162 pa = __floats2bfloat162_rn(valid ? __expf(k0[j] - kmax[j]) : 0.f, valid ? __expf(k0[j + 1] - kmax[j + 1]) : 0.f);
                const __nv_bfloat162 pb = __floats2bfloat162_rn(valid ? __expf(k1[j] - kmax[8 + j]) : 0.f, valid ? __expf(k1[j + 1] - kmax[8 + j + 1]) : 0.f);
                const float2 fa = __bfloat1622float2(pa), fb = __bfloat1622float2(pb);
                psum[j] += fa.x; psum[j + 1] += fa.y;
                psum[8 + j] += fb.x; psum[8 + j + 1] += fb.y;
                pk[j / 2] = *reinterpret_cast<const uint32_t*>(&pa);
                pk[4 + j / 2] = *reinterpret_cast<const uint32_t*>(&pb);
                pq[j / 2] = pack2(q0[j] * qn, q0[j + 1] * qn);
                pq[4 + j / 2] = pack2(q1[j] * qn, q1[j + 1] * qn);
            }
            if (!valid) { v0 = v1 = g0 = g1 = make_uint4(0, 0, 0, 0); }
            const int so = srow * CTX_PITCH + shalf * 16;
            reinterpret_cast<uint4*>(sP + so)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4*>(sP + so)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            reinterpret_cast<uint4*>(sQ + so)[0] = make_uint4(pq[0], pq[1], pq[2], pq[3]);
            reinterpret_cast<uint4*>(sQ + so)[1] = make_uint4(pq[4], pq[5], pq[6], pq[7]);
            reinterpret_cast<uint4*>(sV + so)[0] = v0;
            reinterpret_cast<uint4*>(sV + so)[1] = v1;
            reinterpret_cast<uint4*>(sG + so)[0] = g0;
            reinterpret_cast<uint4*>(sG + so)[1] = g1;
        }
        __syncthreads();
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const bf16* sA = which == 0 ? sP : sQ;
            const bf16* sB = which == 0 ? sV : sG;
            uint32_t af[2][4], bfr[2][4];
            ldmatrix_x4_trans(af[0], sA + prow * CTX_PITCH + pcol);            // d 0-15
            ldmatrix_x4_trans(af[1], sA + prow * CTX_PITCH + 16 + pcol);       // d 16-31
            ldmatrix_x4_trans(bfr[0], sB + vrow * CTX_PITCH + vcol);           // e 0-15  (two n-tiles)
            ldmatrix_x4_trans(bfr[1], sB + vrow * CTX_PITCH + 16 + vcol);      // e 16-31
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    if (which == 0) mma_bf16_16816(acc_c[mt][nt], af[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
                    else mma_bf16_16816(acc_d[mt][nt], af[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
                }
        }
        __syncthreads();
    }

    // ---- softmax denominators of k: lanes with equal parity share the channel half
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1)
#pragma unroll
        for (int j = 0; j < 16; ++j) psum[j] += __shfl_xor_sync(0xffffffffu, psum[j], off);
    if (lane < 2)
#pragma unroll
        for (int j = 0; j < 16; ++j) s_wsum[warp][lane * 16 + j] = psum[j];

    // ---- cross-warp reduction (fixed order): ctx, then dctx through the same buffer
    const int g = lane >> 2, t = lane & 3;
    const float inv_n = 1.0f / static_cast<float>(n);
    float* out = cd + static_cast<size_t>(bh) * 2 * DH * DH;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        float* mine = s_red + warp * DH * DH;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int d = mt * 16 + g, e = nt * 8 + 2 * t;
                const float* a4 = which == 0 ? acc_c[mt][nt] : acc_d[mt][nt];
                mine[d * DH + e] = a4[0];
                mine[d * DH + e + 1] = a4[1];
                mine[(d + 8) * DH + e] = a4[2];
                mine[(d + 8) * DH + e + 1] = a4[3];
            }
        __syncthreads();
        for (int idx = tid; idx < DH * DH; idx += CTX_THREADS) {
            const int d = idx >> 5;
            float tot = 0.f, ks = 0.f;
#pragma unroll
            for (int w = 0; w < CTX_WARPS; ++w) { tot += s_red[w * DH * DH + idx]; ks += s_wsum[w][d]; }
            if (which == 0) {
                out[idx] = tot / ks * inv_n;
                if ((idx & 31) == 0) { ksum_out[bh * DH + d] = ks; kmax_out[bh * DH + d] = s_max[d]; }
            } else {
                out[DH * DH + idx] = tot;
            }
        }
        __syncthreads();
    }
}
}  // namespace ctxk

namespace gradk {
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// A fragments of a [16 x 32] bf16 tile whose rows are r0 (lane rows g) and r1 (rows g + 8): a[ks][i], ks = K step (16 columns).
// Element map: a[ks][0] = (row g, cols 16 ks + 2 t, + 1), a[ks][1] = (row g + 8, same cols), a[ks][2] / a[ks][3] = cols + 8.
__device__ __forceinline__ void load_a(const bf16* r0, const bf16* r1, int t, uint32_t (&a)[2][4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 16 * ks + 2 * t));
        a[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(r1 + 16 * ks + 2 * t));
        a[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 16 * ks + 8 + 2 * t));
        a[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(r1 + 16 * ks + 8 + 2 * t));
    }
}
// the same tile as floats in the C-fragment arrangement: v[j][0..1] = (row g, cols 8 j + 2 t, + 1), v[j][2..3] = row g + 8;
// column block j = 2 ks + (i >> 1), row half = i & 1
__device__ __forceinline__ void a_to_c(const uint32_t (&a)[2][4], float (&v)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 lo = unpack2(a[j >> 1][(j & 1) * 2]), hi = unpack2(a[j >> 1][(j & 1) * 2 + 1]);
        v[j][0] = lo.x; v[j][1] = lo.y; v[j][2] = hi.x; v[j][3] = hi.y;
    }
}
__device__ __forceinline__ void c_to_a(const float (&v)[4][4], uint32_t (&a)[2][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j >> 1][(j & 1) * 2] = pack2(v[j][0], v[j][1]);
        a[j >> 1][(j & 1) * 2 + 1] = pack2(v[j][2], v[j][3]);
    }
}
// out[16 x 32] = X[16 x 32] * M^T, M[n][k] given as B fragments b[ks][j] (column block j of the output)
__device__ __forceinline__ void tile_matmul(const uint32_t (&a)[2][4], const uint32_t (&b)[2][4][2], float (&acc)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) mma16816(acc[j], a[ks], b[ks][j]);
    }
}
// store a C-arranged tile as bf16 rows
__device__ __forceinline__ void store_c(bf16* r0, bf16* r1, bool ok0, bool ok1, int t, const float (&v)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ok0) *reinterpret_cast<uint32_t*>(r0 + 8 * j + 2 * t) = pack2(v[j][0], v[j][1]);
        if (ok1) *reinterpret_cast<uint32_t*>(r1 + 8 * j + 2 * t) = pack2(v[j][2], v[j][3]);
    }
}

// grid (chunks, B * HEADS), 128 threads = 4 warps, each warp strides over 16-pixel tiles of the chunk
// cd: [bh][2][32][32] fp32 (ctx, dctx) as [d][e]; kmax / ksum / tvec: [bh][32]
__global__ void __launch_bounds__(128)
la_grad_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, int n, float scale, const float* __restrict__ kmax,
                   const float* __restrict__ ksum, const float* __restrict__ cd, bf16* __restrict__ dqkv) {
    __shared__ __align__(16) bf16 s_m[3][DH][DH];         // [0] ctx[d][e], [1] dctx[d][e], [2] dctx^T [e][d]: M[n][k] of the three products
    __shared__ float s_km[DH], s_kinv[DH], s_t[DH];
    const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
    for (int i = threadIdx.x; i < DH * DH; i += blockDim.x) {
        const float c = cd[static_cast<size_t>(bh) * 2 * DH * DH + i];
        const float d = cd[(static_cast<size_t>(bh) * 2 + 1) * DH * DH + i];
        s_m[0][i / DH][i % DH] = __float2bfloat16(c);
        s_m[1][i / DH][i % DH] = __float2bfloat16(d);
        s_m[2][i % DH][i / DH] = __float2bfloat16(d);
        // t[d] = sum_n dks[n,d] ks[n,d] = sum_e dctx[d,e] ctx[d,e] (ctx is exactly sum_n ks vs): no pass over the pixels.  The block
        // is four whole warps and DH = 32, so in every iteration a warp holds exactly row d = i / DH with lane = e.
        float tsum = c * d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
        if ((threadIdx.x & 31) == 0) s_t[i / DH] = tsum;
    }
    if (threadIdx.x < DH) {
        s_km[threadIdx.x] = kmax[bh * DH + threadIdx.x];
        s_kinv[threadIdx.x] = 1.0f / ksum[bh * DH + threadIdx.x];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // B fragments of the three matrices, resident for the whole kernel: b[m][ks][j] = {M[8 j + g][16 ks + 2 t, + 1], M[..][+ 8, + 9]}
    uint32_t bm[3][2][4][2];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bm[m][ks][j][0] = *reinterpret_cast<const uint32_t*>(&s_m[m][8 * j + g][16 * ks + 2 * t]);
                bm[m][ks][j][1] = *reinterpret_cast<const uint32_t*>(&s_m[m][8 * j + g][16 * ks + 8 + 2 * t]);
            }
    // per-column constants of the k softmax / dk, for this thread's columns 8 j + 2 t, + 1
    float km[4][2], kinv[4][2], tv[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            km[j][i] = s_km[8 * j + 2 * t + i]; kinv[j][i] = s_kinv[8 * j + 2 * t + i]; tv[j][i] = s_t[8 * j + 2 * t + i];
        }

    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    const float inv_n = 1.0f / static_cast<float>(n);
    const bf16* base = qkv + static_cast<size_t>(b) * n * QKV_LD + h * DH;
    const bf16* dob = dout + static_cast<size_t>(b) * n * OUT_LD + h * DH;
    bf16* ob = dqkv + static_cast<size_t>(b) * n * QKV_LD + h * DH;
    for (int pt = p0 + warp * 16; pt < p1; pt += 4 * 16) {
        const int pa = pt + g, pb = pt + g + 8;
        const bool oka = pa < p1, okb = pb < p1;
        const size_t ra = static_cast<size_t>(oka ? pa : p1 - 1), rb = static_cast<size_t>(okb ? pb : p1 - 1);   // clamp: loads stay in range
        uint32_t a[2][4];
        float acc[4][4], x[4][4];
        // ---- dq = s * y o (dqs - <dqs, y>), y = softmax_d(q), dqs = dout * ctx^T
        load_a(dob + ra * OUT_LD, dob + rb * OUT_LD, t, a);
        tile_matmul(a, bm[0], acc);
        load_a(base + ra * QKV_LD, base + rb * QKV_LD, t, a);
        a_to_c(a, x);
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) { m0 = fmaxf(m0, fmaxf(x[j][0], x[j][1])); m1 = fmaxf(m1, fmaxf(x[j][2], x[j][3])); }
        m0 = quad_max(m0); m1 = quad_max(m1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j][0] = __expf(x[j][0] - m0); x[j][1] = __expf(x[j][1] - m0); x[j][2] = __expf(x[j][2] - m1); x[j][3] = __expf(x[j][3] - m1);
            s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3];
        }
        const float i0 = 1.0f / quad_sum(s0), i1 = 1.0f / quad_sum(s1);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j][0] *= i0; x[j][1] *= i0; x[j][2] *= i1; x[j][3] *= i1;
            d0 = fmaf(acc[j][0], x[j][0], fmaf(acc[j][1], x[j][1], d0));
            d1 = fmaf(acc[j][2], x[j][2], fmaf(acc[j][3], x[j][3], d1));
        }
        d0 = quad_sum(d0); d1 = quad_sum(d1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j][0] = scale * x[j][0] * (acc[j][0] - d0); x[j][1] = scale * x[j][1] * (acc[j][1] - d0);
            x[j][2] = scale * x[j][2] * (acc[j][2] - d1); x[j][3] = scale * x[j][3] * (acc[j][3] - d1);
        }
        store_c(ob + ra * QKV_LD, ob + rb * QKV_LD, oka, okb, t, x);
        // ---- dk = ks o (dks - t), dks = (v / n) * dctx^T, ks = exp(k - kmax) / ksum
        load_a(base + ra * QKV_LD + 2 * HEADS * DH, base + rb * QKV_LD + 2 * HEADS * DH, t, a);
        tile_matmul(a, bm[1], acc);
        load_a(base + ra * QKV_LD + HEADS * DH, base + rb * QKV_LD + HEADS * DH, t, a);
        a_to_c(a, x);
        float o[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[j][i] = __expf(x[j][i] - km[j][i & 1]) * kinv[j][i & 1];                  // ks
                o[j][i] = x[j][i] * (acc[j][i] * inv_n - tv[j][i & 1]);
            }
        store_c(ob + ra * QKV_LD + HEADS * DH, ob + rb * QKV_LD + HEADS * DH, oka, okb, t, o);
        // ---- dv = (1 / n) ks * dctx  (A = ks rounded to bf16)
        c_to_a(x, a);
        tile_matmul(a, bm[2], acc);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] *= inv_n;
        store_c(ob + ra * QKV_LD + 2 * HEADS * DH, ob + rb * QKV_LD + 2 * HEADS * DH, oka, okb, t, acc);
    }
}

// ------------------------------------------------------------------------------------------------ stand-alone check
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

}  // namespace gradk

}  // namespace la_mma

cudaError_t linear_attention_bwd_mma_run(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int n, float* kmax, float* ksum,
                                         float* cd, cudaStream_t s) {
    using namespace la_mma;
    const int bh = B * HEADS;
    const float scale = 0.17677669529663687f;      // 32^-0.5
    ctxk::la_ctx_mma_kernel<<<bh, CTX_THREADS, 0, s>>>(qkv, dout, n, scale, kmax, ksum, cd);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int chunks = (n + 127) / 128;             // 128 pixels per CTA: four warps of 16-pixel tiles, two tiles each
    gradk::la_grad_mma_kernel<<<dim3(chunks, bh), 128, 0, s>>>(qkv, dout, n, scale, kmax, ksum, cd, dqkv);
    return cudaGetLastError();
}

}  // namespace hd
