// Backward of the Unet's attention cores (building blocks), on the NHWC qkv tensor of the to_qkv conv:
// qkv[b, n, 384] = [ q(4 x 32) | k(4 x 32) | v(4 x 32) ], d_out[b, n, 128] -> d_qkv[b, n, 384].
//
// LinearAttention.forward  /root/reference/src/hicdiff_condition.py:212-227, per (image, head), d = e = 32:
//     qs = softmax_d(q) * s, ks = softmax_n(k), vs = v / n, ctx[d,e] = sum_n ks[n,d] vs[n,e], out[n,e] = sum_d qs[n,d] ctx[d,e]
//   backward:  dctx[d,e] = sum_n qs[n,d] dout[n,e]          dqs[n,d] = sum_e dout[n,e] ctx[d,e]
//              dks[n,d]  = sum_e vs[n,e] dctx[d,e]          dv[n,e]  = (1/n) sum_d ks[n,d] dctx[d,e]
//              dq = s * softmax_d(q) o (dqs - <dqs, softmax_d(q)>)       dk = ks o (dks - t),  t[d] = sum_n dks[n,d] ks[n,d]
//   four streaming passes over the pixels (column max / sum of k; ctx and dctx; t; the three gradients), one warp per pixel
//   with lane = d (or e), 32 x 32 matrices in registers, partial sums per pixel chunk combined in a fixed order.
// Attention.forward (8x8 only, n = 64)  :239-251: P = softmax_j(s q k^T), O = P v; the textbook backward in shared memory,
//   one CTA per (image, head).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

constexpr int HEADS = 4;
constexpr int DH = 32;
constexpr int QKV_LD = 3 * HEADS * DH;
constexpr int OUT_LD = HEADS * DH;
constexpr int LAB_SPLIT = 8;            // pixel chunks per (image, head)

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- pass 1: column max and sum_n exp(k - max) per (b, h, d).  grid B * HEADS, 256 threads = 8 warps striding the pixels
__global__ void __launch_bounds__(256)
la_kstats_kernel(const bf16* __restrict__ qkv, int n, float* __restrict__ kmax, float* __restrict__ ksum) {
    __shared__ float s_m[8][DH];
    const int bh = blockIdx.x, b = bh / HEADS, h = bh % HEADS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bf16* kp = qkv + static_cast<size_t>(b) * n * QKV_LD + HEADS * DH + h * DH + lane;
    float m = -INFINITY;
    for (int p = warp; p < n; p += 8) m = fmaxf(m, __bfloat162float(kp[static_cast<size_t>(p) * QKV_LD]));
    s_m[warp][lane] = m;
    __syncthreads();
    m = s_m[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w][lane]);
    __syncthreads();
    float s = 0.f;
    for (int p = warp; p < n; p += 8) s += __expf(__bfloat162float(kp[static_cast<size_t>(p) * QKV_LD]) - m);
    s_m[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_m[w][lane];
        kmax[bh * DH + lane] = m;
        ksum[bh * DH + lane] = t;
    }
}

// ---- pass 2: partial ctx[d][e] and dctx[d][e] over one pixel chunk.  grid (LAB_SPLIT, B * HEADS); lane = e, registers hold column e
__global__ void __launch_bounds__(256)
la_ctx_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, int n, float scale, const float* __restrict__ kmax,
              const float* __restrict__ ksum, float* __restrict__ part) {
    extern __shared__ float la_smem[];
    float (*s_acc)[2][DH][DH + 1] = reinterpret_cast<float (*)[2][DH][DH + 1]>(la_smem);     // [8 warps][2][d][e]
    const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (n + LAB_SPLIT - 1) / LAB_SPLIT;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    const bf16* base = qkv + static_cast<size_t>(b) * n * QKV_LD + h * DH + lane;
    const bf16* dob = dout + static_cast<size_t>(b) * n * OUT_LD + h * DH + lane;
    const float km = kmax[bh * DH + lane], kinv = 1.0f / ksum[bh * DH + lane];
    const float inv_n = 1.0f / static_cast<float>(n);
    float ctx[DH], dctx[DH];     // [d] for this lane's e
#pragma unroll
    for (int d = 0; d < DH; ++d) { ctx[d] = 0.f; dctx[d] = 0.f; }
    // ncu (source page): 36 % of the stall samples sit on the first use of the 2-byte loads -- with one pixel per warp in flight
    // an SM has ~4 KB outstanding, i.e. 0.6 TB/s for the whole GPU, exactly what the kernel reached.  Eight pixels' loads are
    // issued before any of them is used.
    constexpr int PF = 8;
    for (int pb = p0 + warp; pb < p1; pb += 8 * PF) {
        bf16 qa[PF], ka[PF], va[PF], ga[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int p = min(pb + 8 * u, p1 - 1);
            const size_t ro = static_cast<size_t>(p) * QKV_LD;
            qa[u] = base[ro];
            ka[u] = base[ro + HEADS * DH];
            va[u] = base[ro + 2 * HEADS * DH];
            ga[u] = dob[static_cast<size_t>(p) * OUT_LD];
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (pb + 8 * u >= p1) break;
            const float q = __bfloat162float(qa[u]);
            const float v = __bfloat162float(va[u]) * inv_n, go = __bfloat162float(ga[u]);
            const float qe = __expf(q - warp_max(q));
            const float qs = qe / warp_sum(qe) * scale;            // lane = d
            const float ks = __expf(__bfloat162float(ka[u]) - km) * kinv;                // lane = d
#pragma unroll
            for (int d = 0; d < DH; ++d) {
                ctx[d] = fmaf(__shfl_sync(0xffffffffu, ks, d), v, ctx[d]);       // lane = e
                dctx[d] = fmaf(__shfl_sync(0xffffffffu, qs, d), go, dctx[d]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) { s_acc[warp][0][d][lane] = ctx[d]; s_acc[warp][1][d][lane] = dctx[d]; }
    __syncthreads();
    float* o = part + (static_cast<size_t>(bh) * LAB_SPLIT + blockIdx.x) * 2 * DH * DH;
    for (int i = threadIdx.x; i < 2 * DH * DH; i += 256) {
        const int m = i / (DH * DH), d = (i / DH) % DH, e = i % DH;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_acc[w][m][d][e];
        o[i] = t;
    }
}

// sum the chunk partials: out[bh][i] = sum_c part[bh][c][i]; count values per bh
__global__ void la_reduce_kernel(const float* __restrict__ part, int count, float* __restrict__ out) {
    const int bh = blockIdx.x;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        float t = 0.f;
        for (int c = 0; c < LAB_SPLIT; ++c) t += part[(static_cast<size_t>(bh) * LAB_SPLIT + c) * count + i];
        out[static_cast<size_t>(bh) * count + i] = t;
    }
}

// ---- pass 3: partial t[d] = sum_n dks[n,d] ks[n,d].  lane = d, registers hold row d of dctx
__global__ void __launch_bounds__(256)
la_t_kernel(const bf16* __restrict__ qkv, int n, const float* __restrict__ kmax, const float* __restrict__ ksum,
            const float* __restrict__ cd, float* __restrict__ part) {
    __shared__ float s_t[8][DH];
    const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (n + LAB_SPLIT - 1) / LAB_SPLIT;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    const bf16* base = qkv + static_cast<size_t>(b) * n * QKV_LD + h * DH + lane;
    const float km = kmax[bh * DH + lane], kinv = 1.0f / ksum[bh * DH + lane];
    const float inv_n = 1.0f / static_cast<float>(n);
    // the 32 x 32 matrix goes through shared memory: a direct per-lane row read is 32 strided sectors per instruction, repeated
    // by every warp of 2048 CTAs -- ~70 us of L2 requests per launch whatever n is (measured at the 8x8 level)
    __shared__ float s_d[DH][DH + 1];
    const float* dctx = cd + (static_cast<size_t>(bh) * 2 + 1) * DH * DH;
    for (int i = threadIdx.x; i < DH * DH; i += 256) s_d[i / DH][i % DH] = dctx[i];
    __syncthreads();
    float row[DH];
#pragma unroll
    for (int e = 0; e < DH; ++e) row[e] = s_d[lane][e];
    float t = 0.f;
    constexpr int PF = 8;                                  // eight pixels' loads in flight per warp (see la_ctx_kernel)
    for (int pb = p0 + warp; pb < p1; pb += 8 * PF) {
        bf16 ka[PF], va[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const size_t ro = static_cast<size_t>(min(pb + 8 * u, p1 - 1)) * QKV_LD;
            ka[u] = base[ro + HEADS * DH];
            va[u] = base[ro + 2 * HEADS * DH];
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (pb + 8 * u >= p1) break;
            const float ks = __expf(__bfloat162float(ka[u]) - km) * kinv;
            const float v = __bfloat162float(va[u]) * inv_n;
            float dks = 0.f;
#pragma unroll
            for (int e = 0; e < DH; ++e) dks = fmaf(__shfl_sync(0xffffffffu, v, e), row[e], dks);
            t = fmaf(dks, ks, t);
        }
    }
    s_t[warp][lane] = t;
    __syncthreads();
    if (warp == 0) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += s_t[w][lane];
        part[(static_cast<size_t>(bh) * LAB_SPLIT + blockIdx.x) * DH + lane] = a;
    }
}

// ---- pass 4: the gradients.  grid (chunks, B * HEADS); ONE THREAD PER PIXEL: its q / k / v / d_out rows (32 values each) live in
// registers, the 32 x 32 matrices are read from shared memory as 16-byte broadcasts, so there is no cross-lane traffic at all.
// (The first version -- one warp per pixel, lane = d, 96 shuffles per pixel -- ran at 12 % occupancy on 159 registers and was
// latency-bound: profiles/r01_train_unet_linattn_bwd_ncu.md.)
__device__ __forceinline__ void load_row32(const bf16* p, float (&v)[DH]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
        float2 t;
        t = ptx::unpack_bf16x2(u.x); v[8 * i] = t.x; v[8 * i + 1] = t.y;
        t = ptx::unpack_bf16x2(u.y); v[8 * i + 2] = t.x; v[8 * i + 3] = t.y;
        t = ptx::unpack_bf16x2(u.z); v[8 * i + 4] = t.x; v[8 * i + 5] = t.y;
        t = ptx::unpack_bf16x2(u.w); v[8 * i + 6] = t.x; v[8 * i + 7] = t.y;
    }
}
__device__ __forceinline__ void store_row32(bf16* p, const float (&v)[DH]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = ptx::pack_bf16x2(v[8 * i], v[8 * i + 1]); u.y = ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        reinterpret_cast<uint4*>(p)[i] = u;
    }
}
// out[r] = sum_c m[r][c] * x[c], m row-major [32][32] in shared memory (every thread reads the same addresses: broadcasts)
__device__ __forceinline__ void matvec32(const float* __restrict__ m, const float (&x)[DH], float (&out)[DH]) {
#pragma unroll
    for (int r = 0; r < DH; ++r) {
        const float4* row = reinterpret_cast<const float4*>(m + r * DH);
        float acc = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < DH / 4; ++c4) {
            const float4 w = row[c4];
            acc = fmaf(w.x, x[4 * c4], acc); acc = fmaf(w.y, x[4 * c4 + 1], acc);
            acc = fmaf(w.z, x[4 * c4 + 2], acc); acc = fmaf(w.w, x[4 * c4 + 3], acc);
        }
        out[r] = acc;
    }
}

__global__ void __launch_bounds__(128)
la_grad_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, int n, float scale, const float* __restrict__ kmax,
               const float* __restrict__ ksum, const float* __restrict__ cd, const float* __restrict__ tvec, bf16* __restrict__ dqkv) {
    __shared__ __align__(16) float s_ctx[DH * DH], s_dctx[DH * DH], s_dctxT[DH * DH];
    __shared__ float s_km[DH], s_kinv[DH], s_t[DH];
    const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
    for (int i = threadIdx.x; i < DH * DH; i += blockDim.x) {
        const float c = cd[static_cast<size_t>(bh) * 2 * DH * DH + i];
        const float d = cd[(static_cast<size_t>(bh) * 2 + 1) * DH * DH + i];
        s_ctx[i] = c;                                   // [d][e]
        s_dctx[i] = d;                                  // [d][e]
        s_dctxT[(i % DH) * DH + i / DH] = d;            // [e][d]
    }
    if (threadIdx.x < DH) {
        s_km[threadIdx.x] = kmax[bh * DH + threadIdx.x];
        s_kinv[threadIdx.x] = 1.0f / ksum[bh * DH + threadIdx.x];
        s_t[threadIdx.x] = tvec[bh * DH + threadIdx.x];
    }
    __syncthreads();
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    const float inv_n = 1.0f / static_cast<float>(n);
    const bf16* base = qkv + static_cast<size_t>(b) * n * QKV_LD + h * DH;
    const bf16* dob = dout + static_cast<size_t>(b) * n * OUT_LD + h * DH;
    bf16* ob = dqkv + static_cast<size_t>(b) * n * QKV_LD + h * DH;
    for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
        const size_t ro = static_cast<size_t>(p) * QKV_LD;
        float x[DH], y[DH], o[DH];
        // dq = s * softmax_d(q) o (dqs - <dqs, softmax_d(q)>),  dqs[d] = sum_e dout[e] ctx[d][e]
        load_row32(dob + static_cast<size_t>(p) * OUT_LD, x);       // d_out[e]
        matvec32(s_ctx, x, o);                                       // dqs[d]
        load_row32(base + ro, y);                                    // q[d]
        float m = y[0];
#pragma unroll
        for (int d = 1; d < DH; ++d) m = fmaxf(m, y[d]);
        float sum = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) { y[d] = __expf(y[d] - m); sum += y[d]; }
        const float inv = 1.0f / sum;
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) { y[d] *= inv; dot = fmaf(o[d], y[d], dot); }
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = scale * y[d] * (o[d] - dot);
        store_row32(ob + ro, o);
        // dk = ks o (dks - t),  dks[d] = sum_e vs[e] dctx[d][e]
        load_row32(base + ro + 2 * HEADS * DH, x);                  // v[e]
#pragma unroll
        for (int e = 0; e < DH; ++e) x[e] *= inv_n;
        matvec32(s_dctx, x, o);                                      // dks[d]
        load_row32(base + ro + HEADS * DH, y);                       // k[d]
#pragma unroll
        for (int d = 0; d < DH; ++d) { y[d] = __expf(y[d] - s_km[d]) * s_kinv[d]; o[d] = y[d] * (o[d] - s_t[d]); }
        store_row32(ob + ro + HEADS * DH, o);
        // dv[e] = (1 / n) sum_d ks[d] dctx[d][e]
        matvec32(s_dctxT, y, o);
#pragma unroll
        for (int e = 0; e < DH; ++e) o[e] *= inv_n;
        store_row32(ob + ro + 2 * HEADS * DH, o);
    }
}

// ------------------------------------------------------------------------------------------------ full attention (n = 64)
constexpr int FA_N = 64;
__global__ void __launch_bounds__(256)
full_attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, float scale, bf16* __restrict__ dqkv) {
    extern __shared__ float fa_smem[];
    float (*sq)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(fa_smem);
    float (*sk)[DH + 1] = sq + FA_N;
    float (*sv)[DH + 1] = sk + FA_N;
    float (*sdo)[DH + 1] = sv + FA_N;
    float (*sP)[FA_N + 1] = reinterpret_cast<float (*)[FA_N + 1]>(fa_smem + 4 * FA_N * (DH + 1));
    float (*sdS)[FA_N + 1] = sP + FA_N;
    const int bh = blockIdx.x, b = bh / HEADS, h = bh % HEADS;
    const int tid = threadIdx.x;
    const bf16* base = qkv + static_cast<size_t>(b) * FA_N * QKV_LD + h * DH;
    const bf16* dob = dout + static_cast<size_t>(b) * FA_N * OUT_LD + h * DH;
    for (int i = tid; i < FA_N * DH; i += 256) {
        const int p = i / DH, d = i % DH;
        sq[p][d] = __bfloat162float(base[static_cast<size_t>(p) * QKV_LD + d]);
        sk[p][d] = __bfloat162float(base[static_cast<size_t>(p) * QKV_LD + HEADS * DH + d]);
        sv[p][d] = __bfloat162float(base[static_cast<size_t>(p) * QKV_LD + 2 * HEADS * DH + d]);
        sdo[p][d] = __bfloat162float(dob[static_cast<size_t>(p) * OUT_LD + d]);
    }
    __syncthreads();
    // S = scale q k^T and dP = dO v^T
    for (int i = tid; i < FA_N * FA_N; i += 256) {
        const int r = i / FA_N, c = i % FA_N;
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) { s = fmaf(sq[r][d], sk[c][d], s); dp = fmaf(sdo[r][d], sv[c][d], dp); }
        sP[r][c] = s * scale;
        sdS[r][c] = dp;
    }
    __syncthreads();
    // row softmax and dS = P o (dP - <dP, P>): one warp per row (8 warps, 8 rows each)
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < FA_N; r += 8) {
        const float a0 = sP[r][lane], a1 = sP[r][lane + 32];
        const float m = fmaxf(a0, a1);
        float mx = m;
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e0 = __expf(a0 - mx), e1 = __expf(a1 - mx);
        const float inv = 1.0f / warp_sum(e0 + e1);
        const float p0 = e0 * inv, p1 = e1 * inv;
        const float dot = warp_sum(sdS[r][lane] * p0 + sdS[r][lane + 32] * p1);
        sP[r][lane] = p0; sP[r][lane + 32] = p1;
        sdS[r][lane] = p0 * (sdS[r][lane] - dot);
        sdS[r][lane + 32] = p1 * (sdS[r][lane + 32] - dot);
    }
    __syncthreads();
    bf16* ob = dqkv + static_cast<size_t>(b) * FA_N * QKV_LD + h * DH;
    for (int i = tid; i < FA_N * DH; i += 256) {
        const int p = i / DH, d = i % DH;
        float dq = 0.f, dk = 0.f, dv = 0.f;
#pragma unroll 8
        for (int j = 0; j < FA_N; ++j) {
            dq = fmaf(sdS[p][j], sk[j][d], dq);      // dq[i] = scale sum_j dS[i,j] k[j]
            dk = fmaf(sdS[j][p], sq[j][d], dk);      // dk[j] = scale sum_i dS[i,j] q[i]
            dv = fmaf(sP[j][p], sdo[j][d], dv);      // dv[j] = sum_i P[i,j] dO[i]
        }
        ob[static_cast<size_t>(p) * QKV_LD + d] = __float2bfloat16(dq * scale);
        ob[static_cast<size_t>(p) * QKV_LD + HEADS * DH + d] = __float2bfloat16(dk * scale);
        ob[static_cast<size_t>(p) * QKV_LD + 2 * HEADS * DH + d] = __float2bfloat16(dv);
    }
}

}  // namespace

size_t linattn_bwd_scratch_floats(int B) {
    const size_t bh = static_cast<size_t>(B) * HEADS;
    return bh * DH * 2 /*kmax, ksum*/ + bh * LAB_SPLIT * 2 * DH * DH /*ctx parts*/ + bh * 2 * DH * DH /*ctx, dctx*/ +
           bh * LAB_SPLIT * DH /*t parts*/ + bh * DH /*t*/ + 64;
}

cudaError_t linear_attention_bwd_run(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int n, float* scratch, cudaStream_t s) {
    const int bh = B * HEADS;
    const float scale = 0.17677669529663687f;      // 32^-0.5
    float* kmax = scratch;
    float* ksum = kmax + static_cast<size_t>(bh) * DH;
    float* cpart = ksum + static_cast<size_t>(bh) * DH;
    float* cd = cpart + static_cast<size_t>(bh) * LAB_SPLIT * 2 * DH * DH;
    // tensor-core form (attention_bwd_mma.cu): two launches instead of six.  HD_LA_BWD_LEGACY=1 keeps the CUDA-core passes below.
    static const bool legacy = [] { const char* v = getenv("HD_LA_BWD_LEGACY"); return v && atoi(v) != 0; }();
    if (!legacy) return linear_attention_bwd_mma_run(qkv, dout, dqkv, B, n, kmax, ksum, cd, s);
    float* tpart = cd + static_cast<size_t>(bh) * 2 * DH * DH;
    float* tv = tpart + static_cast<size_t>(bh) * LAB_SPLIT * DH;
    la_kstats_kernel<<<bh, 256, 0, s>>>(qkv, n, kmax, ksum);
    constexpr size_t ctx_smem = static_cast<size_t>(8) * 2 * DH * (DH + 1) * sizeof(float);
    static bool ctx_attr = false;     // once per process (one process per GPU); never inside a stream capture (the first step is eager)
    if (!ctx_attr) {
        cudaError_t ae = cudaFuncSetAttribute(la_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ctx_smem));
        if (ae != cudaSuccess) return ae;
        ctx_attr = true;
    }
    la_ctx_kernel<<<dim3(LAB_SPLIT, bh), 256, ctx_smem, s>>>(qkv, dout, n, scale, kmax, ksum, cpart);
    la_reduce_kernel<<<bh, 256, 0, s>>>(cpart, 2 * DH * DH, cd);
    la_t_kernel<<<dim3(LAB_SPLIT, bh), 256, 0, s>>>(qkv, n, kmax, ksum, cd, tpart);
    la_reduce_kernel<<<bh, 32, 0, s>>>(tpart, DH, tv);
    const int chunks = (n + 127) / 128;     // one pixel per thread, 128 threads per CTA
    la_grad_kernel<<<dim3(chunks, bh), 128, 0, s>>>(qkv, dout, n, scale, kmax, ksum, cd, tv, dqkv);
    return cudaGetLastError();
}

cudaError_t full_attention_bwd_run(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int n, cudaStream_t s) {
    if (n != FA_N) return cudaErrorInvalidValue;
    constexpr size_t fa_bytes = (static_cast<size_t>(4) * FA_N * (DH + 1) + 2 * FA_N * (FA_N + 1)) * sizeof(float);
    static bool fa_attr = false;
    if (!fa_attr) {
        cudaError_t ae = cudaFuncSetAttribute(full_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fa_bytes));
        if (ae != cudaSuccess) return ae;
        fa_attr = true;
    }
    full_attention_bwd_kernel<<<B * HEADS, 256, fa_bytes, s>>>(qkv, dout, 0.17677669529663687f, dqkv);
    return cudaGetLastError();
}

}  // namespace hd
