// DRAFT for round 2 -- NOT built into libhicdiff_b200.so, NOT yet run on a GPU (written after the round's GPU budget was spent).
//
// Tensor-core form of `la_grad_kernel` (hicdiff_b200/csrc/attention_bwd.cu): the third pass of the linear-attention backward
// (LinearAttention.forward, src/hicdiff_condition.py:212-227).  The shipped kernel does three 32 x 32 mat-vecs per (pixel, head)
// on the fp32 FMA pipe (3.2 GFMA at the 64x64 level: 354 us, ~26 % of the FMA rate).  Here a warp owns 16 pixels and the three
// products are mma.sync.m16n8k16 bf16 tiles ([16 px x 32] x [32 x 32], fp32 accumulate):
//     dqs[n,d] = sum_e dout[n,e] ctx[d,e]      dks[n,d] = sum_e v[n,e] dctx[d,e] / n      dv[n,e] = sum_d ks[n,d] dctx[d,e] / n
// The A fragment of a [16 x 32] bf16 tile and the C fragment of a [16 x 32] fp32 result hold the SAME (row, column) elements in
// every thread (row g / g + 8, columns 8 j + 2 t, + 1), so the softmax / product epilogues are thread-local and only the three
// per-row reductions of the q softmax cross the four lanes of a quad.  The matrices are rounded to bf16 once per CTA (ctx and
// dctx come from fp32 partial sums; 2^-9 relative, the same order as the bf16 rounding of the outputs).
//
// Stand-alone check (next round, on the GPU box):
//     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/la_grad_mma experimental/la_grad_mma.cu
//     /tmp/la_grad_mma            # compares against a CPU evaluation of the same formulas, prints max / rms errors and the time
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

using bf16 = __nv_bfloat16;
constexpr int HEADS = 4, DH = 32, QKV_LD = 3 * HEADS * DH, OUT_LD = HEADS * DH;

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
                   const float* __restrict__ ksum, const float* __restrict__ cd, const float* __restrict__ tvec, bf16* __restrict__ dqkv) {
    __shared__ __align__(16) bf16 s_m[3][DH][DH];         // [0] ctx[d][e], [1] dctx[d][e], [2] dctx^T [e][d]: M[n][k] of the three products
    __shared__ float s_km[DH], s_kinv[DH], s_t[DH];
    const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
    for (int i = threadIdx.x; i < DH * DH; i += blockDim.x) {
        const float c = cd[static_cast<size_t>(bh) * 2 * DH * DH + i];
        const float d = cd[(static_cast<size_t>(bh) * 2 + 1) * DH * DH + i];
        s_m[0][i / DH][i % DH] = __float2bfloat16(c);
        s_m[1][i / DH][i % DH] = __float2bfloat16(d);
        s_m[2][i % DH][i / DH] = __float2bfloat16(d);
    }
    if (threadIdx.x < DH) {
        s_km[threadIdx.x] = kmax[bh * DH + threadIdx.x];
        s_kinv[threadIdx.x] = 1.0f / ksum[bh * DH + threadIdx.x];
        s_t[threadIdx.x] = tvec[bh * DH + threadIdx.x];
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

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
    const int B = 2, n = 1000;                       // ragged on purpose: n is not a multiple of 16 or of the chunk size
    const int bh = B * HEADS;
    const float scale = 0.17677669529663687f;
    std::vector<float> qkv(static_cast<size_t>(B) * n * QKV_LD), dout(static_cast<size_t>(B) * n * OUT_LD), cd(static_cast<size_t>(bh) * 2 * DH * DH),
        kmax(bh * DH), ksum(bh * DH), tv(bh * DH);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (static_cast<float>(seed >> 8) / 8388608.0f) - 1.0f; };
    for (auto& v : qkv) v = bf(2.0f * rnd());
    for (auto& v : dout) v = bf(0.2f * rnd());
    for (auto& v : cd) v = 0.05f * rnd();
    for (auto& v : kmax) v = 2.0f + 0.1f * rnd();
    for (auto& v : ksum) v = 300.0f + 50.0f * rnd();
    for (auto& v : tv) v = 0.01f * rnd();
    std::vector<bf16> hq(qkv.size()), hd(dout.size());
    for (size_t i = 0; i < qkv.size(); ++i) hq[i] = __float2bfloat16(qkv[i]);
    for (size_t i = 0; i < dout.size(); ++i) hd[i] = __float2bfloat16(dout[i]);
    // CPU evaluation of the same formulas in double
    std::vector<double> want(qkv.size());
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < HEADS; ++h) {
            const int x = b * HEADS + h;
            const float* ctx = &cd[static_cast<size_t>(x) * 2 * DH * DH];
            const float* dctx = ctx + DH * DH;
            for (int p = 0; p < n; ++p) {
                const float* row = &qkv[(static_cast<size_t>(b) * n + p) * QKV_LD + h * DH];
                const float* dor = &dout[(static_cast<size_t>(b) * n + p) * OUT_LD + h * DH];
                double* wr = &want[(static_cast<size_t>(b) * n + p) * QKV_LD + h * DH];
                double y[DH], dqs[DH], ks[DH], m = -1e30, s = 0, dot = 0;
                for (int d = 0; d < DH; ++d) m = std::fmax(m, row[d]);
                for (int d = 0; d < DH; ++d) { y[d] = std::exp(row[d] - m); s += y[d]; }
                for (int d = 0; d < DH; ++d) {
                    y[d] /= s;
                    dqs[d] = 0;
                    for (int e = 0; e < DH; ++e) dqs[d] += static_cast<double>(dor[e]) * ctx[d * DH + e];
                    dot += dqs[d] * y[d];
                }
                for (int d = 0; d < DH; ++d) wr[d] = scale * y[d] * (dqs[d] - dot);
                for (int d = 0; d < DH; ++d) {
                    ks[d] = std::exp(row[HEADS * DH + d] - kmax[x * DH + d]) / ksum[x * DH + d];
                    double dks = 0;
                    for (int e = 0; e < DH; ++e) dks += static_cast<double>(row[2 * HEADS * DH + e]) / n * dctx[d * DH + e];
                    wr[HEADS * DH + d] = ks[d] * (dks - tv[x * DH + d]);
                }
                for (int e = 0; e < DH; ++e) {
                    double dv = 0;
                    for (int d = 0; d < DH; ++d) dv += ks[d] * dctx[d * DH + e];
                    wr[2 * HEADS * DH + e] = dv / n;
                }
            }
        }
    bf16 *dq, *dd, *dg;
    float *dcd, *dkm, *dks, *dtv;
    CK(cudaMalloc(&dq, hq.size() * 2)); CK(cudaMalloc(&dd, hd.size() * 2)); CK(cudaMalloc(&dg, hq.size() * 2));
    CK(cudaMalloc(&dcd, cd.size() * 4)); CK(cudaMalloc(&dkm, kmax.size() * 4)); CK(cudaMalloc(&dks, ksum.size() * 4)); CK(cudaMalloc(&dtv, tv.size() * 4));
    CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dd, hd.data(), hd.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcd, cd.data(), cd.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dkm, kmax.data(), kmax.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dks, ksum.data(), ksum.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dtv, tv.data(), tv.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dg, 0xff, hq.size() * 2));
    const int chunks = (n + 127) / 128;
    la_grad_mma_kernel<<<dim3(chunks, bh), 128>>>(dq, dd, n, scale, dkm, dks, dcd, dtv, dg);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<bf16> got(hq.size());
    CK(cudaMemcpy(got.data(), dg, got.size() * 2, cudaMemcpyDeviceToHost));
    const char* names[3] = {"dq", "dk", "dv"};
    int bad = 0;
    for (int part = 0; part < 3; ++part) {
        double se = 0, sw = 0, mx = 0;
        for (size_t r = 0; r < static_cast<size_t>(B) * n; ++r)
            for (int c = 0; c < HEADS * DH; ++c) {
                const size_t i = r * QKV_LD + part * HEADS * DH + c;
                const double e = static_cast<double>(__bfloat162float(got[i])) - want[i];
                se += e * e; sw += want[i] * want[i]; mx = std::fmax(mx, std::fabs(e));
            }
        const double rel = std::sqrt(se / (sw + 1e-300));
        printf("%s: rel-RMS %.3e  max abs err %.3e\n", names[part], rel, mx);
        if (!(rel <= 1e-2)) bad = 1;
    }
    printf(bad ? "FAILED\n" : "ok\n");
    return bad;
}
