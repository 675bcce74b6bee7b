// DRAFT for round 2 -- NOT built into libhicdiff_b200.so, NOT yet run on a GPU (written after the round's GPU budget was spent).
//
// Tensor-core form of the first two passes of the linear-attention backward (hicdiff_b200/csrc/attention_bwd.cu:
// la_kstats_kernel + la_ctx_kernel + la_reduce_kernel = 0.2 + 1.05 + 0.05 ms of the 13.4 ms Unet training step, the ctx / dctx
// accumulation running on shuffles and fp32 FMAs).  One kernel, one CTA per (image, head), adapted from the FORWARD context kernel
// (hicdiff_b200/csrc/attention.cu: linattn_context_kernel, 64 us at the 64x64 level), which already is an mma.sync / ldmatrix
// pipeline for ctx = ks^T vs:
//     kmax[d] = max_n k[n,d]            ksum[d] = sum_n p[n,d],  p = bf16(exp(k - kmax))
//     ctx[d,e]  = (sum_n p[n,d] v[n,e]) / ksum[d] / n                                  (fp32, [d][e], no 32^-0.5)
//     dctx[d,e] = sum_n qs[n,d] dout[n,e],  qs = bf16(softmax_d(q[n,:]) * 32^-0.5)     (fp32, [d][e])
// i.e. exactly the `kmax`, `ksum`, `cd` arrays la_grad_kernel consumes.  Deterministic: fixed-order cross-warp reduction.
//
// Stand-alone check (next round, on the GPU box):
//     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/la_ctx_mma experimental/la_ctx_mma.cu
//     /tmp/la_ctx_mma
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

using bf16 = __nv_bfloat16;
constexpr int HEADS = 4, DH = 32, QKV_LD = 3 * HEADS * DH, OUT_LD = HEADS * DH;
constexpr int CTX_THREADS = 256, CTX_WARPS = CTX_THREADS / 32;
constexpr int CTX_TILE = 128;           // pixels staged per iteration (16 per warp)
constexpr int CTX_PITCH = 40;           // bf16 elements per staged row (32 + 8 pad) = 80 bytes: conflict-free ldmatrix

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

// ------------------------------------------------------------------------------------------------ stand-alone check
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
    const int B = 2;
    const float scale = 0.17677669529663687f;
    int bad = 0;
    for (int n : {64, 200, 1024}) {                      // one partial tile, a ragged tail, whole tiles
        const int bh = B * HEADS;
        std::vector<float> qkv(static_cast<size_t>(B) * n * QKV_LD), dout(static_cast<size_t>(B) * n * OUT_LD);
        uint32_t seed = 777u + n;
        auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (static_cast<float>(seed >> 8) / 8388608.0f) - 1.0f; };
        for (auto& v : qkv) v = bf(2.0f * rnd());
        for (auto& v : dout) v = bf(0.2f * rnd());
        std::vector<bf16> hq(qkv.size()), hd(dout.size());
        for (size_t i = 0; i < qkv.size(); ++i) hq[i] = __float2bfloat16(qkv[i]);
        for (size_t i = 0; i < dout.size(); ++i) hd[i] = __float2bfloat16(dout[i]);
        std::vector<double> wctx(static_cast<size_t>(bh) * DH * DH, 0.0), wdctx(wctx.size(), 0.0), wmax(bh * DH), wsum(bh * DH, 0.0);
        for (int b = 0; b < B; ++b)
            for (int h = 0; h < HEADS; ++h) {
                const int x = b * HEADS + h;
                for (int d = 0; d < DH; ++d) {
                    double m = -1e30;
                    for (int p = 0; p < n; ++p) m = std::fmax(m, qkv[(static_cast<size_t>(b) * n + p) * QKV_LD + HEADS * DH + h * DH + d]);
                    wmax[x * DH + d] = m;
                    for (int p = 0; p < n; ++p) wsum[x * DH + d] += std::exp(qkv[(static_cast<size_t>(b) * n + p) * QKV_LD + HEADS * DH + h * DH + d] - m);
                }
                for (int p = 0; p < n; ++p) {
                    const float* row = &qkv[(static_cast<size_t>(b) * n + p) * QKV_LD + h * DH];
                    const float* dor = &dout[(static_cast<size_t>(b) * n + p) * OUT_LD + h * DH];
                    double y[DH], m = -1e30, s = 0;
                    for (int d = 0; d < DH; ++d) m = std::fmax(m, row[d]);
                    for (int d = 0; d < DH; ++d) { y[d] = std::exp(row[d] - m); s += y[d]; }
                    for (int d = 0; d < DH; ++d) {
                        const double qs = y[d] / s * scale;
                        const double ks = std::exp(row[HEADS * DH + d] - wmax[x * DH + d]) / wsum[x * DH + d];
                        for (int e = 0; e < DH; ++e) {
                            wctx[(static_cast<size_t>(x) * DH + d) * DH + e] += ks * row[2 * HEADS * DH + e] / n;
                            wdctx[(static_cast<size_t>(x) * DH + d) * DH + e] += qs * dor[e];
                        }
                    }
                }
            }
        bf16 *dq, *dd;
        float *dcd, *dkm, *dks;
        CK(cudaMalloc(&dq, hq.size() * 2)); CK(cudaMalloc(&dd, hd.size() * 2));
        CK(cudaMalloc(&dcd, static_cast<size_t>(bh) * 2 * DH * DH * 4)); CK(cudaMalloc(&dkm, bh * DH * 4)); CK(cudaMalloc(&dks, bh * DH * 4));
        CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dd, hd.data(), hd.size() * 2, cudaMemcpyHostToDevice));
        la_ctx_mma_kernel<<<bh, CTX_THREADS>>>(dq, dd, n, scale, dkm, dks, dcd);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> cd(static_cast<size_t>(bh) * 2 * DH * DH), km(bh * DH), ks(bh * DH);
        CK(cudaMemcpy(cd.data(), dcd, cd.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(km.data(), dkm, km.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ks.data(), dks, ks.size() * 4, cudaMemcpyDeviceToHost));
        double ec = 0, wc = 0, ed = 0, wd = 0, em = 0, es = 0;
        for (int x = 0; x < bh; ++x) {
            for (int i = 0; i < DH * DH; ++i) {
                const double c = cd[static_cast<size_t>(x) * 2 * DH * DH + i] - wctx[static_cast<size_t>(x) * DH * DH + i];
                const double d = cd[(static_cast<size_t>(x) * 2 + 1) * DH * DH + i] - wdctx[static_cast<size_t>(x) * DH * DH + i];
                ec += c * c; wc += wctx[static_cast<size_t>(x) * DH * DH + i] * wctx[static_cast<size_t>(x) * DH * DH + i];
                ed += d * d; wd += wdctx[static_cast<size_t>(x) * DH * DH + i] * wdctx[static_cast<size_t>(x) * DH * DH + i];
            }
            for (int d = 0; d < DH; ++d) {
                em = std::fmax(em, std::fabs(km[x * DH + d] - wmax[x * DH + d]));
                es = std::fmax(es, std::fabs(ks[x * DH + d] - wsum[x * DH + d]) / wsum[x * DH + d]);
            }
        }
        const double rc = std::sqrt(ec / wc), rd = std::sqrt(ed / wd);
        printf("n = %4d: ctx rel-RMS %.3e  dctx rel-RMS %.3e  kmax max err %.3e  ksum max rel err %.3e\n", n, rc, rd, em, es);
        if (!(rc <= 5e-3 && rd <= 5e-3 && em == 0.0 && es <= 5e-3)) bad = 1;
        cudaFree(dq); cudaFree(dd); cudaFree(dcd); cudaFree(dkm); cudaFree(dks);
    }
    printf(bad ? "FAILED\n" : "ok\n");
    return bad;
}
