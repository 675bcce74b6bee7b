// Backward of the Unet's normalisation layers (building blocks of the Unet training step):
//
//   Block           s = SiLU( GN_8(y) * (scale + 1) + shift ),  GN_8(y) = (y - mu_g) * r_g * gamma + beta
//                   /root/reference/src/hicdiff_condition.py:155-171 (nn.GroupNorm(8, C), eps 1e-5, biased variance)
//   LayerNorm       z = (x - mu) * rsqrt(var + eps) * g over the channel dimension of every pixel      :99-108
//   WeightStandardizedConv2d   w~ = (w - mu_o) * rsqrt(var_o + eps) per output channel                :84-97
//
// GroupNorm backward, with dh = ds * SiLU'(h), xhat = (y - mu) * r and per (sample, channel) U = sum_p dh, V = sum_p dh xhat:
//   d shift = U,  d scale = gamma V + beta U,  d beta = sum_b (scale + 1) U,  d gamma = sum_b (scale + 1) V,
//   dy = r * ( dh (scale + 1) gamma  -  m1  -  xhat m2 ),  m1 = mean_g[(scale + 1) gamma U] , m2 = mean_g[(scale + 1) gamma V]
// (means over the group's channels and all pixels).  Three streaming passes (statistics, (U, V), dy); every reduction is
// two-stage in a fixed order -> bit-reproducible.
#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

constexpr int G = 8;   // groups

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    float2 t;
    t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
    t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
    t = ptx::unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
    t = ptx::unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 o;
    o.x = ptx::pack_bf16x2(v[0], v[1]); o.y = ptx::pack_bf16x2(v[2], v[3]);
    o.z = ptx::pack_bf16x2(v[4], v[5]); o.w = ptx::pack_bf16x2(v[6], v[7]);
    return o;
}
__device__ __forceinline__ float block_sum(float v, float* s_red) {   // 256 threads; result valid in every thread
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_red[k];
    return t;
}

// ---------------------------------------------------------------------------------------------- statistics
// One coalesced pass: grid (nchunk, B), thread = (8-channel chunk cc, pixel lane pl) like the passes below; every thread's eight
// channels lie in ONE group, so it keeps a (sum, sum of squares) pair; per-group partials per pixel chunk are merged in double
// precision by gn_stats_finish_kernel (E[x^2] - mean^2 on conv outputs: |mean| ~ std, far from cancellation in fp64).
__global__ void __launch_bounds__(256)
gn_stats_part_kernel(const uint4* __restrict__ y, int P, int C, int nchunk, float2* __restrict__ spart) {
    __shared__ float s_s[256], s_q[256];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cpp = C / 8, lanes = 256 / cpp;
    const int cc = threadIdx.x % cpp, pl = threadIdx.x / cpp;
    const int ppc = P / nchunk;
    const size_t base = (static_cast<size_t>(b) * P + static_cast<size_t>(chunk) * ppc) * cpp + cc;
    float s1 = 0.f, s2 = 0.f;
    for (int p = pl; p < ppc; p += lanes) {
        float v[8];
        unpack8(__ldg(y + base + static_cast<size_t>(p) * cpp), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
    }
    s_s[threadIdx.x] = s1;
    s_q[threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < G) {
        const int g = threadIdx.x, gch = cpp / G;        // chunks cc in [g * gch, (g + 1) * gch) belong to group g
        float a = 0.f, q = 0.f;
        for (int l = 0; l < lanes; ++l)
            for (int k = 0; k < gch; ++k) { a += s_s[l * cpp + g * gch + k]; q += s_q[l * cpp + g * gch + k]; }
        spart[(static_cast<size_t>(b) * nchunk + chunk) * G + g] = make_float2(a, q);
    }
}
__global__ void gn_stats_finish_kernel(const float2* __restrict__ spart, int nchunk, int P, int C, float eps, float2* __restrict__ stats) {
    const int b = blockIdx.x, g = threadIdx.x;
    if (g >= G) return;
    double a = 0.0, q = 0.0;
    for (int k = 0; k < nchunk; ++k) { const float2 v = spart[(static_cast<size_t>(b) * nchunk + k) * G + g]; a += v.x; q += v.y; }
    const double n = static_cast<double>(P) * (C / G);
    const double mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[b * G + g] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
}

__device__ __forceinline__ float silu_grad(float h) {     // sigmoid through one tanh.approx (see train_kernels.cu): these passes are issue-bound
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * h));
    const float sg = fmaf(0.5f, t, 0.5f);
    return sg * fmaf(h, 1.0f - sg, 1.0f);
}

constexpr int GB_CHUNKS_MAX = 16;
constexpr int GB_PIX = 4;          // pixels in flight per thread in the two streaming passes

// ---------------------------------------------------------------------------------------------- pass 1: (U, V) partials
// grid (nchunk, B); thread = (8-channel chunk cc, pixel lane pl), lanes = 256 / (C / 8)
__global__ void __launch_bounds__(256)
gn_bwd_sums_kernel(const uint4* __restrict__ y, const uint4* __restrict__ ds, const float2* __restrict__ stats,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                   const float* __restrict__ shift, int ld, int P, int C, int nchunk, float* __restrict__ part) {
    extern __shared__ float gs_red[];   // [lanes][4 * C]: U, V, the plain sum of ds (SR3 post-activation embedding's gradient), sum of y
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cpp = C / 8, lanes = 256 / cpp;
    const int cc = threadIdx.x % cpp, pl = threadIdx.x / cpp;
    const int ppc = P / nchunk;
    const float2 st = stats[b * G + (cc * 8) / (C / G)];
    float k1[8], k0[8], U[8], V[8], Ws[8], Ys[8];      // h = xhat * k1 + k0, k1 = gamma (scale + 1), k0 = beta (scale + 1) + shift
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cc * 8 + j;
        const float sc1 = scale ? scale[static_cast<size_t>(b) * ld + c] + 1.0f : 1.0f;
        const float sh = shift ? shift[static_cast<size_t>(b) * ld + c] : 0.0f;
        k1[j] = gamma[c] * sc1;
        k0[j] = fmaf(beta[c], sc1, sh);
        U[j] = 0.f; V[j] = 0.f; Ws[j] = 0.f; Ys[j] = 0.f;
    }
    const size_t base = (static_cast<size_t>(b) * P + static_cast<size_t>(chunk) * ppc) * cpp + cc;
    // GB_PIX pixels' loads are issued before any of them is consumed (the pass is latency-bound at ~80 registers / 3 CTAs per SM);
    // a pixel past the chunk reads as zeros and adds nothing, so the per-thread summation order is unchanged
    for (int p0 = pl; p0 < ppc; p0 += lanes * GB_PIX) {
        uint4 ry[GB_PIX], rg[GB_PIX];
#pragma unroll
        for (int k = 0; k < GB_PIX; ++k) {
            const int p = p0 + k * lanes;
            ry[k] = rg[k] = make_uint4(0u, 0u, 0u, 0u);
            if (p < ppc) {
                ry[k] = __ldg(y + base + static_cast<size_t>(p) * cpp);
                rg[k] = __ldg(ds + base + static_cast<size_t>(p) * cpp);
            }
        }
#pragma unroll
        for (int k = 0; k < GB_PIX; ++k) {
            float yv[8], g[8];
            unpack8(ry[k], yv);
            unpack8(rg[k], g);
            if (p0 + k * lanes < ppc) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (yv[j] - st.x) * st.y;
                    const float dh = g[j] * silu_grad(fmaf(xh, k1[j], k0[j]));
                    U[j] += dh;
                    V[j] = fmaf(dh, xh, V[j]);
                    Ws[j] += g[j];
                    Ys[j] += yv[j];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float* r = gs_red + pl * 4 * C + cc * 8 + j;
        r[0] = U[j]; r[C] = V[j]; r[2 * C] = Ws[j]; r[3 * C] = Ys[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * C; i += 256) {
        float t = 0.f;
        for (int k = 0; k < lanes; ++k) t += gs_red[k * 4 * C + i];
        part[(static_cast<size_t>(b) * nchunk + chunk) * 4 * C + i] = t;
    }
}

// ---------------------------------------------------------------------------------------------- per-sample coefficients
// grid B, 2 * 256 threads max -> one thread per channel (C <= 512): UV[b][2][C], group means m[b][G][2], d scale / d shift
__global__ void __launch_bounds__(512)
gn_bwd_coef_kernel(const float* __restrict__ part, int nchunk, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ scale, int ld, int P, int C, float* __restrict__ UV, float* __restrict__ gm,
                   float* __restrict__ dscale, float* __restrict__ dshift, float* __restrict__ dpost) {
    __shared__ float s_a[512], s_b[512];
    const int b = blockIdx.x, c = threadIdx.x;
    float U = 0.f, V = 0.f, Wp = 0.f, Yp = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int k = 0; k < nchunk; ++k) {
            const float* pp = part + (static_cast<size_t>(b) * nchunk + k) * 4 * C + c;
            U += pp[0]; V += pp[C]; Wp += pp[2 * C]; Yp += pp[3 * C];
        }
        if (dpost) dpost[static_cast<size_t>(b) * ld + c] = Wp;
        UV[(static_cast<size_t>(b) * 3) * C + c] = U;
        UV[(static_cast<size_t>(b) * 3 + 1) * C + c] = V;
        UV[(static_cast<size_t>(b) * 3 + 2) * C + c] = Yp;
        if (dshift) dshift[static_cast<size_t>(b) * ld + c] = U;
        if (dscale) dscale[static_cast<size_t>(b) * ld + c] = fmaf(gamma[c], V, beta[c] * U);
        const float k1 = (scale ? scale[static_cast<size_t>(b) * ld + c] + 1.0f : 1.0f) * gamma[c];
        s_a[c] = k1 * U;
        s_b[c] = k1 * V;
    }
    __syncthreads();
    if (c < G) {
        const int cpg = C / G;
        float a = 0.f, v = 0.f;
        for (int j = 0; j < cpg; ++j) { a += s_a[c * cpg + j]; v += s_b[c * cpg + j]; }
        const float inv = 1.0f / (static_cast<float>(cpg) * static_cast<float>(P));
        gm[(b * G + c) * 2] = a * inv;
        gm[(b * G + c) * 2 + 1] = v * inv;
    }
}

// d gamma[c] = sum_b (scale + 1) V, d beta[c] = sum_b (scale + 1) U; optionally the bias gradient of the conv that produced y:
// sum_{b,p} dy = sum_b r (k1 U - P m1 - m2 sum_p xhat), sum_p xhat = r (sum_p y - P mu)  -- no extra pass over dy
// grid C / 32, 256 threads = 32 channels x 8 sample slices (a single thread per channel walking all B samples was a chain of
// dependent loads: 42 us per launch, 9 % of the Unet training step); the slices are combined in a fixed order.
__global__ void __launch_bounds__(256)
gn_bwd_affine_kernel(const float* __restrict__ UV, const float* __restrict__ scale, int ld, int B, int C, int P,
                     const float* __restrict__ gamma, const float2* __restrict__ stats, const float* __restrict__ gm,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dconv_bias) {
    __shared__ float s_r[3][8][32];
    const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int g = c / (C / G);
    const int per = (B + 7) / 8;
    const int b0 = sl * per, b1 = min(B, b0 + per);
    float dg = 0.f, db = 0.f, dcb = 0.f;
    for (int b = b0; b < b1; ++b) {
        const float s1 = scale ? scale[static_cast<size_t>(b) * ld + c] + 1.0f : 1.0f;
        const float U = UV[(static_cast<size_t>(b) * 3) * C + c], V = UV[(static_cast<size_t>(b) * 3 + 1) * C + c];
        db = fmaf(s1, U, db);
        dg = fmaf(s1, V, dg);
        if (dconv_bias != nullptr) {
            const float2 st = stats[b * G + g];
            const float m1 = gm[(b * G + g) * 2], m2 = gm[(b * G + g) * 2 + 1];
            const float xs = st.y * (UV[(static_cast<size_t>(b) * 3 + 2) * C + c] - static_cast<float>(P) * st.x);
            dcb += st.y * (gamma[c] * s1 * U - static_cast<float>(P) * m1 - m2 * xs);
        }
    }
    s_r[0][sl][cl] = dg; s_r[1][sl][cl] = db; s_r[2][sl][cl] = dcb;
    __syncthreads();
    if (sl == 0) {
        float a = 0.f, bsum = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += s_r[0][k][cl]; bsum += s_r[1][k][cl]; d += s_r[2][k][cl]; }
        dgamma[c] = a;
        dbeta[c] = bsum;
        if (dconv_bias != nullptr) dconv_bias[c] = d;
    }
}

// ---------------------------------------------------------------------------------------------- pass 2: dy
__global__ void __launch_bounds__(256)
gn_bwd_dx_kernel(const uint4* __restrict__ y, const uint4* ds, const float2* __restrict__ stats, const float* __restrict__ gm,
                 const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                 const float* __restrict__ shift, int ld, int P, int C, int nchunk, uint4* dy) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cpp = C / 8, lanes = 256 / cpp;
    const int cc = threadIdx.x % cpp, pl = threadIdx.x / cpp;
    const int ppc = P / nchunk;
    const int g = (cc * 8) / (C / G);
    const float2 st = stats[b * G + g];
    const float m1 = gm[(b * G + g) * 2], m2 = gm[(b * G + g) * 2 + 1];
    float k1[8], k0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cc * 8 + j;
        const float sc1 = scale ? scale[static_cast<size_t>(b) * ld + c] + 1.0f : 1.0f;
        const float sh = shift ? shift[static_cast<size_t>(b) * ld + c] : 0.0f;
        k1[j] = gamma[c] * sc1;
        k0[j] = fmaf(beta[c], sc1, sh);
    }
    const size_t base = (static_cast<size_t>(b) * P + static_cast<size_t>(chunk) * ppc) * cpp + cc;
    for (int p0 = pl; p0 < ppc; p0 += lanes * GB_PIX) {
        uint4 ry[GB_PIX], rg[GB_PIX];
#pragma unroll
        for (int k = 0; k < GB_PIX; ++k) {          // all loads first (dy may alias ds: each pixel is read before it is written)
            const int p = p0 + k * lanes;
            if (p < ppc) {
                ry[k] = __ldg(y + base + static_cast<size_t>(p) * cpp);
                rg[k] = ds[base + static_cast<size_t>(p) * cpp];
            }
        }
#pragma unroll
        for (int k = 0; k < GB_PIX; ++k) {
            const int p = p0 + k * lanes;
            if (p >= ppc) break;
            float yv[8], gr[8];
            unpack8(ry[k], yv);
            unpack8(rg[k], gr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (yv[j] - st.x) * st.y;
                const float dh = gr[j] * silu_grad(fmaf(xh, k1[j], k0[j]));
                gr[j] = st.y * (dh * k1[j] - m1 - xh * m2);
            }
            dy[base + static_cast<size_t>(p) * cpp] = pack8(gr);
        }
    }
}

// ---------------------------------------------------------------------------------------------- channel LayerNorm backward
// z = xhat * g, xhat = (x - mu) * r over the C channels of a pixel.  One warp per pixel, PB pixels in flight per warp (their
// shuffle reductions interleave):
//   dx = r * (dz g - mean_c(dz g) - xhat mean_c(dz g xhat));  dg partial sums per block -> part[blk][C]
template <int PER, int PB>      // PER = C / 32 channels per lane
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dz, const float* __restrict__ gain, long long M, float eps,
              bf16* __restrict__ dx, float* __restrict__ part) {
    constexpr int C = PER * 32;
    extern __shared__ float ln_dg[];   // [8 warps][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float dg[PER], gn[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) { dg[j] = 0.f; gn[j] = gain[j * 32 + lane]; }
    const long long stride = static_cast<long long>(gridDim.x) * 8 * PB;
    for (long long m0 = (static_cast<long long>(blockIdx.x) * 8 + warp) * PB; m0 < M; m0 += stride) {
        float xv[PB][PER], gv[PB][PER], s[PB], vs[PB], a[PB], bs[PB], r[PB];
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            const long long m = m0 + q < M ? m0 + q : M - 1;          // tail: recompute the last pixel, store guarded below
            s[q] = 0.f;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                xv[q][j] = __bfloat162float(x[m * C + j * 32 + lane]);
                gv[q][j] = __bfloat162float(dz[m * C + j * 32 + lane]);
                s[q] += xv[q][j];
            }
        }
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < PB; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            s[q] /= C;
            vs[q] = 0.f;
#pragma unroll
            for (int j = 0; j < PER; ++j) { const float d = xv[q][j] - s[q]; vs[q] = fmaf(d, d, vs[q]); }
        }
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < PB; ++q) vs[q] += __shfl_xor_sync(0xffffffffu, vs[q], o);
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            r[q] = rsqrtf(vs[q] / C + eps);
            a[q] = 0.f; bs[q] = 0.f;
            const bool live = m0 + q < M;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const float xh = (xv[q][j] - s[q]) * r[q];
                const float t = gv[q][j] * gn[j];
                if (live) dg[j] = fmaf(gv[q][j], xh, dg[j]);
                a[q] += t;
                bs[q] = fmaf(t, xh, bs[q]);
                xv[q][j] = xh;
                gv[q][j] = t;
            }
        }
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < PB; ++q) { a[q] += __shfl_xor_sync(0xffffffffu, a[q], o); bs[q] += __shfl_xor_sync(0xffffffffu, bs[q], o); }
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            if (m0 + q < M) {
                const float am = a[q] / C, bm = bs[q] / C;
#pragma unroll
                for (int j = 0; j < PER; ++j) dx[(m0 + q) * C + j * 32 + lane] = __float2bfloat16(r[q] * (gv[q][j] - am - xv[q][j] * bm));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) ln_dg[warp * C + j * 32 + lane] = dg[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += ln_dg[k * C + c];
        part[static_cast<size_t>(blockIdx.x) * C + c] = t;
    }
}

// ---------------------------------------------------------------------------------------------- weight standardisation backward
// grid Cout: dw = rstd * (dwt - mean(dwt) - what * mean(dwt * what)), what = (w - mean) * rstd over K = Cin * k * k
__global__ void __launch_bounds__(256)
ws_bwd_kernel(const float* __restrict__ w, const float* dwt, int K, float eps, float* dw) {
    __shared__ float s_red[8];
    const float* wr = w + static_cast<size_t>(blockIdx.x) * K;
    const float* gr = dwt + static_cast<size_t>(blockIdx.x) * K;
    float* orow = dw + static_cast<size_t>(blockIdx.x) * K;
    float s = 0.f;
    for (int i = threadIdx.x; i < K; i += 256) s += wr[i];
    const float mean = block_sum(s, s_red) / K;
    float v = 0.f;
    for (int i = threadIdx.x; i < K; i += 256) { const float d = wr[i] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(block_sum(v, s_red) / K + eps);
    float a = 0.f, bsum = 0.f;
    for (int i = threadIdx.x; i < K; i += 256) {
        const float g = gr[i];
        a += g;
        bsum = fmaf(g, (wr[i] - mean) * rstd, bsum);
    }
    a = block_sum(a, s_red) / K;
    bsum = block_sum(bsum, s_red) / K;
    for (int i = threadIdx.x; i < K; i += 256) orow[i] = rstd * (gr[i] - a - (wr[i] - mean) * rstd * bsum);
}

// out[i] = sum_k part[k][i]: 32 columns per block, 32 threads per column over contiguous slices (eight loads in flight each), fixed combination order
constexpr int SUM_SLICES = 32;      // threads per column (each sums a contiguous slice of the partials)
__global__ void __launch_bounds__(32 * SUM_SLICES)
sum_rows_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ out) {
    __shared__ float s_p[SUM_SLICES][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;
    const int per = (nparts + SUM_SLICES - 1) / SUM_SLICES;
    const int k0 = sl * per, k1 = min(nparts, k0 + per);
    float t = 0.f;
    if (col < n) {
        int k = k0;
        for (; k + 8 <= k1; k += 8) {              // eight independent loads in flight, added in order
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(part + static_cast<size_t>(k + j) * n + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) t += v[j];
        }
        for (; k < k1; ++k) t += __ldg(part + static_cast<size_t>(k) * n + col);
    }
    s_p[sl][threadIdx.x & 31] = t;
    __syncthreads();
    if (sl == 0 && col < n) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < SUM_SLICES; ++k) v += s_p[k][threadIdx.x];
        out[col] = v;
    }
}

int gn_chunks(int P, int C) {
    const int lanes = 256 / (C / 8);
    int nchunk = GB_CHUNKS_MAX;
    while (nchunk > 1 && (P % nchunk != 0 || (P / nchunk) < lanes)) nchunk >>= 1;
    return nchunk;
}

}  // namespace

size_t gn_bwd_scratch_floats(int B, int P, int C) {
    return static_cast<size_t>(B) * G * 2 /*stats*/ + static_cast<size_t>(B) * gn_chunks(P, C) * 4 * C /*part*/ +
           static_cast<size_t>(B) * 3 * C /*U, V, sum y*/ + static_cast<size_t>(B) * GB_CHUNKS_MAX * G * 2 /*gm, statistics partials*/ + 64;
}

cudaError_t groupnorm_silu_bwd_run(const GroupNormBwdArgs& a, float* scratch, cudaStream_t s) {
    const int B = a.B, P = a.P, C = a.C;
    if (C % 64 != 0 || C > 512 || 256 % (C / 8) != 0) return cudaErrorInvalidValue;
    const int nchunk = gn_chunks(P, C);
    const int lanes = 256 / (C / 8);
    const int ld = a.ld > 0 ? a.ld : C;
    float2* stats_buf = reinterpret_cast<float2*>(scratch);
    const float2* stats = a.stats_in != nullptr ? a.stats_in : stats_buf;
    float* part = scratch + static_cast<size_t>(B) * G * 2;
    float* UV = part + static_cast<size_t>(B) * nchunk * 4 * C;
    float* gm = UV + static_cast<size_t>(B) * 3 * C;
    const uint4* y = reinterpret_cast<const uint4*>(a.y);
    const uint4* ds = reinterpret_cast<const uint4*>(a.ds);
    if (a.stats_in == nullptr) {
        gn_stats_part_kernel<<<dim3(nchunk, B), 256, 0, s>>>(y, P, C, nchunk, reinterpret_cast<float2*>(gm));   // gm is free until the coefficient pass
        gn_stats_finish_kernel<<<B, 32, 0, s>>>(reinterpret_cast<const float2*>(gm), nchunk, P, C, a.eps, stats_buf);
    }
    gn_bwd_sums_kernel<<<dim3(nchunk, B), 256, static_cast<size_t>(lanes) * 4 * C * sizeof(float), s>>>(
        y, ds, stats, a.gamma, a.beta, a.scale, a.shift, ld, P, C, nchunk, part);
    gn_bwd_coef_kernel<<<B, 512, 0, s>>>(part, nchunk, a.gamma, a.beta, a.scale, ld, P, C, UV, gm, a.dscale, a.dshift, a.dpost);
    gn_bwd_affine_kernel<<<C / 32, 256, 0, s>>>(UV, a.scale, ld, B, C, P, a.gamma, stats, gm, a.dgamma, a.dbeta, a.dconv_bias);
    gn_bwd_dx_kernel<<<dim3(nchunk, B), 256, 0, s>>>(y, ds, stats, gm, a.gamma, a.beta, a.scale, a.shift, ld, P, C, nchunk,
                                                      reinterpret_cast<uint4*>(a.dy));
    return cudaGetLastError();
}

int ln_bwd_blocks(long long M) { return static_cast<int>(M / 8 < 592 ? (M + 7) / 8 : 592); }

cudaError_t channel_layernorm_bwd_run(const bf16* x, const bf16* dz, const float* gain, long long M, int C, float eps, bf16* dx,
                                      float* dgain, float* part, cudaStream_t s) {
    const int blocks = ln_bwd_blocks(M);
    const size_t smem = static_cast<size_t>(8) * C * sizeof(float);
    switch (C) {
        case 64: ln_bwd_kernel<2, 4><<<blocks, 256, smem, s>>>(x, dz, gain, M, eps, dx, part); break;
        case 128: ln_bwd_kernel<4, 4><<<blocks, 256, smem, s>>>(x, dz, gain, M, eps, dx, part); break;
        case 256: ln_bwd_kernel<8, 2><<<blocks, 256, smem, s>>>(x, dz, gain, M, eps, dx, part); break;
        case 512: ln_bwd_kernel<16, 1><<<blocks, 256, smem, s>>>(x, dz, gain, M, eps, dx, part); break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_rows_kernel<<<(C + 31) / 32, 32 * SUM_SLICES, 0, s>>>(part, blocks, C, dgain);
    return cudaGetLastError();
}

cudaError_t weight_standardize_bwd_run(const float* w, const float* dwt, int Cout, int K, float eps, float* dw, cudaStream_t s) {
    ws_bwd_kernel<<<Cout, 256, 0, s>>>(w, dwt, K, eps, dw);
    return cudaGetLastError();
}

}  // namespace hd
