// Bandwidth-bound kernels of the HiCEDRN training step (train.py:109-136 trains hicedrn_Diff through
// GaussianDiffusion.p_losses, hicdiff_condition.py:715-746).  The convolutions run on conv_gemm.cu (forward and dgrad) and
// wgrad.cu; everything around them is here:
//
//   prep_dgrad_weight   [Cout, Cin, 3, 3] fp32 -> [Cin, (tap', cout)] bf16 with tap' = 8 - tap: the data gradient of a 3x3
//                       "same" conv is the same implicit GEMM over the flipped, transposed filter
//   film_silu_fwd       s = SiLU(a * (scale + 1) + shift)                  hicedrn_Diff.py:175-178,201-203
//   film_silu_bwd       da = ds * SiLU'(h) * (scale + 1), per-(sample, channel) sums for d scale / d shift
//   colsum              per-channel sums over all pixels (bias gradients)
//   thin_wgrad          weight gradient of the 1- or 2-plane head conv / the 1-channel tail conv
//   loss_grad           weighted l1 / l2 loss and its gradient             hicdiff_condition.py:706-713,741-746
//   linear_bwd_*        the time-embedding MLPs (a few rows; fp32)         hicedrn_Diff.py:232-246,189-200
// All reductions are two-pass with a fixed order (no atomics): gradients are bit-reproducible run to run.
#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

// sigmoid(v) = 0.5 * tanh(0.5 v) + 0.5 with one MUFU op (tanh.approx, abs. error ~2^-11: below the bf16 rounding of every
// tensor these kernels write); the exact-division form made the FiLM kernels issue-bound instead of HBM-bound
__device__ __forceinline__ float sigmoid_f(float v) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
    return fmaf(0.5f, t, 0.5f);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    float2 t;
    t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
    t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
    t = ptx::unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
    t = ptx::unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 o;
    o.x = ptx::pack_bf16x2(v[0], v[1]);
    o.y = ptx::pack_bf16x2(v[2], v[3]);
    o.z = ptx::pack_bf16x2(v[4], v[5]);
    o.w = ptx::pack_bf16x2(v[6], v[7]);
    return o;
}

// ------------------------------------------------------------------------------------------------ weights
__global__ void __launch_bounds__(256)
prep_dgrad_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin) {
    const int ci = blockIdx.x;
    bf16* orow = out + static_cast<size_t>(ci) * 9 * Cout;
    for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) {
        const int tapp = i / Cout, co = i - tapp * Cout;
        orow[i] = __float2bfloat16(w[(static_cast<size_t>(co) * Cin + ci) * 9 + (8 - tapp)]);
    }
}

// tail conv [1, C, 3, 3] -> the [C, 1, 3, 3] filter whose forward conv over d_eps is the tail's data gradient
__global__ void flip_tail_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C * 9) {
        const int c = i / 9, t = i - c * 9;
        out[i] = w[c * 9 + (8 - t)];
    }
}

// ------------------------------------------------------------------------------------------------ FiLM + SiLU
constexpr int FB_CHUNKS = 16;   // pixel chunks per image (P = 4096 -> 256 pixels per CTA)

// grid (FB_CHUNKS, B), C == 256: thread = (8-channel chunk cc = tid % 32, pixel lane pl = tid / 32); the thread's scale / shift
// stay in registers and loads are issued four pixels ahead of their use
__global__ void __launch_bounds__(256)
film_silu_fwd_kernel(const uint4* __restrict__ a, uint4* __restrict__ s, const float* __restrict__ film, int ld, int off, int P,
                     int has_scale) {
    constexpr int C = 256, CPP = C / 8;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cc = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int ppc = P / FB_CHUNKS;
    const float* row = film + static_cast<size_t>(b) * ld + off + cc * 8;
    float sc1[8], sh[8];   // has_scale == 0 (SR3 FeatureWiseAffine): the row holds the additive term only
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc1[j] = has_scale ? __ldg(row + j) + 1.0f : 1.0f; sh[j] = __ldg(row + (has_scale ? C : 0) + j); }
    const size_t base = (static_cast<size_t>(b) * P + static_cast<size_t>(chunk) * ppc) * CPP + cc;
    for (int p = pl; p < ppc; p += 32) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(a + base + static_cast<size_t>(p + 8 * k) * CPP);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v[8];
            unpack8(u[k], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float h = fmaf(v[j], sc1[j], sh[j]);
                v[j] = h * sigmoid_f(h);
            }
            s[base + static_cast<size_t>(p + 8 * k) * CPP] = pack8(v);
        }
    }
}

// same thread layout; ds and da may be the same buffer (each element is read before it is written by the same thread)
__global__ void __launch_bounds__(256)
film_silu_bwd_kernel(const uint4* ds, const uint4* __restrict__ a, uint4* da, const float* __restrict__ film, int ld, int off, int P,
                     float* __restrict__ part, int has_scale) {
    constexpr int C = 256, CPP = C / 8;
    __shared__ float s_red[8][2 * C];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cc = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int ppc = P / FB_CHUNKS;
    const float* row = film + static_cast<size_t>(b) * ld + off + cc * 8;
    float sc1[8], sh[8], acc_sc[8], acc_sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc1[j] = has_scale ? __ldg(row + j) + 1.0f : 1.0f;
        sh[j] = __ldg(row + (has_scale ? C : 0) + j);
        acc_sc[j] = 0.f; acc_sh[j] = 0.f;
    }
    const size_t base = (static_cast<size_t>(b) * P + static_cast<size_t>(chunk) * ppc) * CPP + cc;
    for (int p = pl; p < ppc; p += 32) {
        uint4 ug[4], ua[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t i = base + static_cast<size_t>(p + 8 * k) * CPP;
            ug[k] = ds[i];
            ua[k] = __ldg(a + i);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float g[8], av[8];
            unpack8(ug[k], g);
            unpack8(ua[k], av);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float h = fmaf(av[j], sc1[j], sh[j]);
                const float sg = sigmoid_f(h);
                const float dh = g[j] * (sg * (1.0f + h * (1.0f - sg)));
                acc_sc[j] = fmaf(dh, av[j], acc_sc[j]);
                acc_sh[j] += dh;
                g[j] = dh * sc1[j];
            }
            da[base + static_cast<size_t>(p + 8 * k) * CPP] = pack8(g);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[pl][cc * 8 + j] = acc_sc[j]; s_red[pl][C + cc * 8 + j] = acc_sh[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_red[k][i];
        part[(static_cast<size_t>(b) * FB_CHUNKS + chunk) * 2 * C + i] = t;
    }
}

// dfilm[b, off + i] = sum over the chunks (i < 2C: d scale then d shift; has_scale == 0: d shift only, i < C); grid B, 2C threads
__global__ void film_grad_finish_kernel(const float* __restrict__ part, float* __restrict__ dfilm, int ld, int off, int C, int has_scale) {
    const int b = blockIdx.x, i = threadIdx.x;
    if (!has_scale && i < C) return;
    float t = 0.f;
    for (int k = 0; k < FB_CHUNKS; ++k) t += part[(static_cast<size_t>(b) * FB_CHUNKS + k) * 2 * C + i];
    dfilm[static_cast<size_t>(b) * ld + off + (has_scale ? i : i - C)] = t;
}

// the shared conv's bias gradient: its first use contributes sum_b (scale + 1) * dshift (da = dh * (scale + 1)), its second
// 0.1 * colsum(g)
__global__ void __launch_bounds__(256)
edrn_bias_grad_kernel(const float* __restrict__ film, const float* __restrict__ dfilm, int ld, int off, int B,
                      const float* __restrict__ colsum_g, float g_scale, float* __restrict__ dbias, int C, int has_scale) {
    // 32 channels x 8 sample slices per block (a single thread per channel over all B samples is a chain of dependent loads)
    __shared__ float s_p[8][32];
    const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int per = (B + 7) / 8;
    const int b0 = sl * per, b1 = min(B, b0 + per);
    float t = 0.f;
    if (c < C) {
        for (int b = b0; b < b1; ++b) {
            if (has_scale) t = fmaf(film[static_cast<size_t>(b) * ld + off + c] + 1.0f, dfilm[static_cast<size_t>(b) * ld + off + C + c], t);
            else t += dfilm[static_cast<size_t>(b) * ld + off + c];
        }
    }
    s_p[sl][cl] = t;
    __syncthreads();
    if (sl == 0 && c < C) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += s_p[k][cl];
        dbias[c] = v + g_scale * colsum_g[c];
    }
}

// ------------------------------------------------------------------------------------------------ column sums
// x [M, C] bf16, C/8 divides 256.  Block k sums rows [k * rpb, (k + 1) * rpb) -> part[k][C]
__global__ void __launch_bounds__(256)
colsum_part_kernel(const uint4* __restrict__ x, float* __restrict__ part, long long M, int C, int rpb) {
    extern __shared__ float cs_red[];   // [lanes][C]
    const int cpp = C / 8;
    const int lanes = 256 / cpp;
    const int cc = threadIdx.x % cpp, pl = threadIdx.x / cpp;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const long long r0 = static_cast<long long>(blockIdx.x) * rpb;
    const long long r1 = r0 + rpb < M ? r0 + rpb : M;
    long long r = r0 + pl;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {     // four rows in flight
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(x + (r + k * lanes) * cpp + cc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v[8];
            unpack8(u[k], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
    }
    for (; r < r1; r += lanes) {
        float v[8];
        unpack8(__ldg(x + r * cpp + cc), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) cs_red[pl * C + cc * 8 + j] = acc[j];
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) {
        float t = 0.f;
        for (int k = 0; k < lanes; ++k) t += cs_red[k * C + i];
        part[static_cast<size_t>(blockIdx.x) * C + i] = t;
    }
}

// out[i] = scale * sum_k part[k][i] (+ out[i] when accumulate); n columns.  32 columns per block, 32 threads per column each
// summing a contiguous slice of the partials, combined in a fixed order (a single thread per column was latency-bound).
constexpr int SUM_SLICES = 32;      // threads per column (each sums a contiguous slice of the partials)
__global__ void __launch_bounds__(32 * SUM_SLICES)
sum_parts_kernel(const float* __restrict__ part, int nparts, int n, float scale, int accumulate, float* __restrict__ out) {
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
        out[col] = accumulate ? fmaf(scale, v, out[col]) : scale * v;
    }
}

// y = a + b (bf16, 16-byte chunks)
__global__ void __launch_bounds__(256)
add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y, long long chunks) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < chunks; i += gridDim.x * 256ll) {
        float x[8], z[8];
        unpack8(__ldg(a + i), x);
        unpack8(__ldg(b + i), z);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += z[j];
        y[i] = pack8(x);
    }
}

// ------------------------------------------------------------------------------------------------ thin wgrad
// part[(b * 8 + rg)][k][tap][c] = sum over the 8 rows of row group rg of G[b, y, x, c] * u_k[b, y + sgn*(ky-1), x + sgn*(kx-1)]
// G [B, 64, 64, 256] bf16, u_k fp32 planes [B, 64, 64]; grid (8, B), 256 threads = channels.
__global__ void __launch_bounds__(256)
thin_wgrad_kernel(const bf16* __restrict__ G, const float* __restrict__ u0, const float* __restrict__ u1, int sgn,
                  float* __restrict__ part) {
    constexpr int C = 256, W = 64, H = 64, ROWS = 8, PITCH = 66;
    __shared__ float s_u[2][ROWS + 2][PITCH];
    const int rg = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    const int nk = u1 != nullptr ? 2 : 1;
    const int y0 = rg * ROWS;
    for (int i = threadIdx.x; i < nk * (ROWS + 2) * PITCH; i += 256) {
        const int k = i / ((ROWS + 2) * PITCH);
        const int rem = i - k * (ROWS + 2) * PITCH;
        const int yy = rem / PITCH + y0 - 1, xx = rem % PITCH - 1;
        const float* u = k == 0 ? u0 : u1;
        s_u[k][rem / PITCH][rem % PITCH] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(u + (static_cast<size_t>(b) * H + yy) * W + xx) : 0.f;
    }
    __syncthreads();
    float acc[2][9];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
    const bf16* g = G + ((static_cast<size_t>(b) * H + y0) * W) * C + c;
    // eight pixels per step: their gradients are loaded together and the 3 x 10 window of u they touch is read from shared
    // memory once (the per-(pixel, tap) broadcast reads made the loop LDS-issue bound: one LDS per FMA)
    for (int yl = 0; yl < ROWS; ++yl) {
        for (int x0 = 0; x0 < W; x0 += 8) {
            float gv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) gv[i] = __bfloat162float(g[(static_cast<size_t>(yl) * W + x0 + i) * C]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < nk) {
                    float uw[3][10];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int j = 0; j < 10; ++j) uw[r][j] = s_u[k][yl + r][x0 + j];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const int dy = t / 3 - 1, dx = t % 3 - 1;
                            const float uv = sgn > 0 ? uw[1 + dy][i + 1 + dx] : uw[1 - dy][i + 1 - dx];
                            acc[k][t] = fmaf(gv[i], uv, acc[k][t]);
                        }
                }
            }
        }
    }
    float* o = part + (static_cast<size_t>(b) * 8 + rg) * nk * 9 * C;
    for (int k = 0; k < nk; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) o[(k * 9 + t) * C + c] = acc[k][t];
}

// dW[(c * nk + k) * 9 + tap] = sum_parts part[.][k][tap][c]: 32 columns per block, 32 threads per column over contiguous slices
// of the partials (eight loads in flight each), combined in a fixed order
__global__ void __launch_bounds__(1024)
thin_wgrad_finish_kernel(const float* __restrict__ part, int nparts, int nk, float* __restrict__ dw) {
    constexpr int C = 256;
    __shared__ float s_p[32][32];
    const int n = nk * 9 * C;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;    // i indexes [k][tap][c]
    const int per = (nparts + 31) / 32;
    const int k0 = sl * per, k1 = min(nparts, k0 + per);
    float t = 0.f;
    if (i < n) {
        int k = k0;
        for (; k + 8 <= k1; k += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(part + static_cast<size_t>(k + j) * n + i);
#pragma unroll
            for (int j = 0; j < 8; ++j) t += v[j];
        }
        for (; k < k1; ++k) t += __ldg(part + static_cast<size_t>(k) * n + i);
    }
    s_p[sl][threadIdx.x & 31] = t;
    __syncthreads();
    if (sl == 0 && i < n) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) v += s_p[k][threadIdx.x];
        const int c = i % C, kt = i / C, kk = kt / 9, tap = kt - kk * 9;
        dw[(c * nk + kk) * 9 + tap] = v;
    }
}

// ------------------------------------------------------------------------------------------------ loss
// loss_type 0: l1, 1: l2.  part[blk] = sum over the block's elements of |d|^p * w[b]; d_eps = dL/d eps for L = mean(...)
__global__ void __launch_bounds__(256)
loss_grad_kernel(const float* __restrict__ eps, const float* __restrict__ target, const float* __restrict__ w, int loss_type,
                 long long n, int tile_elems, float* __restrict__ d_eps, float* __restrict__ part) {
    __shared__ float s_red[8];
    const float inv_n = 1.0f / static_cast<float>(n);
    float acc = 0.f;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const float wb = __ldg(w + i / tile_elems);
        const float d = eps[i] - target[i];
        if (loss_type == 1) {
            acc = fmaf(d * d, wb, acc);
            d_eps[i] = 2.0f * d * wb * inv_n;
        } else {
            acc = fmaf(fabsf(d), wb, acc);
            d_eps[i] = (d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f)) * wb * inv_n;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += s_red[k];
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
sum_f32_part_kernel(const float* __restrict__ x, long long n, float* __restrict__ part) {
    __shared__ float s_red[8];
    float acc = 0.f;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) acc += x[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += s_red[k];
        part[blockIdx.x] = t;
    }
}

__global__ void sum_scalar_kernel(const float* __restrict__ part, int nparts, float scale, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < nparts; ++k) t += part[k];
        out[0] = static_cast<float>(t * scale);
    }
}

// ------------------------------------------------------------------------------------------------ small linears
__device__ __forceinline__ float act_f(float v, int act) {
    if (act == 1) return v / (1.0f + expf(-v));
    if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    return v;
}
__device__ __forceinline__ float act_grad_f(float v, int act) {
    if (act == 1) { const float s = 1.0f / (1.0f + expf(-v)); return s * (1.0f + v * (1.0f - s)); }
    if (act == 2) return 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
    return 1.0f;
}

// dW[j, k] = sum_r dY[r, off + j] * act(X[r, k]);  db[j] = sum_r dY[r, off + j].  grid (ceil(in_f / 256), out_f / 8): a thread
// owns column k for 8 output features, so each X element is loaded once per 8 (not once per 1) features
constexpr int LBW_J = 8;
__global__ void __launch_bounds__(256)
linear_bwd_weight_kernel(const float* __restrict__ dY, int ldy, int off, const float* __restrict__ X, int ldx, int rows,
                         int in_f, int in_act, float* __restrict__ dW, float* __restrict__ db) {
    const int j0 = blockIdx.y * LBW_J;
    const int k = blockIdx.x * 256 + threadIdx.x;
    float acc[LBW_J], accb[LBW_J];
#pragma unroll
    for (int j = 0; j < LBW_J; ++j) { acc[j] = 0.f; accb[j] = 0.f; }
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
        const float xv = k < in_f ? act_f(__ldg(X + static_cast<size_t>(r) * ldx + k), in_act) : 0.f;
        const float* g = dY + static_cast<size_t>(r) * ldy + off + j0;
#pragma unroll
        for (int j = 0; j < LBW_J; ++j) {
            const float gv = __ldg(g + j);
            accb[j] += gv;
            acc[j] = fmaf(gv, xv, acc[j]);
        }
    }
    if (k < in_f) {
#pragma unroll
        for (int j = 0; j < LBW_J; ++j) dW[static_cast<size_t>(j0 + j) * in_f + k] = acc[j];
    }
    if (k == 0 && db != nullptr) {
#pragma unroll
        for (int j = 0; j < LBW_J; ++j) db[j0 + j] = accb[j];
    }
}

// dX[r, k] (+)= sum_j dY[r, off + j] * W[j, k].  grid (ceil(in_f / 256), rows)
__global__ void __launch_bounds__(256)
linear_bwd_input_kernel(const float* __restrict__ dY, int ldy, int off, const float* __restrict__ W, int out_f, int in_f,
                        int accumulate, float* __restrict__ dX, int ldx) {
    const int r = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= in_f) return;
    float acc = 0.f;
#pragma unroll 16
    for (int j = 0; j < out_f; ++j) acc = fmaf(__ldg(dY + static_cast<size_t>(r) * ldy + off + j), __ldg(W + static_cast<size_t>(j) * in_f + k), acc);
    float* o = dX + static_cast<size_t>(r) * ldx + k;
    *o = accumulate ? *o + acc : acc;
}

// y[i] = act(x[i])
__global__ void act_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int act) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i < n) y[i] = act_f(x[i], act);
}

// ---- the per-block time-embedding Linears, all slots in ONE launch each (19 / 32 launches of the single-slot kernels were
// ~36 us of latency apiece: 8 % of the Unet training step)
__global__ void __launch_bounds__(256)
linear_bwd_weight_batched_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, int rows, int in_f,
                                 const LinSlot* __restrict__ slots) {
    const LinSlot sl = slots[blockIdx.z];
    const int j0 = blockIdx.y * LBW_J;
    if (j0 >= sl.width) return;
    const int k = blockIdx.x * 256 + threadIdx.x;
    float acc[LBW_J], accb[LBW_J];
#pragma unroll
    for (int j = 0; j < LBW_J; ++j) { acc[j] = 0.f; accb[j] = 0.f; }
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
        const float xv = k < in_f ? __ldg(X + static_cast<size_t>(r) * ldx + k) : 0.f;
        const float* g = dY + static_cast<size_t>(r) * ldy + sl.off + j0;
#pragma unroll
        for (int j = 0; j < LBW_J; ++j) {
            const float gv = __ldg(g + j);
            accb[j] += gv;
            acc[j] = fmaf(gv, xv, acc[j]);
        }
    }
    if (k < in_f) {
#pragma unroll
        for (int j = 0; j < LBW_J; ++j) sl.gW[static_cast<size_t>(j0 + j) * in_f + k] = acc[j];
    }
    if (k == 0) {
#pragma unroll
        for (int j = 0; j < LBW_J; ++j) sl.gb[j0 + j] = accb[j];
    }
}
// part[slot][r][k] = sum_j dY[r, off + j] * W_slot[j, k]; grid (ceil(in_f / 256), rows, nslots)
__global__ void __launch_bounds__(256)
linear_bwd_input_batched_kernel(const float* __restrict__ dY, int ldy, int rows, int in_f, const LinSlot* __restrict__ slots,
                                float* __restrict__ part) {
    const LinSlot sl = slots[blockIdx.z];
    const int r = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= in_f) return;
    const float* g = dY + static_cast<size_t>(r) * ldy + sl.off;
    float acc = 0.f;
#pragma unroll 16
    for (int j = 0; j < sl.width; ++j) acc = fmaf(__ldg(g + j), __ldg(sl.W + static_cast<size_t>(j) * in_f + k), acc);
    part[(static_cast<size_t>(blockIdx.z) * rows + r) * in_f + k] = acc;
}

// d[i] *= act'(x[i])
__global__ void act_grad_kernel(float* __restrict__ d, const float* __restrict__ x, long long n, int act) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i < n) d[i] *= act_grad_f(x[i], act);
}

// ---- fused Adam (train.py:111 torch.optim.Adam) over many parameter tensors in one launch ----
// The arithmetic follows torch's own multi-tensor Adam operation by operation (lerp, mul, addcmul, sqrt, div, add, addcdiv with
// the scalars rounded to fp32 the way its kernels receive them) so a model trained with either optimiser follows the same path.
struct AdamScalars {
    float w1, beta2, w2, bc2_sqrt, eps, neg_step, wd;
};
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamScalars& a) {
    if (a.wd != 0.f) g = fmaf(a.wd, p, g);
    const float d = g - m;
    m = a.w1 < 0.5f ? fmaf(a.w1, d, m) : g - d * (1.f - a.w1);
    v = __fmul_rn(v, a.beta2);
    v = fmaf(a.w2, __fmul_rn(g, g), v);
    const float den = __fadd_rn(__fdiv_rn(sqrtf(v), a.bc2_sqrt), a.eps);
    p = fmaf(a.neg_step, __fdiv_rn(m, den), p);
}
__global__ void __launch_bounds__(256) adam_step_kernel(const AdamChunk* __restrict__ chunks, float* __restrict__ exp_avg,
                                                        float* __restrict__ exp_avg_sq, AdamScalars a) {
    const AdamChunk c = chunks[blockIdx.x];
    float* p = c.param;
    const float* g = c.grad;
    float* m = exp_avg + c.state_off;
    float* v = exp_avg_sq + c.state_off;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const int n4 = vec ? c.n / 4 : 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        adam_update(pv.x, gv.x, mv.x, vv.x, a);
        adam_update(pv.y, gv.y, mv.y, vv.y, a);
        adam_update(pv.z, gv.z, mv.z, vv.z, a);
        adam_update(pv.w, gv.w, mv.w, vv.w, a);
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = n4 * 4 + threadIdx.x; i < c.n; i += blockDim.x) adam_update(p[i], g[i], m[i], v[i], a);
}

}  // namespace

// ================================================================================================ launchers
cudaError_t prep_dgrad_weight_run(const float* w, bf16* out, int Cout, int Cin, cudaStream_t s) {
    prep_dgrad_weight_kernel<<<Cin, 256, 0, s>>>(w, out, Cout, Cin);
    return cudaGetLastError();
}
cudaError_t flip_tail_weight_run(const float* w, float* out, int C, cudaStream_t s) {
    flip_tail_weight_kernel<<<(C * 9 + 255) / 256, 256, 0, s>>>(w, out, C);
    return cudaGetLastError();
}
cudaError_t film_silu_fwd_run(const bf16* a, bf16* sout, const float* film, int ld, int off, int B, int P, int C, int has_scale,
                              cudaStream_t s) {
    if (C != 256 || P % (FB_CHUNKS * 32) != 0) return cudaErrorInvalidValue;
    film_silu_fwd_kernel<<<dim3(FB_CHUNKS, B), 256, 0, s>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<uint4*>(sout), film, ld, off, P, has_scale);
    return cudaGetLastError();
}
int film_bwd_part_floats(int B, int C) { return B * FB_CHUNKS * 2 * C; }
cudaError_t film_silu_bwd_run(const bf16* ds, const bf16* a, bf16* da, const float* film, float* dfilm, int ld, int off, int B,
                              int P, int C, float* part, int has_scale, cudaStream_t s) {
    if (C != 256 || P % (FB_CHUNKS * 32) != 0) return cudaErrorInvalidValue;
    film_silu_bwd_kernel<<<dim3(FB_CHUNKS, B), 256, 0, s>>>(reinterpret_cast<const uint4*>(ds), reinterpret_cast<const uint4*>(a),
                                                           reinterpret_cast<uint4*>(da), film, ld, off, P, part, has_scale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    film_grad_finish_kernel<<<B, 2 * C, 0, s>>>(part, dfilm, ld, off, C, has_scale);
    return cudaGetLastError();
}
cudaError_t edrn_bias_grad_run(const float* film, const float* dfilm, int ld, int off, int B, const float* colsum_g, float g_scale,
                               float* dbias, int C, int has_scale, cudaStream_t s) {
    edrn_bias_grad_kernel<<<(C + 31) / 32, 256, 0, s>>>(film, dfilm, ld, off, B, colsum_g, g_scale, dbias, C, has_scale);
    return cudaGetLastError();
}
int colsum_parts(long long M) { return static_cast<int>(M / 256 < 592 ? (M + 255) / 256 : 592); }
cudaError_t colsum_run(const bf16* x, long long M, int C, float* part, float scale, int accumulate, float* out, cudaStream_t s) {
    if (C % 8 != 0 || 256 % (C / 8) != 0) return cudaErrorInvalidValue;
    const int nparts = colsum_parts(M);
    const int rpb = static_cast<int>((M + nparts - 1) / nparts);
    const int lanes = 256 / (C / 8);
    colsum_part_kernel<<<nparts, 256, lanes * C * sizeof(float), s>>>(reinterpret_cast<const uint4*>(x), part, M, C, rpb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_parts_kernel<<<(C + 31) / 32, 32 * SUM_SLICES, 0, s>>>(part, nparts, C, scale, accumulate, out);
    return cudaGetLastError();
}
cudaError_t sum_parts_run(const float* part, int nparts, int n, float scale, int accumulate, float* out, cudaStream_t s) {
    sum_parts_kernel<<<(n + 31) / 32, 32 * SUM_SLICES, 0, s>>>(part, nparts, n, scale, accumulate, out);
    return cudaGetLastError();
}
cudaError_t add_bf16_run(const bf16* a, const bf16* b, bf16* y, long long n, cudaStream_t s) {
    const long long chunks = n / 8;
    const int grid = static_cast<int>(chunks / 256 < 148 * 16 ? (chunks + 255) / 256 : 148 * 16);
    add_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(y), chunks);
    return cudaGetLastError();
}
cudaError_t thin_wgrad_run(const bf16* G, const float* u0, const float* u1, int sgn, int B, float* part, float* dw, cudaStream_t s) {
    const int nk = u1 != nullptr ? 2 : 1;
    thin_wgrad_kernel<<<dim3(8, B), 256, 0, s>>>(G, u0, u1, sgn, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    thin_wgrad_finish_kernel<<<(nk * 9 * 256 + 31) / 32, 1024, 0, s>>>(part, B * 8, nk, dw);
    return cudaGetLastError();
}
int loss_parts() { return 592; }
cudaError_t loss_grad_run(const float* eps, const float* target, const float* w, int loss_type, int B, int tile_elems, float* d_eps,
                          float* part, float* loss, cudaStream_t s) {
    const long long n = static_cast<long long>(B) * tile_elems;
    loss_grad_kernel<<<loss_parts(), 256, 0, s>>>(eps, target, w, loss_type, n, tile_elems, d_eps, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_scalar_kernel<<<1, 32, 0, s>>>(part, loss_parts(), 1.0f / static_cast<float>(n), loss);
    return cudaGetLastError();
}
// out[0] = sum x[0..n); part: loss_parts() floats
cudaError_t sum_f32_run(const float* x, long long n, float* part, float* out, cudaStream_t s) {
    sum_f32_part_kernel<<<loss_parts(), 256, 0, s>>>(x, n, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_scalar_kernel<<<1, 32, 0, s>>>(part, loss_parts(), 1.0f, out);
    return cudaGetLastError();
}
cudaError_t linear_bwd_weight_run(const float* dY, int ldy, int off, const float* X, int ldx, int rows, int in_f, int out_f,
                                  int in_act, float* dW, float* db, cudaStream_t s) {
    if (out_f % LBW_J != 0) return cudaErrorInvalidValue;
    linear_bwd_weight_kernel<<<dim3((in_f + 255) / 256, out_f / LBW_J), 256, 0, s>>>(dY, ldy, off, X, ldx, rows, in_f, in_act, dW, db);
    return cudaGetLastError();
}
cudaError_t linear_bwd_input_run(const float* dY, int ldy, int off, const float* W, int rows, int in_f, int out_f, int accumulate,
                                 float* dX, int ldx, cudaStream_t s) {
    linear_bwd_input_kernel<<<dim3((in_f + 255) / 256, rows), 256, 0, s>>>(dY, ldy, off, W, out_f, in_f, accumulate, dX, ldx);
    return cudaGetLastError();
}
// weight / bias gradients of all slots (X is shared: the activated time embedding); every slot width must be a multiple of 8
cudaError_t linear_bwd_weight_batched_run(const float* dY, int ldy, const float* X, int ldx, int rows, int in_f, const LinSlot* slots,
                                          int nslots, int max_width, cudaStream_t s) {
    if (max_width % LBW_J != 0) return cudaErrorInvalidValue;
    linear_bwd_weight_batched_kernel<<<dim3((in_f + 255) / 256, max_width / LBW_J, nslots), 256, 0, s>>>(dY, ldy, X, ldx, rows, in_f, slots);
    return cudaGetLastError();
}
// dX[r, k] = sum over the slots (in slot order) of dY_slot W_slot; part: nslots * rows * in_f floats
cudaError_t linear_bwd_input_batched_run(const float* dY, int ldy, int rows, int in_f, const LinSlot* slots, int nslots, float* part,
                                         float* dX, cudaStream_t s) {
    linear_bwd_input_batched_kernel<<<dim3((in_f + 255) / 256, rows, nslots), 256, 0, s>>>(dY, ldy, rows, in_f, slots, part);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return sum_parts_run(part, nslots, rows * in_f, 1.0f, 0, dX, s);
}
// one Adam update; hyper-parameters in double as Python holds them, rounded to fp32 exactly where torch rounds them
cudaError_t adam_step_run(const AdamChunk* chunks, int nchunks, float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2,
                          double eps, double weight_decay, long long step, cudaStream_t s) {
    if (nchunks <= 0) return cudaSuccess;
    const double bc1 = 1.0 - pow(beta1, static_cast<double>(step)), bc2 = 1.0 - pow(beta2, static_cast<double>(step));
    AdamScalars a;
    a.w1 = static_cast<float>(1.0 - beta1);
    a.beta2 = static_cast<float>(beta2);
    a.w2 = static_cast<float>(1.0 - beta2);
    a.bc2_sqrt = static_cast<float>(pow(bc2, 0.5));
    a.eps = static_cast<float>(eps);
    a.neg_step = static_cast<float>((lr / bc1) * -1.0);
    a.wd = static_cast<float>(weight_decay);
    adam_step_kernel<<<nchunks, 256, 0, s>>>(chunks, exp_avg, exp_avg_sq, a);
    return cudaGetLastError();
}
cudaError_t act_apply_run(const float* x, float* y, long long n, int act, cudaStream_t s) {
    act_apply_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(x, y, n, act);
    return cudaGetLastError();
}
cudaError_t act_grad_run(float* d, const float* x, long long n, int act, cudaStream_t s) {
    act_grad_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(d, x, n, act);
    return cudaGetLastError();
}

}  // namespace hd
