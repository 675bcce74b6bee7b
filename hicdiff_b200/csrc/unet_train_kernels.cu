// Glue kernels of the Unet training step: the resampling layers as explicit copies, the thin convs at the
// two ends of the net, and the data-gradient weight layout with the weight standardisation applied.
//
//   Downsample  'b c (h p1) (w p2) -> b (c p1 p2) h w' + 1x1 conv     /root/reference/src/hicdiff_condition.py:78-82
//   Upsample    nearest x2 + 3x3 conv                                 :72-76
//   init_conv   7x7, Cin = 1 or 2 -> dim                              :278-279
//   final_conv  1x1, dim -> 1                                         :343
// (The sampling path folds both resampling layers into the conv's TMA addressing; training materialises the rearranged
// tensor once so that the conv, its dgrad and its wgrad are the plain kernels.)
#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

// out[b, h, w, c * 4 + p1 * 2 + p2] = in[b, 2h + p1, 2w + p2, c]   (inverse == 1: the scatter back, out is [B, 2H, 2W, C])
// One thread per (low-resolution pixel, 8 channels): four 16-byte chunks on the high-resolution side (one per (p1, p2), 8
// consecutive channels each) <-> four consecutive 16-byte chunks on the low-resolution side, transposed in registers.  (The
// 2-byte-per-thread gather this replaces ran at a tenth of the copy bandwidth.)
__global__ void __launch_bounds__(256)
unshuffle_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int C, int inverse) {
    const int cpp = C / 8;
    const long long total = static_cast<long long>(B) * H * W * cpp;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int cc = static_cast<int>(i % cpp);
        const long long pix = i / cpp;
        const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H);
        const long long b = pix / (static_cast<long long>(W) * H);
        long long hi[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) hi[p] = ((b * 2 * H + 2 * h + (p >> 1)) * 2 * W + 2 * w + (p & 1)) * cpp + cc;
        const long long lo = pix * 4 * cpp + cc * 4;          // chunk index of channel (cc * 8) * 4 in the [.., 4C] row
        union { uint4 v[4]; unsigned short e[32]; } a, t;
        if (!inverse) {
#pragma unroll
            for (int p = 0; p < 4; ++p) a.v[p] = __ldg(in + hi[p]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int p = 0; p < 4; ++p) t.e[k * 4 + p] = a.e[p * 8 + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) out[lo + q] = t.v[q];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) a.v[q] = __ldg(in + lo + q);
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int p = 0; p < 4; ++p) t.e[p * 8 + k] = a.e[k * 4 + p];
#pragma unroll
            for (int p = 0; p < 4; ++p) out[hi[p]] = t.v[p];
        }
    }
}

// nearest x2: out[b, y, x, :] = in[b, y / 2, x / 2, :]   (16-byte chunks)
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int cpp) {
    const long long total = static_cast<long long>(B) * 2 * H * 2 * W * cpp;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int cc = static_cast<int>(i % cpp);
        const long long pix = i / cpp;
        const int x = static_cast<int>(pix % (2 * W)), y = static_cast<int>((pix / (2 * W)) % (2 * H));
        const long long b = pix / (static_cast<long long>(4) * W * H);
        out[i] = __ldg(in + ((b * H + y / 2) * W + x / 2) * cpp + cc);
    }
}
// its backward: out[b, h, w, :] = sum of the 2x2 block of in [B, 2H, 2W, C]
__global__ void __launch_bounds__(256)
sumpool2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int cpp) {
    const long long total = static_cast<long long>(B) * H * W * cpp;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int cc = static_cast<int>(i % cpp);
        const long long pix = i / cpp;
        const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H);
        const long long b = pix / (static_cast<long long>(W) * H);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 u = __ldg(in + ((b * 2 * H + 2 * h + (q >> 1)) * 2 * W + 2 * w + (q & 1)) * cpp + cc);
            float2 t;
            t = ptx::unpack_bf16x2(u.x); acc[0] += t.x; acc[1] += t.y;
            t = ptx::unpack_bf16x2(u.y); acc[2] += t.x; acc[3] += t.y;
            t = ptx::unpack_bf16x2(u.z); acc[4] += t.x; acc[5] += t.y;
            t = ptx::unpack_bf16x2(u.w); acc[6] += t.x; acc[7] += t.y;
        }
        uint4 o;
        o.x = ptx::pack_bf16x2(acc[0], acc[1]); o.y = ptx::pack_bf16x2(acc[2], acc[3]);
        o.z = ptx::pack_bf16x2(acc[4], acc[5]); o.w = ptx::pack_bf16x2(acc[6], acc[7]);
        out[i] = o;
    }
}

// ---- init_conv weight gradient: part[(b * 8 + rg)][k][tap][c] over 8 image rows; G [B, 64, 64, C] bf16, u_k fp32 planes, k x k taps.
// grid (8, B, ksize): blockIdx.z = filter row ky; threads = channels (C <= 256)
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const bf16* __restrict__ G, const float* __restrict__ u0, const float* __restrict__ u1, int C, int ksize,
                  float* __restrict__ part) {
    constexpr int W = 64, H = 64, ROWS = 8;
    extern __shared__ float sw_u[];                      // [nk][ROWS][W + ksize - 1] rows y0 + ky - pad .. (only the ky of this block)
    const int pad = ksize / 2, pitch = W + ksize - 1;
    const int rg = blockIdx.x, b = blockIdx.y, ky = blockIdx.z, c = threadIdx.x;
    const int nk = u1 != nullptr ? 2 : 1;
    const int y0 = rg * ROWS;
    for (int i = threadIdx.x; i < nk * ROWS * pitch; i += blockDim.x) {
        const int k = i / (ROWS * pitch), rem = i - k * ROWS * pitch;
        const int yy = y0 + rem / pitch + ky - pad, xx = rem % pitch - pad;
        const float* u = k == 0 ? u0 : u1;
        sw_u[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(u + (static_cast<size_t>(b) * H + yy) * W + xx) : 0.f;
    }
    __syncthreads();
    if (c >= C) return;
    float acc[2][7];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int t = 0; t < 7; ++t) acc[k][t] = 0.f;
    const bf16* g = G + ((static_cast<size_t>(b) * H + y0) * W) * C + c;
    // eight pixels per step: their gradients are loaded together and the (8 + ksize - 1)-wide window of u is read from shared
    // memory once per step instead of once per (pixel, tap) -- the loop was bound by broadcast LDS issue, one per FMA
    for (int yl = 0; yl < ROWS; ++yl)
        for (int x0 = 0; x0 < W; x0 += 8) {
            float gv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) gv[i] = __bfloat162float(g[(static_cast<size_t>(yl) * W + x0 + i) * C]);
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (k < nk) {
                    const float* ur = sw_u + (k * ROWS + yl) * pitch + x0;
                    float uw[14];
#pragma unroll
                    for (int j = 0; j < 14; ++j) uw[j] = x0 + j < pitch ? ur[j] : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int t = 0; t < 7; ++t)
                            if (t < ksize) acc[k][t] = fmaf(gv[i], uw[i + t], acc[k][t]);
                }
        }
    float* o = part + ((static_cast<size_t>(b) * 8 + rg) * nk * ksize * ksize) * C;
    for (int k = 0; k < nk; ++k)
        for (int t = 0; t < ksize; ++t) o[((k * ksize + ky) * ksize + t) * C + c] = acc[k][t];
}
// dw[(c * nk + k) * taps + tap] = sum_parts part[.][k][tap][c]: 32 columns per block, 32 threads per column over contiguous
// slices of the partials (eight loads in flight each), combined in a fixed order
__global__ void __launch_bounds__(1024)
stem_wgrad_finish_kernel(const float* __restrict__ part, int nparts, int nk, int taps, int C, float* __restrict__ dw) {
    __shared__ float s_p[32][32];
    const int n = nk * taps * C;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;
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
        const int c = i % C, kt = i / C, kk = kt / taps, tap = kt - kk * taps;
        dw[(c * nk + kk) * taps + tap] = v;
    }
}

// ---- final_conv (C -> 1) backward: dx[m, c] = d_eps[m] * w[c];  dw partials part[blk][c] = sum_m d_eps[m] x[m, c]
__global__ void __launch_bounds__(256)
head_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ d_eps, const float* __restrict__ w, long long M, int C,
                bf16* __restrict__ dx, float* __restrict__ part, int rpb) {
    extern __shared__ float hb_red[];    // [lanes][C]
    const int lanes = 256 / C;           // C = 64 -> 4 row lanes
    const int c = threadIdx.x % C, pl = threadIdx.x / C;
    const float wc = w[c];
    const long long r0 = static_cast<long long>(blockIdx.x) * rpb, r1 = r0 + rpb < M ? r0 + rpb : M;
    float acc = 0.f;
    for (long long m = r0 + pl; m < r1; m += lanes) {
        const float g = d_eps[m];
        acc = fmaf(g, __bfloat162float(x[m * C + c]), acc);
        dx[m * C + c] = __float2bfloat16(g * wc);
    }
    hb_red[pl * C + c] = acc;
    __syncthreads();
    if (pl == 0) {
        float t = 0.f;
        for (int k = 0; k < lanes; ++k) t += hb_red[k * C + c];
        part[static_cast<size_t>(blockIdx.x) * C + c] = t;
    }
}

inline int grid_for(long long n) { return static_cast<int>(n / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16); }

}  // namespace

cudaError_t unshuffle_run(const bf16* in, bf16* out, int B, int H, int W, int C, int inverse, cudaStream_t s) {
    if (C % 8 != 0) return cudaErrorInvalidValue;
    unshuffle_kernel<<<grid_for(static_cast<long long>(B) * H * W * (C / 8)), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), B, H, W, C, inverse);
    return cudaGetLastError();
}
cudaError_t upsample2x_run(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s) {
    upsample2x_kernel<<<grid_for(static_cast<long long>(B) * 4 * H * W * (C / 8)), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), B, H, W, C / 8);
    return cudaGetLastError();
}
cudaError_t sumpool2x_run(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s) {
    sumpool2x_kernel<<<grid_for(static_cast<long long>(B) * H * W * (C / 8)), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), B, H, W, C / 8);
    return cudaGetLastError();
}
// G [B, 64, 64, C] (C <= 256); part: B * 8 * nk * k * k * C floats; dw in the reference layout [C, nk, k, k]
cudaError_t stem_wgrad_run(const bf16* G, const float* u0, const float* u1, int B, int C, int ksize, float* part, float* dw, cudaStream_t s) {
    if (C > 256 || ksize > 7 || (ksize & 1) == 0) return cudaErrorInvalidValue;
    const int nk = u1 != nullptr ? 2 : 1;
    const size_t smem = static_cast<size_t>(nk) * 8 * (64 + ksize - 1) * sizeof(float);
    stem_wgrad_kernel<<<dim3(8, B, ksize), (C + 31) / 32 * 32, smem, s>>>(G, u0, u1, C, ksize, part);   // one thread per channel
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n = nk * ksize * ksize * C;
    stem_wgrad_finish_kernel<<<(n + 31) / 32, 1024, 0, s>>>(part, B * 8, nk, ksize * ksize, C, dw);
    return cudaGetLastError();
}
int head_bwd_parts(long long M) { return static_cast<int>(M / 64 < 592 ? (M + 63) / 64 : 592); }
// x [M, C] (256 % C == 0), dx [M, C], dw [C] (via part: head_bwd_parts(M) * C floats)
cudaError_t head_bwd_run(const bf16* x, const float* d_eps, const float* w, long long M, int C, bf16* dx, float* part, float* dw,
                         cudaStream_t s) {
    if (C > 256 || 256 % C != 0) return cudaErrorInvalidValue;
    const int nparts = head_bwd_parts(M);
    const int rpb = static_cast<int>((M + nparts - 1) / nparts);
    head_bwd_kernel<<<nparts, 256, 256 * sizeof(float), s>>>(x, d_eps, w, M, C, dx, part, rpb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return sum_parts_run(part, nparts, C, 1.0f, 0, dw, s);
}

}  // namespace hd
