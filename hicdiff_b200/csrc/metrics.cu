// Per-tile SSIM and MSE of 64x64 tiles on the GPU (the measurement side of the path).
//
// reference: _ssim  /root/reference/src/Utils/loss/SSIM.py:17-36 (11x11 gaussian window sigma 1.5, zero "same" padding,
//                   C1 = 0.01^2, C2 = 0.03^2, mean of the SSIM map)
//            inverse_data_transform('rescaled')  /root/reference/src/datasets/__init__.py:214-223: clamp((x + 1) / 2, 0, 1)
//            PSNR = 10 log10(1 / mse)  /root/reference/pretrain/train_unet_Diff_cond_n.py:125-133
// One CTA per tile: both tiles are staged in shared memory (optionally rescaled), every thread evaluates the five windowed
// moments of 16 pixels with the reference's own 2-D window (passed in, 121 floats) and the CTA reduces the SSIM map and
// the squared error in a fixed order.  The caller averages the per-tile values (tiles have equal size, so the mean of the
// per-tile means IS the reference's global mean).
#include "kernels.h"

namespace hd {
namespace {

constexpr int MT = 64;           // tile edge
constexpr int MWIN = 11;
constexpr int MPAD = MWIN / 2;
constexpr int MPITCH = MT + 2 * MPAD;      // zero halo staged in shared memory

__global__ void __launch_bounds__(256)
ssim_mse_tiles_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ window,
                      float* __restrict__ ssim_out, float* __restrict__ mse_out, int rescale) {
    extern __shared__ float msm[];
    float* sa = msm;                               // [MPITCH][MPITCH]
    float* sb = sa + MPITCH * MPITCH;
    float* sw = sb + MPITCH * MPITCH;              // [121]
    __shared__ float s_red[2][8];
    const int tile = blockIdx.x;
    const int tid = threadIdx.x;
    const float* ta = a + static_cast<size_t>(tile) * MT * MT;
    const float* tb = b + static_cast<size_t>(tile) * MT * MT;
    for (int i = tid; i < MPITCH * MPITCH; i += 256) {
        const int y = i / MPITCH - MPAD, x = i % MPITCH - MPAD;
        float va = 0.f, vb = 0.f;
        if (y >= 0 && y < MT && x >= 0 && x < MT) {
            va = __ldg(ta + y * MT + x);
            vb = __ldg(tb + y * MT + x);
            if (rescale) {
                va = fminf(fmaxf((va + 1.0f) / 2.0f, 0.0f), 1.0f);
                vb = fminf(fmaxf((vb + 1.0f) / 2.0f, 0.0f), 1.0f);
            }
        }
        sa[i] = va;
        sb[i] = vb;
    }
    if (tid < MWIN * MWIN) sw[tid] = __ldg(window + tid);
    __syncthreads();
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float ssim_sum = 0.f, se_sum = 0.f;
    for (int p = tid; p < MT * MT; p += 256) {
        const int y = p / MT, x = p % MT;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
        for (int dy = 0; dy < MWIN; ++dy) {
            const float* ra = sa + (y + dy) * MPITCH + x;
            const float* rb = sb + (y + dy) * MPITCH + x;
#pragma unroll
            for (int dx = 0; dx < MWIN; ++dx) {
                const float w = sw[dy * MWIN + dx];
                const float va = ra[dx], vb = rb[dx];
                mu1 = fmaf(w, va, mu1);
                mu2 = fmaf(w, vb, mu2);
                e11 = fmaf(w, va * va, e11);
                e22 = fmaf(w, vb * vb, e22);
                e12 = fmaf(w, va * vb, e12);
            }
        }
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
        ssim_sum += ((2.0f * mu12 + C1) * (2.0f * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2));
        const float d = sa[(y + MPAD) * MPITCH + x + MPAD] - sb[(y + MPAD) * MPITCH + x + MPAD];
        se_sum = fmaf(d, d, se_sum);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ssim_sum += __shfl_xor_sync(0xffffffffu, ssim_sum, off);
        se_sum += __shfl_xor_sync(0xffffffffu, se_sum, off);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = ssim_sum; s_red[1][tid >> 5] = se_sum; }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f, e = 0.f;
        for (int w = 0; w < 8; ++w) { s += s_red[0][w]; e += s_red[1][w]; }
        ssim_out[tile] = s / (MT * MT);
        mse_out[tile] = e / (MT * MT);
    }
}

}  // namespace

cudaError_t ssim_mse_tiles_run(const float* a, const float* b, const float* window, float* ssim_out, float* mse_out, int B,
                               int rescale, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = (2 * MPITCH * MPITCH + MWIN * MWIN) * sizeof(float);   // ~45 KiB
    ssim_mse_tiles_kernel<<<B, 256, smem, s>>>(a, b, window, ssim_out, mse_out, rescale);
    return cudaGetLastError();
}

}  // namespace hd
