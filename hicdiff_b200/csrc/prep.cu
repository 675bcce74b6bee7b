// One-off (per checkpoint) preparation kernels and the time-embedding path.
//
// prep_conv_weight    WeightStandardizedConv2d's weight normalisation, hoisted out of the step because it does
//                     not depend on the input (/root/reference/src/hicdiff_condition.py:89-95): per output channel
//                     w~ = (w - mean) * rsqrt(var_biased + 1e-5) in fp32, then re-laid out [Cout, (ky, kx, cin)]
//                     bf16 = the K-major B operand of conv_gemm.cu.  standardize = 0 gives the plain re-layout.
// prep_unshuffle      Downsample's 1x1 weight: K order (c, p1, p2) -> (p1, p2, c)   (:78-82)
// posenc_rows         SinusoidalPosEmb (:122-134) / SR3 PositionalEncoding (hicdiff_sr3.py:155-165)
// linear_rows         nn.Linear over a handful of rows with optional SiLU on the input (ResnetBlock.mlp :176-179)
//                     or exact-erf GELU on the output (time_mlp :300-305)
// Because p_sample feeds the SAME t to every sample (:594), the sampling path evaluates these once per
// checkpoint for all T timesteps and keeps the [T, sum(2*Cout)] FiLM table resident in HBM.
#include "kernels.h"
#include "ptx.cuh"

namespace hd {
namespace {

// one output channel (GEMM row) per CTA; stat_out (optional) receives the row's (mean, rstd) when standardising
//
// split = 1 ("bf16w2" precision): every weight is kept as hi + lo, hi = bf16(v), lo = bf16(v - hi) (together ~16
// mantissa bits); a GEMM row is then 2K long, per tap [hi(0..Cin) | lo(0..Cin)] -- the conv walks its input channels twice per
// tap (conv_gemm.cu, KArgs::wrap), so the product is x * hi + x * lo with fp32 accumulation.
__device__ __forceinline__ void store_weight(bf16* __restrict__ orow, int tap, int ci, int Cin, float v, int split) {
    const bf16 hi = __float2bfloat16(v);
    if (split) {
        orow[(2 * tap) * Cin + ci] = hi;
        orow[(2 * tap + 1) * Cin + ci] = __float2bfloat16(v - __bfloat162float(hi));
    } else {
        orow[tap * Cin + ci] = hi;
    }
}

__device__ __forceinline__ void prep_conv_weight_row(const float* __restrict__ w, bf16* __restrict__ out, const int co, int Cout, int Cin,
                                                     int ksize, int standardize, float eps, float2* stat_out, int split = 0) {
    __shared__ float s_red[8];
    __shared__ float s_stat[2];
    const int tid = threadIdx.x;
    const int kk = ksize * ksize;
    const int K = Cin * kk;
    const int Kout = split ? 2 * K : K;
    bf16* orow = out + static_cast<size_t>(co) * Kout;
    if (co >= Cout) {   // zero padding rows (N padded up to the GEMM tile)
        for (int i = tid; i < Kout; i += blockDim.x) orow[i] = __float2bfloat16(0.f);
        return;
    }
    const float* wrow = w + static_cast<size_t>(co) * K;
    float mean = 0.f, rstd = 1.f;
    if (standardize) {
        float s = 0.f;
        for (int i = tid; i < K; i += blockDim.x) s += wrow[i];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((tid & 31) == 0) s_red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; ++i) t += s_red[i];
            s_stat[0] = t / K;
        }
        __syncthreads();
        mean = s_stat[0];
        float ss = 0.f;
        for (int i = tid; i < K; i += blockDim.x) ss += (wrow[i] - mean) * (wrow[i] - mean);
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = ss;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; ++i) t += s_red[i];
            s_stat[1] = rsqrtf(t / K + eps);
        }
        __syncthreads();
        rstd = s_stat[1];
        if (stat_out != nullptr && tid == 0) *stat_out = make_float2(mean, rstd);
    }
    // reference layout [Cout][Cin][ky][kx]  ->  [Cout][(ky*ks + kx) * Cin + ci]
    for (int i = tid; i < K; i += blockDim.x) {
        const int tap = i / Cin;
        const int ci = i - tap * Cin;
        const float v = (wrow[ci * kk + tap] - mean) * rstd;
        store_weight(orow, tap, ci, Cin, v, split);
    }
}

__global__ void __launch_bounds__(256)
prep_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int ksize,
                        int standardize, float eps, int Npad, int split) {
    prep_conv_weight_row(w, out, blockIdx.x, Cout, Cin, ksize, standardize, eps, nullptr, split);
}

// ---- the training step's per-step weight preparation for ALL convs in two launches (it was three short launches per conv:
// ~180 graph nodes whose launch latency, not their work, was the cost).  rows[i] = (slot, row): forward rows are output
// channels (the standardised bf16 GEMM layout + the (mean, rstd) the backward reuses), backward rows are input channels.
__global__ void __launch_bounds__(256)
prep_weights_fwd_batched_kernel(const PrepSlot* __restrict__ slots, const int2* __restrict__ rows, float eps) {
    const int2 r = rows[blockIdx.x];
    const PrepSlot sl = slots[r.x];
    prep_conv_weight_row(sl.w, sl.qf, r.y, sl.Cout, sl.Cin, sl.ksize, sl.ws, eps, sl.ws ? sl.stats + r.y : nullptr);
}
// qd[ci][(tap', co)] = wt[co][ci][taps - 1 - tap'] (the flipped / transposed weight dgrad convolves with), wt standardised when ws
__global__ void __launch_bounds__(256)
prep_weights_bwd_batched_kernel(const PrepSlot* __restrict__ slots, const int2* __restrict__ rows) {
    const int2 r = rows[blockIdx.x];
    const PrepSlot sl = slots[r.x];
    const int ci = r.y, taps = sl.ksize * sl.ksize, Cout = sl.Cout, Cin = sl.Cin;
    bf16* orow = sl.qd + static_cast<size_t>(ci) * taps * Cout;
    for (int i = threadIdx.x; i < taps * Cout; i += blockDim.x) {
        const int tapp = i / Cout, co = i - tapp * Cout;
        float v = sl.w[(static_cast<size_t>(co) * Cin + ci) * taps + (taps - 1 - tapp)];
        if (sl.ws) { const float2 st = sl.stats[co]; v = (v - st.x) * st.y; }
        orow[i] = __float2bfloat16(v);
    }
}

__global__ void __launch_bounds__(256)
prep_unshuffle_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int C, int split) {
    const int co = blockIdx.x;
    const int K = 4 * C;
    bf16* orow = out + static_cast<size_t>(co) * (split ? 2 * K : K);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int tap = i / C;        // p1 * 2 + p2
        const int c = i - tap * C;
        store_weight(orow, tap, c, C, w[static_cast<size_t>(co) * K + c * 4 + tap], split);
    }
}

// Upsample = nearest x2 followed by a 3x3 conv (/root/reference/src/hicdiff_condition.py:72-76).  Output pixel
// (2i + a, 2j + b) only ever sees the low-res pixels (i + a - 1 + u, j + b - 1 + v), u, v in {0, 1}: original tap ky
// lands on u = (a + ky + 1) / 2 - a  (a = 0: ky 0 -> u 0, ky 1,2 -> u 1;  a = 1: ky 0,1 -> u 0, ky 2 -> u 1).
// out[phase = a*2 + b][co][(u*2 + v) * Cin + ci] = sum of the original taps that collapse onto (u, v), in fp32.
__global__ void __launch_bounds__(256)
prep_upsample_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int split) {
    const int co = blockIdx.x;
    const int phase = blockIdx.y;
    const int a = phase >> 1, b = phase & 1;
    const int K = 4 * Cin;
    bf16* orow = out + (static_cast<size_t>(phase) * Cout + co) * (split ? 2 * K : K);
    const float* wrow = w + static_cast<size_t>(co) * Cin * 9;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int tap = i / Cin;
        const int ci = i - tap * Cin;
        const int u = tap >> 1, v = tap & 1;
        float acc = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            if ((a + ky + 1) / 2 - a != u) continue;
            for (int kx = 0; kx < 3; ++kx) {
                if ((b + kx + 1) / 2 - b != v) continue;
                acc += wrow[ci * 9 + ky * 3 + kx];
            }
        }
        store_weight(orow, tap, ci, Cin, acc, split);
    }
}

__global__ void __launch_bounds__(128)
posenc_rows_kernel(const float* __restrict__ t, float* __restrict__ y, int rows, int dim, int mode) {
    const int r = blockIdx.x;
    const int half = dim / 2;
    const float tv = t[r];
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float freq;
        if (mode == 0) {
            const float emb = logf(10000.0f) / static_cast<float>(half - 1);
            freq = expf(static_cast<float>(i) * -emb);
        } else {
            const float step = static_cast<float>(i) / static_cast<float>(half);
            freq = expf(-logf(10000.0f) * step);
        }
        const float arg = tv * freq;
        y[static_cast<size_t>(r) * dim + i] = sinf(arg);
        y[static_cast<size_t>(r) * dim + half + i] = cosf(arg);
    }
}

// one warp per output feature, grid.y = row
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                   float* __restrict__ y, int ldy, int off, int rows, int in_f, int out_f, int in_act, int out_act) {
    const int r = blockIdx.y;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= out_f) return;
    const float* xr = x + static_cast<size_t>(r) * ldx;
    const float* wr = W + static_cast<size_t>(j) * in_f;
    float acc = 0.f;
    for (int k = lane; k < in_f; k += 32) {
        float xv = xr[k];
        if (in_act == 1) xv = xv / (1.0f + expf(-xv));
        else if (in_act == 2) xv = 0.5f * xv * (1.0f + erff(xv * 0.70710678118654752440f));
        acc = fmaf(xv, wr[k], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        float v = acc + (bias != nullptr ? bias[j] : 0.f);
        if (out_act == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        y[static_cast<size_t>(r) * ldy + off + j] = v;
    }
}

}  // namespace

cudaError_t prep_weights_batched_run(const PrepSlot* slots, const int2* fwd_rows, int nfwd, const int2* bwd_rows, int nbwd, float eps,
                                     cudaStream_t s) {
    if (nfwd > 0) prep_weights_fwd_batched_kernel<<<nfwd, 256, 0, s>>>(slots, fwd_rows, eps);
    if (nbwd > 0) prep_weights_bwd_batched_kernel<<<nbwd, 256, 0, s>>>(slots, bwd_rows);
    return cudaGetLastError();
}
cudaError_t prep_conv_weight_run(const float* w, bf16* out, int Cout, int Cin, int ksize, int standardize, float eps,
                                 int Npad, cudaStream_t s, int split) {
    prep_conv_weight_kernel<<<Npad, 256, 0, s>>>(w, out, Cout, Cin, ksize, standardize, eps, Npad, split);
    return cudaGetLastError();
}

cudaError_t prep_unshuffle_weight_run(const float* w, bf16* out, int Cout, int C, cudaStream_t s, int split) {
    prep_unshuffle_weight_kernel<<<Cout, 256, 0, s>>>(w, out, Cout, C, split);
    return cudaGetLastError();
}

cudaError_t prep_upsample_weight_run(const float* w, bf16* out, int Cout, int Cin, cudaStream_t s, int split) {
    prep_upsample_weight_kernel<<<dim3(Cout, 4), 256, 0, s>>>(w, out, Cout, Cin, split);
    return cudaGetLastError();
}

cudaError_t linear_rows_run(const float* x, int ldx, const float* W, const float* bias, float* y, int ldy, int off,
                            int rows, int in_f, int out_f, int in_act, int out_act, cudaStream_t s) {
    dim3 grid((out_f + 7) / 8, rows);
    linear_rows_kernel<<<grid, 256, 0, s>>>(x, ldx, W, bias, y, ldy, off, rows, in_f, out_f, in_act, out_act);
    return cudaGetLastError();
}

cudaError_t posenc_rows_run(const float* t, float* y, int rows, int dim, int mode, cudaStream_t s) {
    posenc_rows_kernel<<<rows, 128, 0, s>>>(t, y, rows, dim, mode);
    return cudaGetLastError();
}

}  // namespace hd
