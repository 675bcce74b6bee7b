// Fused normalisation kernels (HBM-bandwidth bound; one read + one write of the activation).
//
// groupnorm_film_silu: nn.GroupNorm(8, C, eps=1e-5) -> x*(scale+1)+shift -> SiLU [-> + postadd] [-> + residual]
//   reference: Block.forward   /root/reference/src/hicdiff_condition.py:162-171
//              ResnetBlock     :185-197 (the "+ res_conv(x)" is the optional residual operand here)
//              SR3 ResnetBlock /root/reference/src/hicdiff_sr3.py:246-251 (additive noise embedding = postadd)
//   One image (b) is owned by a thread-block CLUSTER of S CTAs; each CTA stages its contiguous slab of the
//   NHWC image (<= 64 KiB) in shared memory once, the per-group mean and then the centred sum of squares are
//   reduced with warp shuffles -> shared memory -> DSMEM across the cluster (two-pass variance, fp32), and the
//   normalised / modulated / activated result is written straight from shared memory.  No atomics, deterministic.
//
// channel_layernorm: LayerNorm over channels with gain g (eps 1e-5), optional residual add and optional
//   nearest-neighbour 2x upsample on the way out.
//   reference: LayerNorm :99-108, PreNorm :110-118, Residual :64-70, Upsample's nn.Upsample :74
#include <cooperative_groups.h>

#include "kernels.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace hd {
namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_GROUPS = 8;
constexpr int GN_MAX_SLAB = 64 * 1024;

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

struct GNK {
    GroupNormArgs a;
    int slab_chunks;   // 16-byte chunks per CTA slab
    int S;             // cluster size
};

__global__ void __launch_bounds__(GN_THREADS)
groupnorm_kernel(const GNK k) {
    extern __shared__ __align__(16) uint8_t gn_smem[];
    __shared__ float s_part[GN_GROUPS];      // this CTA's partial per group (read by cluster peers)
    __shared__ float s_wp[GN_THREADS / 32][GN_GROUPS];
    __shared__ float s_stat[GN_GROUPS];

    cg::cluster_group cluster = cg::this_cluster();
    const GroupNormArgs& a = k.a;
    const int S = k.S;
    const int b = blockIdx.x / S;
    const int part = blockIdx.x - b * S;
    const int tid = threadIdx.x;
    const int chunks_per_pixel = a.C / 8;                    // 16 B = 8 bf16
    const int cpg = a.C / GN_GROUPS;                         // channels per group (multiple of 8)
    const int my_cp = tid % chunks_per_pixel;                // fixed: 256 % chunks_per_pixel == 0
    const int my_c = my_cp * 8;                              // first channel this thread ever touches
    const int my_g = my_c / cpg;
    const size_t img_chunks = static_cast<size_t>(a.P) * chunks_per_pixel;
    const size_t base_chunk = static_cast<size_t>(b) * img_chunks + static_cast<size_t>(part) * k.slab_chunks;
    const uint4* xin = reinterpret_cast<const uint4*>(a.x) + base_chunk;
    uint4* sx = reinterpret_cast<uint4*>(gn_smem);
    const int nchunks = k.slab_chunks;
    const float inv_count = 1.0f / (static_cast<float>(a.P) * static_cast<float>(cpg));

    // ---- stage slab, pass 1: sum
    float acc = 0.f;
    for (int i = tid; i < nchunks; i += GN_THREADS) {
        const uint4 u = __ldg(xin + i);
        sx[i] = u;
        float2 t;
        t = ptx::unpack_bf16x2(u.x); acc += t.x + t.y;
        t = ptx::unpack_bf16x2(u.y); acc += t.x + t.y;
        t = ptx::unpack_bf16x2(u.z); acc += t.x + t.y;
        t = ptx::unpack_bf16x2(u.w); acc += t.x + t.y;
    }
    // Deterministic per-group reduction: warp shuffles -> per-warp slots -> fixed-order sum -> DSMEM across the
    // cluster (every CTA sums the S partials in rank order, so all CTAs of an image see identical statistics).
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int lanes_per_group = cpg / 8;                                   // consecutive chunk slots of one group
    const int lanes_active = chunks_per_pixel < 32 ? chunks_per_pixel : 32;  // distinct chunk slots inside a warp
    auto group_reduce = [&](float v) -> float {
        if (tid < 64) s_wp[tid >> 3][tid & 7] = 0.f;
        __syncthreads();
        for (int off = 16; off >= lanes_active; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        for (int off = lanes_per_group >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane < lanes_active && (lane & (lanes_per_group - 1)) == 0) s_wp[warp][my_g] = v;
        __syncthreads();
        if (tid < GN_GROUPS) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < GN_THREADS / 32; ++w) tot += s_wp[w][tid];
            s_part[tid] = tot;
            if (S == 1) s_stat[tid] = tot;
        }
        if (S > 1) {
            cluster.sync();
            if (tid < GN_GROUPS) {
                float tot = 0.f;
                for (int r = 0; r < S; ++r) tot += *cluster.map_shared_rank(&s_part[tid], r);
                s_stat[tid] = tot;
            }
            cluster.sync();   // peers are done reading s_part before it is rewritten / before this CTA exits
        } else {
            __syncthreads();
        }
        return s_stat[my_g];
    };
    const float mean = group_reduce(acc) * inv_count;

    // ---- pass 2: centred sum of squares (from shared memory)
    float acc2 = 0.f;
    for (int i = tid; i < nchunks; i += GN_THREADS) {
        const uint4 u = sx[i];
        float2 t;
        t = ptx::unpack_bf16x2(u.x); acc2 += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
        t = ptx::unpack_bf16x2(u.y); acc2 += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
        t = ptx::unpack_bf16x2(u.z); acc2 += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
        t = ptx::unpack_bf16x2(u.w); acc2 += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
    }
    const float var = group_reduce(acc2) * inv_count;
    const float rstd = rsqrtf(var + a.eps);

    // ---- per-thread channel constants (this thread always handles channels my_c .. my_c+7)
    float mul[8], add[8], post[8];
    {
        const float* frow = nullptr;
        const float* prow = nullptr;
        if (a.film != nullptr || a.postadd != nullptr) {
            const int row = a.film_row[b * a.film_row_stride];
            if (a.film != nullptr) frow = a.film + static_cast<size_t>(row) * a.film_ld + a.film_off;
            if (a.postadd != nullptr) prow = a.postadd + static_cast<size_t>(row) * a.film_ld + a.postadd_off;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = my_c + j;
            float gm = __ldg(a.gamma + c) * rstd;
            float bt = __ldg(a.beta + c) - mean * gm;
            if (frow != nullptr) {
                const float sc = __ldg(frow + c) + 1.0f;
                const float sh = __ldg(frow + a.C + c);
                gm *= sc;
                bt = bt * sc + sh;
            }
            mul[j] = gm;
            add[j] = bt;
            post[j] = prow != nullptr ? __ldg(prow + c) : 0.f;
        }
    }

    // ---- apply + write
    uint4* yout = reinterpret_cast<uint4*>(a.y) + base_chunk;
    const uint4* rin = a.res != nullptr ? reinterpret_cast<const uint4*>(a.res) + base_chunk : nullptr;
    for (int i = tid; i < nchunks; i += GN_THREADS) {
        const uint4 u = sx[i];
        float v[8];
        float2 t;
        t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
        t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
        t = ptx::unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
        t = ptx::unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = silu_f(fmaf(v[j], mul[j], add[j])) + post[j];
        if (rin != nullptr) {
            const uint4 r = __ldg(rin + i);
            t = ptx::unpack_bf16x2(r.x); v[0] += t.x; v[1] += t.y;
            t = ptx::unpack_bf16x2(r.y); v[2] += t.x; v[3] += t.y;
            t = ptx::unpack_bf16x2(r.z); v[4] += t.x; v[5] += t.y;
            t = ptx::unpack_bf16x2(r.w); v[6] += t.x; v[7] += t.y;
        }
        uint4 o;
        o.x = ptx::pack_bf16x2(v[0], v[1]);
        o.y = ptx::pack_bf16x2(v[2], v[3]);
        o.z = ptx::pack_bf16x2(v[4], v[5]);
        o.w = ptx::pack_bf16x2(v[6], v[7]);
        yout[i] = o;
    }
}

// ----------------------------------------------------------------------------------------------- channel LN
template <int C>
__global__ void __launch_bounds__(256)
channel_layernorm_kernel(const LayerNormArgs a) {
    constexpr int VEC = C / 32;            // elements per lane (2, 4, 8, 16)
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int m = warp_global;
    if (m >= a.M) return;
    const bf16* xp = a.x + static_cast<size_t>(m) * C + lane * VEC;
    float v[VEC];
    if constexpr (VEC == 2) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(xp));
        const float2 t = ptx::unpack_bf16x2(u); v[0] = t.x; v[1] = t.y;
    } else if constexpr (VEC == 4) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(xp));
        float2 t = ptx::unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
        t = ptx::unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
    } else {
#pragma unroll
        for (int q = 0; q < VEC / 8; ++q) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(xp) + q);
            float2 t = ptx::unpack_bf16x2(u.x); v[q * 8 + 0] = t.x; v[q * 8 + 1] = t.y;
            t = ptx::unpack_bf16x2(u.y); v[q * 8 + 2] = t.x; v[q * 8 + 3] = t.y;
            t = ptx::unpack_bf16x2(u.z); v[q * 8 + 4] = t.x; v[q * 8 + 5] = t.y;
            t = ptx::unpack_bf16x2(u.w); v[q * 8 + 6] = t.x; v[q * 8 + 7] = t.y;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) s += v[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) ss += (v[j] - mean) * (v[j] - mean);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float rstd = rsqrtf(ss * (1.0f / C) + a.eps);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = (v[j] - mean) * rstd * __ldg(a.g + lane * VEC + j);
    if (a.res != nullptr) {
        const bf16* rp = a.res + static_cast<size_t>(m) * C + lane * VEC;
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            const float2 t = ptx::unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(rp + j)));
            v[j] += t.x; v[j + 1] += t.y;
        }
    }
    uint32_t o[VEC / 2];
#pragma unroll
    for (int j = 0; j < VEC; j += 2) o[j / 2] = ptx::pack_bf16x2(v[j], v[j + 1]);

    auto store_row = [&](size_t row) {
        bf16* yp = a.y + row * C + lane * VEC;
        if constexpr (VEC == 2) {
            *reinterpret_cast<uint32_t*>(yp) = o[0];
        } else if constexpr (VEC == 4) {
            *reinterpret_cast<uint2*>(yp) = make_uint2(o[0], o[1]);
        } else {
#pragma unroll
            for (int q = 0; q < VEC / 8; ++q)
                reinterpret_cast<uint4*>(yp)[q] = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
        }
    };
    if (!a.upsample2x) {
        store_row(static_cast<size_t>(m));
    } else {
        const int P = a.H * a.W;
        const int b = m / P;
        const int hw = m - b * P;
        const int h = hw / a.W;
        const int w = hw - h * a.W;
        const size_t W2 = 2 * a.W;
        const size_t base = (static_cast<size_t>(b) * 2 * a.H + 2 * h) * W2 + 2 * w;
        store_row(base);
        store_row(base + 1);
        store_row(base + W2);
        store_row(base + W2 + 1);
    }
}

}  // namespace

cudaError_t groupnorm_film_silu_run(const GroupNormArgs& a, cudaStream_t s) {
    if (a.C % 64 != 0 || a.C > 512 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
    const size_t img_bytes = static_cast<size_t>(a.P) * a.C * 2;
    int S = 1;
    while (img_bytes / S > GN_MAX_SLAB) S *= 2;
    if (S > 8 || (img_bytes / 16) % S != 0 || a.P % S != 0) return cudaErrorInvalidValue;
    GNK k;
    k.a = a;
    k.S = S;
    k.slab_chunks = static_cast<int>(img_bytes / 16 / S);
    const int smem = k.slab_chunks * 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(groupnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_MAX_SLAB);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.B * S);
    cfg.blockDim = dim3(GN_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, groupnorm_kernel, k);
}

cudaError_t channel_layernorm_run(const LayerNormArgs& a, cudaStream_t s) {
    const int threads = 256;
    const int rows_per_block = threads / 32;
    const int grid = (a.M + rows_per_block - 1) / rows_per_block;
    switch (a.C) {
        case 64: channel_layernorm_kernel<64><<<grid, threads, 0, s>>>(a); break;
        case 128: channel_layernorm_kernel<128><<<grid, threads, 0, s>>>(a); break;
        case 256: channel_layernorm_kernel<256><<<grid, threads, 0, s>>>(a); break;
        case 512: channel_layernorm_kernel<512><<<grid, threads, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace hd
