// Stand-alone tcgen05.mma issue-rate microbenchmark: separates a hardware operand-fetch floor from a
// pipeline bubble of conv_gemm_kernel.  No TMA, no epilogue: operands are RESIDENT in shared memory (or TMEM), one elected thread
// issues a long run of MMAs into one accumulator, one commit, one mbarrier wait; clock64() around the run.
//
//   forms   SS (A and B from shared memory)  |  TS (A from TMEM, B from shared memory)
//   shapes  M = 128 or 64 per CTA, N = 64 / 128 / 256, K = 16 (kind::f16, bf16 operands, fp32 accumulate)
//   pairs   cta_group::1  |  cta_group::2 (cluster of 2: M = 256 or 128 over the pair, each CTA holds N/2 rows of B)
//
// Prints one JSON line per configuration: cycles per MMA (median / min / max over the CTAs of the grid) and the implied
// shared-memory operand bytes per cycle.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench_tcgen05
// ubench_tcgen05.cu -I../hicdiff_b200/csrc.  Run on a B200: ./ubench_tcgen05 [iters]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace hd;

namespace {

constexpr int THREADS = 128;
constexpr uint32_t A_BYTES = 128 * 128;        // 128 rows x 64 bf16 (one 128-byte swizzle span per row)
constexpr uint32_t B_BYTES = 256 * 128;        // up to 256 rows
constexpr int NTILES = 3;                      // distinct operand tiles the run cycles through (rotate = 1)
constexpr uint32_t SMEM_BYTES = NTILES * (A_BYTES + B_BYTES) + 1024 + 64;

struct Params {
    int M;        // rows per CTA: 128 or 64
    int N;        // accumulator columns (over the pair for cta_group::2)
    int ts;       // 1: A operand from TMEM
    int rotate;   // 1: cycle through K slices and operand tiles like a real K loop; 0: the same operands every time
    int iters;    // MMAs in the timed run
    long long* cycles;   // [grid] (leader CTAs write theirs; peers write 0)
};

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(1u)
        : "memory");
}
__device__ __forceinline__ void umma_ts_cg2(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(1u)
        : "memory");
}

template <int CG>
__global__ void __launch_bounds__(THREADS, 1) ubench_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + NTILES * A_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NTILES * (A_BYTES + B_BYTES));
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;

    // operands: small finite bf16 values (0x3c00.. ~ 0.0078 .. ), different per element so the datapath toggles
    for (uint32_t i = threadIdx.x; i < NTILES * (A_BYTES + B_BYTES) / 4; i += THREADS) {
        const uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u | (h & 0x007f007fu);
    }
    ptx::fence_proxy_async_smem();
    if (warp == 0) {
        if (CG == 2) { ptx::tmem_alloc_cg2(slot, 512); ptx::tmem_relinquish_cg2(); }
        else { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
    } else if (warp == 1 && (threadIdx.x & 31) == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
    }
    ptx::tc_fence_before();
    if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *slot;

    long long cyc = 0;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = ptx::make_idesc_bf16(static_cast<uint32_t>(p.M * CG), static_cast<uint32_t>(p.N));
        const uint64_t dA = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA));
        const uint64_t dB = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB));
        const uint32_t a_tmem0 = tmem + 256;           // TS: A tiles start at column 256 (8 columns per K = 16 slice)
        if (ptx::elect_one()) {
            // warm-up run (not timed): 64 MMAs + commit + wait
            for (int rep = 0; rep < 2; ++rep) {
                const int n = rep == 0 ? 64 : p.iters;
                const long long t0 = clock64();
                for (int i = 0; i < n; ++i) {
                    const int k = p.rotate ? (i & 3) : 0;
                    const int tile = p.rotate ? ((i >> 2) % NTILES) : 0;
                    const uint64_t db = dB + static_cast<uint64_t>((tile * B_BYTES) >> 4) + 2u * k;
                    if (p.ts) {
                        const uint32_t at = a_tmem0 + static_cast<uint32_t>((tile * 4 + k) * 8);
                        if (CG == 2) umma_ts_cg2(tmem, at, db, idesc); else umma_ts(tmem, at, db, idesc);
                    } else {
                        const uint64_t da = dA + static_cast<uint64_t>((tile * A_BYTES) >> 4) + 2u * k;
                        if (CG == 2) ptx::umma_bf16_cg2(tmem, da, db, idesc, 1u); else ptx::umma_bf16(tmem, da, db, idesc, 1u);
                    }
                }
                if (CG == 2) ptx::umma_commit_cg2(bar, 3); else ptx::umma_commit(bar);
                ptx::mbar_wait(bar, static_cast<uint32_t>(rep & 1));
                cyc = clock64() - t0;
            }
        }
        __syncwarp();
    } else if (CG == 2 && warp == 1 && rank == 1) {
        // the peer's barrier receives the multicast commits too: consume them so the phases stay aligned
        if (ptx::elect_one()) {
            ptx::mbar_wait(bar, 0);
            ptx::mbar_wait(bar, 1);
        }
        __syncwarp();
    }
    ptx::tc_fence_before();
    if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (threadIdx.x == 32) p.cycles[blockIdx.x] = (rank == 0) ? cyc : 0;
    if (warp == 0) {
        if (CG == 2) ptx::tmem_dealloc_cg2(tmem, 512); else ptx::tmem_dealloc(tmem, 512);
    }
}

template <int CG>
cudaError_t launch(const Params& p, int grid) {
    cudaError_t e = cudaFuncSetAttribute(ubench_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ubench_kernel<CG>, p);
}

}  // namespace

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4096;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess || prop.major != 10) {
        fprintf(stderr, "needs an sm_100 GPU\n");
        return 2;
    }
    const int sms = prop.multiProcessorCount;
    long long* d_cyc = nullptr;
    cudaMalloc(&d_cyc, sizeof(long long) * sms);
    std::vector<long long> h(sms);
    printf("{\"device\": \"%s\", \"sms\": %d, \"iters\": %d}\n", prop.name, sms, iters);
    for (int grid_mode = 0; grid_mode < 2; ++grid_mode)          // 0: one CTA (pair) alone, 1: every SM busy
        for (int cg = 1; cg <= 2; ++cg)
            for (int ts = 0; ts < 2; ++ts)
                for (int M : {128, 64})
                    for (int N : {64, 128, 256})
                        for (int rotate = 1; rotate >= 0; --rotate) {
                            if (grid_mode == 0 && rotate == 0) continue;
                            Params p{M, N, ts, rotate, iters, d_cyc};
                            const int grid = grid_mode == 0 ? cg : (sms / cg) * cg;
                            cudaMemset(d_cyc, 0, sizeof(long long) * sms);
                            cudaError_t e = cg == 1 ? launch<1>(p, grid) : launch<2>(p, grid);
                            if (e == cudaSuccess) e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) {
                                printf("{\"cg\": %d, \"ts\": %d, \"M\": %d, \"N\": %d, \"rotate\": %d, \"grid\": %d, \"error\": \"%s\"}\n", cg, ts,
                                       M, N, rotate, grid, cudaGetErrorString(e));
                                return 1;
                            }
                            cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
                            std::vector<double> v;
                            for (int i = 0; i < grid; ++i)
                                if (h[i] > 0) v.push_back(static_cast<double>(h[i]) / iters);
                            std::sort(v.begin(), v.end());
                            const double med = v[v.size() / 2];
                            // shared-memory operand bytes one CTA's tensor core reads per MMA (K = 16 bf16 = 32 B per row)
                            const double bytes = (ts ? 0.0 : M * 32.0) + (N / cg) * 32.0;
                            const double floor_clk = N / 2.0;      // guide: max(M_atom,128) * N / (256 * cta_group) with M_atom = 128 * cta_group
                            printf("{\"cg\": %d, \"form\": \"%s\", \"M_per_cta\": %d, \"N\": %d, \"rotate\": %d, \"ctas\": %d, "
                                   "\"clk_per_mma_median\": %.1f, \"min\": %.1f, \"max\": %.1f, \"smem_operand_bytes\": %.0f, "
                                   "\"smem_bytes_per_clk\": %.1f, \"math_floor_clk\": %.0f, \"tensor_rate_frac\": %.3f}\n",
                                   cg, ts ? "TS" : "SS", M, N, rotate, grid, med, v.front(), v.back(), bytes, bytes / med,
                                   floor_clk, (M / 128.0) * floor_clk / med);
                            fflush(stdout);
                        }
    cudaFree(d_cyc);
    return 0;
}
