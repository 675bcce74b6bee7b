// Data preparation on the GPU: the step BEFORE the sampling / training path.
//
// reference: loadBothConstraints  /root/reference/processdata/PrepareData_linear.py:48-103
//              :55-76   (row, col, value) triples -> dense symmetric matrix, a Python loop over the non-zeros, later entries
//                       overwrite earlier ones
//              :77-85   bins whose diagonal is 0 or NaN are deleted (rows and columns)
//              :88-92   per = np.percentile(mat, 99.0); mat = 2 * clip(mat, 0, per) / per - 1
//            split_numpy :183-213  tiles (splitPieces -> tiles.cu) and y = x + sigma_0 * randn for the 'deno' operator
// Everything here is indexing, order statistics or IEEE fp32 arithmetic in the reference's order, so results are BIT-EXACT:
//   * the scatter resolves duplicate cells like the sequential loop (highest entry index wins) through an atomicMax owner map;
//   * the percentile's two order statistics come from an exact 3-pass radix select on the monotone integer image of the
//     floats (integer histograms: no floating-point reduction order anywhere); numpy's interpolation of the two runs on the host.
#include "kernels.h"

namespace hd {
namespace {

// ---------------------------------------------------------------------------------------------- COO -> dense
// entry i writes (r, c) then (c, r); the surviving writer of a cell is the highest i (the mirrored write of entry i comes after
// its direct write, which only matters on the diagonal where both carry the same value)
__global__ void coo_owner_kernel(const long long* __restrict__ rows, const long long* __restrict__ cols, long long nnz,
                                 long long smallbin, long long n, int* __restrict__ owner, int* __restrict__ bad) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= nnz) return;
    const long long r = rows[i] - smallbin, c = cols[i] - smallbin;
    if (r < 0 || c < 0 || r >= n || c >= n) { atomicExch(bad, 1); return; }
    atomicMax(owner + r * n + c, static_cast<int>(i));
    atomicMax(owner + c * n + r, static_cast<int>(i));
}
__global__ void coo_fill_kernel(const float* __restrict__ vals, const int* __restrict__ owner, long long cells, float* __restrict__ mat) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= cells) return;
    const int o = owner[i];
    mat[i] = o >= 0 ? vals[o] : 0.0f;
}

// ---------------------------------------------------------------------------------------------- empty-bin removal
// one block: keep[i] = !(diag == 0 || isnan(diag)); map[j] = j-th kept index; *n_kept
__global__ void __launch_bounds__(1024)
keep_map_kernel(const float* __restrict__ mat, long long n, long long* __restrict__ map, long long* __restrict__ n_kept) {
    __shared__ int s_cnt[1024];
    const int tid = threadIdx.x;
    const long long per = (n + 1023) / 1024;
    const long long i0 = tid * per, i1 = i0 + per < n ? i0 + per : n;
    int cnt = 0;
    for (long long i = i0; i < i1; ++i) {
        const float d = mat[i * n + i];
        cnt += !(d == 0.0f || d != d);
    }
    s_cnt[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < 1024; ++k) { const int c = s_cnt[k]; s_cnt[k] = run; run += c; }
        *n_kept = run;
    }
    __syncthreads();
    long long pos = s_cnt[tid];
    for (long long i = i0; i < i1; ++i) {
        const float d = mat[i * n + i];
        if (!(d == 0.0f || d != d)) map[pos++] = i;
    }
}
__global__ void compact_kernel(const float* __restrict__ mat, long long n, const long long* __restrict__ map, long long m,
                               float* __restrict__ out) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= m * m) return;
    const long long r = i / m, c = i - r * m;
    out[i] = mat[map[r] * n + map[c]];
}

// ---------------------------------------------------------------------------------------------- exact order statistic
__device__ __forceinline__ unsigned int key_of(float v) {
    const unsigned int u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // monotone: a < b  <=>  key(a) < key(b)   (-0 < +0, NaNs at the ends)
}
struct SelectState {        // device-resident between the passes
    unsigned long long rank;   // rank still to be located inside the current prefix
    unsigned int prefix;       // key bits fixed so far (left-aligned)
    unsigned int hist[2048];
};
// pass p (0, 1, 2) looks at bits [21,32), [10,21), [0,10) of the keys whose higher bits equal the prefix
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ x, long long n, int pass, SelectState* __restrict__ st) {
    __shared__ unsigned int s_h[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s_h[i] = 0;
    __syncthreads();
    const unsigned int prefix = st->prefix;
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const unsigned int mask_hi = pass == 0 ? 0u : (pass == 1 ? 0xFFE00000u : 0xFFFFFC00u);
    const unsigned int nb = pass == 2 ? 1023u : 2047u;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const unsigned int k = key_of(x[i]);
        if ((k & mask_hi) == prefix) atomicAdd(&s_h[(k >> shift) & nb], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += 256)
        if (s_h[i]) atomicAdd(&st->hist[i], s_h[i]);
}
__global__ void select_scan_kernel(int pass, SelectState* st, float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const int bins = pass == 2 ? 1024 : 2048;
    unsigned long long r = st->rank;
    int b = 0;
    for (; b < bins - 1; ++b) {
        const unsigned int h = st->hist[b];
        if (r < h) break;
        r -= h;
    }
    st->rank = r;
    st->prefix |= static_cast<unsigned int>(b) << shift;
    for (int i = 0; i < 2048; ++i) st->hist[i] = 0;
    if (pass == 2) {
        const unsigned int k = st->prefix;
        const unsigned int u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
        *out = __uint_as_float(u);
    }
}
__global__ void select_init_kernel(SelectState* st, unsigned long long rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2048) st->hist[i] = 0;
    if (i == 0) { st->rank = rank; st->prefix = 0; }
}

// ---------------------------------------------------------------------------------------------- normalisation, noise
__global__ void normalize_contacts_kernel(float* __restrict__ x, long long n, float per) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    const float c = fminf(fmaxf(x[i], 0.0f), per);            // np.clip(mat, 0, per)
    const float s = __fdiv_rn(c, per);                        // mat / per
    x[i] = __fsub_rn(__fmul_rn(2.0f, s), 1.0f);               // 2 * mat - 1.0
}
__global__ void axpy_noise_kernel(const float* __restrict__ x, const float* __restrict__ z, float sigma, long long n, float* __restrict__ y) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i < n) y[i] = __fadd_rn(x[i], __fmul_rn(sigma, z[i]));   // data + sigma_0 * randn_like(data): two roundings, like torch
}

inline int blocks_for(long long n) { return static_cast<int>((n + 255) / 256); }

}  // namespace

size_t coo_scratch_bytes(long long n) { return static_cast<size_t>(n) * n * sizeof(int) + 16; }

cudaError_t coo_to_dense_run(const long long* rows, const long long* cols, const float* vals, long long nnz, long long smallbin,
                             long long n, float* mat, void* scratch, int* bad_host, cudaStream_t s) {
    int* owner = static_cast<int*>(scratch);
    int* bad = owner + n * n;
    cudaError_t e = cudaMemsetAsync(owner, 0xFF, static_cast<size_t>(n) * n * sizeof(int), s);   // -1
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(bad, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (nnz > 0) coo_owner_kernel<<<blocks_for(nnz), 256, 0, s>>>(rows, cols, nnz, smallbin, n, owner, bad);
    if (n > 0) coo_fill_kernel<<<blocks_for(n * n), 256, 0, s>>>(vals, owner, n * n, mat);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(bad_host, bad, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}

cudaError_t keep_map_run(const float* mat, long long n, long long* map, long long* n_kept_dev, cudaStream_t s) {
    keep_map_kernel<<<1, 1024, 0, s>>>(mat, n, map, n_kept_dev);
    return cudaGetLastError();
}
cudaError_t compact_run(const float* mat, long long n, const long long* map, long long m, float* out, cudaStream_t s) {
    if (m > 0) compact_kernel<<<blocks_for(m * m), 256, 0, s>>>(mat, n, map, m, out);
    return cudaGetLastError();
}

size_t select_scratch_bytes() { return sizeof(SelectState); }
// out[0] = the element of rank `rank` (0-based, ascending) of x[0..n)
cudaError_t select_rank_run(const float* x, long long n, unsigned long long rank, float* out, void* scratch, cudaStream_t s) {
    SelectState* st = static_cast<SelectState*>(scratch);
    select_init_kernel<<<8, 256, 0, s>>>(st, rank);
    const int grid = static_cast<int>(n / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    for (int pass = 0; pass < 3; ++pass) {
        select_hist_kernel<<<grid, 256, 0, s>>>(x, n, pass, st);
        select_scan_kernel<<<1, 32, 0, s>>>(pass, st, out);
    }
    return cudaGetLastError();
}

cudaError_t normalize_contacts_run(float* x, long long n, float per, cudaStream_t s) {
    if (n > 0) normalize_contacts_kernel<<<blocks_for(n), 256, 0, s>>>(x, n, per);
    return cudaGetLastError();
}
cudaError_t axpy_noise_run(const float* x, const float* z, float sigma, long long n, float* y, cudaStream_t s) {
    if (n > 0) axpy_noise_kernel<<<blocks_for(n), 256, 0, s>>>(x, z, sigma, n, y);
    return cudaGetLastError();
}

}  // namespace hd
