// Tile extraction / reassembly in the enumeration order of the reference's splitPieces
// (/root/reference/processdata/PrepareData_linear.py:25-46):
//   pad n -> n' = ceil(n / piece) * piece with zeros (bottom / right);
//   for i in range(0, n', piece): for j in range(i, n', piece): if |i - j| <= piece*4*scal + 1: take [i:i+p, j:j+p]
// With step == piece the band test is (bj - bi) <= band_blocks, band_blocks = 4 * int(40000 / res).
// Pure index arithmetic on fp32 payloads: results are bit-exact by construction.
#include "kernels.h"

namespace hd {
namespace {

__device__ __forceinline__ void tile_to_block(int k, int P, int band, int* bi, int* bj) {
    int row = 0, before = 0;
    for (;;) {
        const int cnt = min(band + 1, P - row);
        if (k < before + cnt) break;
        before += cnt;
        ++row;
    }
    *bi = row;
    *bj = row + (k - before);
}

__global__ void __launch_bounds__(256)
tile_extract_kernel(const float* __restrict__ mat, int n, float* __restrict__ tiles, int piece, int band, int P) {
    int bi, bj;
    tile_to_block(blockIdx.x, P, band, &bi, &bj);
    float* t = tiles + static_cast<size_t>(blockIdx.x) * piece * piece;
    for (int i = threadIdx.x; i < piece * piece; i += blockDim.x) {
        const int y = i / piece, x = i - y * piece;
        const int gy = bi * piece + y, gx = bj * piece + x;
        t[i] = (gy < n && gx < n) ? mat[static_cast<size_t>(gy) * n + gx] : 0.0f;
    }
}

__global__ void __launch_bounds__(256)
tile_scatter_kernel(const float* __restrict__ tiles, float* __restrict__ mat, int n, int piece, int band, int P) {
    int bi, bj;
    tile_to_block(blockIdx.x, P, band, &bi, &bj);
    const float* t = tiles + static_cast<size_t>(blockIdx.x) * piece * piece;
    for (int i = threadIdx.x; i < piece * piece; i += blockDim.x) {
        const int y = i / piece, x = i - y * piece;
        const int gy = bi * piece + y, gx = bj * piece + x;
        if (gy < n && gx < n) {
            const float v = t[i];
            mat[static_cast<size_t>(gy) * n + gx] = v;
            if (bi != bj) mat[static_cast<size_t>(gx) * n + gy] = v;   // symmetric counterpart
        }
    }
}

}  // namespace

int tile_count(int n, int piece, int band) {
    const int P = (n + piece - 1) / piece;
    int cnt = 0;
    for (int r = 0; r < P; ++r) cnt += (band + 1 < P - r) ? band + 1 : P - r;
    return cnt;
}

cudaError_t tile_extract_run(const float* mat, int n, float* tiles, int piece, int band, cudaStream_t s) {
    const int P = (n + piece - 1) / piece;
    const int cnt = tile_count(n, piece, band);
    if (cnt == 0) return cudaSuccess;
    tile_extract_kernel<<<cnt, 256, 0, s>>>(mat, n, tiles, piece, band, P);
    return cudaGetLastError();
}

cudaError_t tile_scatter_run(const float* tiles, float* mat, int n, int piece, int band, cudaStream_t s) {
    const int P = (n + piece - 1) / piece;
    const int cnt = tile_count(n, piece, band);
    if (cnt == 0) return cudaSuccess;
    tile_scatter_kernel<<<cnt, 256, 0, s>>>(tiles, mat, n, piece, band, P);
    return cudaGetLastError();
}

}  // namespace hd
