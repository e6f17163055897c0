// Dense row-major mat-vecs for the dense single-LP path (the `interior` caller, main.py:707-757):
//   gemv_rows : y = A v      one warp per row, 128-bit loads, fixed shuffle tree
//   gemv_cols : w = A^T u    one thread per column over a chunk of rows; chunks combined in index order
// HBM-bound: each reads A exactly once (8 m n bytes).
#pragma once
#include "common.cuh"

namespace ipm {

constexpr int GEMV_NT = 256;
constexpr int GEMVC_ROWS = 256;      // rows per chunk in gemv_cols

#ifdef __CUDACC__
static __global__ void __launch_bounds__(GEMV_NT) k_gemv_rows(int m, int n, const double* __restrict__ A, int64_t lda,
                                                       const double* __restrict__ v, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    for (int64_t r = warp; r < m; r += nwarps) {
        const double* row = A + (size_t)r * lda;
        double acc0 = 0.0, acc1 = 0.0;
        if (vec) {
            const int n2 = n >> 1;
            const double2* row2 = reinterpret_cast<const double2*>(row);
            const double2* v2 = reinterpret_cast<const double2*>(v);
#pragma unroll 4
            for (int k = lane; k < n2; k += 32) {
                const double2 a = row2[k], b = v2[k];
                acc0 += a.x * b.x;
                acc1 += a.y * b.y;
            }
            if ((n & 1) && lane == 0) acc0 += row[n - 1] * v[n - 1];
        } else {
            for (int k = lane; k < n; k += 32) acc0 += row[k] * v[k];
        }
        const double acc = warp_sum(acc0 + acc1);
        if (lane == 0) y[r] = acc;
    }
}

// partial[chunk][k] = sum_{i in chunk} A[i][k] u[i];  grid (ceil(n/GEMV_NT), nchunks)
static __global__ void __launch_bounds__(GEMV_NT) k_gemv_cols_partial(int m, int n, const double* __restrict__ A, int64_t lda,
                                                               const double* __restrict__ u,
                                                               double* __restrict__ partial) {
    __shared__ double us[GEMVC_ROWS];
    const int i0 = blockIdx.y * GEMVC_ROWS;
    const int rows = (m - i0 < GEMVC_ROWS) ? (m - i0) : GEMVC_ROWS;
    for (int i = threadIdx.x; i < rows; i += GEMV_NT) us[i] = u[i0 + i];
    __syncthreads();
    const int k = blockIdx.x * GEMV_NT + threadIdx.x;
    if (k >= n) return;
    const double* col = A + (size_t)i0 * lda + k;
    double acc = 0.0;
#pragma unroll 8
    for (int i = 0; i < rows; ++i) acc += col[(size_t)i * lda] * us[i];
    partial[(size_t)blockIdx.y * n + k] = acc;
}
static __global__ void k_gemv_cols_combine(int n, int nchunks, const double* __restrict__ partial, double* __restrict__ w) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double acc = 0.0;
    for (int c = 0; c < nchunks; ++c) acc += partial[(size_t)c * n + k];
    w[k] = acc;
}
#endif

}  // namespace ipm
