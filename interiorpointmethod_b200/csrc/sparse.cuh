// Sparse kernels for the Netlib path: CSR mat-vecs and the numeric phase of the SpGEMM
// M = A diag(x/s) A^T (main.py:223-224) on a pattern fixed once per problem.
#pragma once
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace ipm {

// Symbolic product pattern, built on the host once per problem (ipm_load_csr):
// entry e of the lower triangle of M lives at linear index out_idx[e] of the dense row-major M and is
//   sum_{t in [prod_ptr[e], prod_ptr[e+1])}  (val[pa[t]] * d[colind[pa[t]]]) * val[pb[t]]
// with the terms ordered by the shared column index k, i.e. the same (A D) A^T association and summation
// order scipy's csr_matmat uses for main.py:224.
struct SpgemmPattern {
    std::vector<int64_t> out_idx;
    std::vector<int64_t> prod_ptr;
    std::vector<int32_t> pa, pb;
};

// rows tiled in row order: entry list is sorted by (i, j), so consecutive threads write neighbouring M entries.
inline void spgemm_symbolic(int m, int n, const int32_t* rowptr, const int32_t* colind, int64_t ldm,
                            SpgemmPattern& pat) {
    const int64_t nnz = rowptr[m];
    // CSC view: for each column the (row, position) pairs in row order
    std::vector<int64_t> cptr(n + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) cptr[colind[p] + 1]++;
    for (int k = 0; k < n; ++k) cptr[k + 1] += cptr[k];
    std::vector<int32_t> crow(nnz), cpos(nnz);
    {
        std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
        for (int i = 0; i < m; ++i)
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                const int64_t q = fill[colind[p]]++;
                crow[q] = i;
                cpos[q] = (int32_t)p;
            }
    }
    pat.out_idx.clear(); pat.prod_ptr.clear(); pat.pa.clear(); pat.pb.clear();
    pat.prod_ptr.push_back(0);
    std::vector<int32_t> mark(m, -1), cnt(m, 0), uniq;
    std::vector<int64_t> start(m, 0);
    for (int i = 0; i < m; ++i) {
        uniq.clear();
        // pass 1: which j <= i appear, and how many terms each has
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const int k = colind[p];
            for (int64_t q = cptr[k]; q < cptr[k + 1]; ++q) {
                const int j = crow[q];
                if (j > i) break;                      // rows inside a column are ascending
                if (mark[j] != i) { mark[j] = i; cnt[j] = 0; uniq.push_back(j); }
                cnt[j]++;
            }
        }
        std::sort(uniq.begin(), uniq.end());
        int64_t base = (int64_t)pat.pa.size();
        for (int j : uniq) {
            start[j] = base;
            base += cnt[j];
            pat.out_idx.push_back((int64_t)i * ldm + j);
            pat.prod_ptr.push_back(base);
        }
        pat.pa.resize(base);
        pat.pb.resize(base);
        // pass 2: fill terms; p ascends => k ascends inside every (i, j) list when colind is sorted
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const int k = colind[p];
            for (int64_t q = cptr[k]; q < cptr[k + 1]; ++q) {
                const int j = crow[q];
                if (j > i) break;
                const int64_t t = start[j]++;
                pat.pa[t] = (int32_t)p;
                pat.pb[t] = cpos[q];
            }
        }
    }
}

#ifdef __CUDACC__
// y = A v for CSR A, 8 lanes per row (Netlib rows are short), fixed shuffle tree.
static __device__ __forceinline__ void d_spmv_csr(int rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                           const double* __restrict__ val, const double* v, double* y) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t rw = (gt >> 5) * 4; rw < rows; rw += nwarps * 4) {      // warp-uniform trip count
        const int64_t r = rw + (lane >> 3);
        double acc = 0.0;
        if (r < rows) {
            const int p1 = rowptr[r + 1];
            for (int p = rowptr[r] + sub; p < p1; p += 8) acc += val[p] * v[colind[p]];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (sub == 0 && r < rows) y[r] = acc;
    }
}
static __global__ void k_spmv_csr(int rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                           const double* __restrict__ val, const double* __restrict__ v, double* __restrict__ y) { d_spmv_csr(rows, rowptr, colind, val, v, y); }

// ad[p] = val[p] * d[colind[p]]   (the A @ D_square factor of main.py:224)
static __device__ __forceinline__ void d_scale_vals(int64_t nnz, const int32_t* __restrict__ colind, const double* __restrict__ val,
                             const double* d, double* ad) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
        ad[p] = val[p] * d[colind[p]];
}
static __global__ void k_scale_vals(int64_t nnz, const int32_t* __restrict__ colind, const double* __restrict__ val,
                             const double* __restrict__ d, double* __restrict__ ad) { d_scale_vals(nnz, colind, val, d, ad); }

// numeric SpGEMM: one thread per lower-triangle entry of M, terms summed in pattern order (deterministic).
static __device__ __forceinline__ void d_spgemm_numeric(int64_t nent, const int64_t* __restrict__ out_idx,
                                 const int64_t* __restrict__ prod_ptr, const int32_t* __restrict__ pa,
                                 const int32_t* __restrict__ pb, const double* ad,
                                 const double* __restrict__ val, double* M) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nent; e += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        const int64_t t1 = prod_ptr[e + 1];
        for (int64_t t = prod_ptr[e]; t < t1; ++t) acc += ad[pa[t]] * val[pb[t]];
        M[out_idx[e]] = acc;
    }
}
static __global__ void k_spgemm_numeric(int64_t nent, const int64_t* __restrict__ out_idx,
                                 const int64_t* __restrict__ prod_ptr, const int32_t* __restrict__ pa,
                                 const int32_t* __restrict__ pb, const double* __restrict__ ad,
                                 const double* __restrict__ val, double* __restrict__ M) { d_spgemm_numeric(nent, out_idx, prod_ptr, pa, pb, ad, val, M); }
#endif

}  // namespace ipm
