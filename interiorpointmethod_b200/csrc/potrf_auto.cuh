// Picks the Cholesky variant for ONE matrix by its order.
#pragma once
#include "chol.cuh"
#include "chol_batched.cuh"

namespace ipm {
#ifdef __CUDACC__
//   m <= 512  : the fused one-CTA kernel of the batched solver (a single launch; 512 threads above 256)
//   m <= 2048 : blocked, 64-wide panels (the panel kernels are latency-bound: narrower panels shorten the chain)
//   larger    : blocked, 128-wide panels (the trailing update dominates: wider panels raise its intensity)
inline int potrf_single_auto(double* M, int64_t ldm, int m, double* scal, double tau, cudaStream_t st,
                             DepMask dm = DepMask()) {
    if (m <= KBC_MAX_M_BIG) return potrf_batched_fused(M, ldm, 0, m, 1, scal, 0, tau, nullptr, st, dm.mask, dm.mode);
    if (m <= 2048) return potrf_blocked<64, 256, 128>(M, ldm, 0, m, 1, scal, 0, tau, nullptr, st, dm);
    return potrf_blocked<128, 512, 64>(M, ldm, 0, m, 1, scal, 0, tau, nullptr, st, dm);
}
#endif
}  // namespace ipm
