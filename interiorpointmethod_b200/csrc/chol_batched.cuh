// Fused per-LP Cholesky for the batched workload: ONE CTA factors one m x m matrix (m <= KBC_MAX_M) end to
// end, left-looking in 32-wide panels, so the whole factorisation of all LPs is a single launch and M makes one
// trip through HBM (read) and one back (L).  Replaces solve_linear (main.py:176-182) on M (main.py:223-224).
//
// Panel J (columns j0..j0+31, rows j0..m-1):
//   1. update   acc = M[:, J] - sum_{K<J} L[:, K] L[J, K]^T     DMMA.8x8x4; A/B fragments are read straight
//                                                                from global memory (L written by this CTA a
//                                                                moment ago -> L1/L2 hits), accumulators = panel
//   2. diag     32x32 block: warp 0, lane i owns row i in registers, finished rows are broadcast through
//               shared memory; pivot safeguard p <= tau*maxdiag or NaN -> 1e128;
//               the pivot itself travels by warp shuffle
//   3. trsm     rows below: one thread per row, forward substitution against the diagonal block
//   4. store    panel -> global L
#pragma once
#include "common.cuh"

namespace ipm {

constexpr int KBC_NT = 256;
constexpr int KBC_NW = KBC_NT / 32;
constexpr int KBC_MAX_M = 256;      // 4 row tiles x 4 column tiles of accumulators per warp
constexpr int KBC_LD = 33;
constexpr int KBC_LDT = 34;     // transposed diagonal block: even so that 128-bit reads stay aligned
constexpr int KBC_KC = 64;      // columns of L[J, :] staged in shared memory per chunk of the update
constexpr int KBC_LDB = KBC_KC + 4;   // 68 = 4 mod 16: conflict-free 64-bit fragment reads

struct CholBatchedArgs {
    double* M; int64_t ldm; int64_t strideM;
    double* scal; int64_t strideScal;     // S_MAXDIAG, S_NFIXED written per LP (nullable)
    double tau;
    int m;
    const int* active;
};

inline size_t kbc_smem_bytes(int m) {
    const int below = m > 32 ? m - 32 : 0;
    return (size_t)(32 * KBC_LD + 2 + 32 * KBC_LDT + 32 * KBC_LDB + (size_t)below * KBC_LD + 32) * sizeof(double);
}

#ifdef __CUDACC__
// One chunk of the left-looking update for the NTI row tiles of a warp: A fragments straight from global memory
// (L written by this CTA a moment ago) with a one-step register prefetch, B fragments from the staged chunk.
template <int NTI>
__device__ __forceinline__ void kbc_update_chunk(double (&acc)[4][4][2], const double* __restrict__ Mb,
                                                 const int (&aoff)[4], const bool (&aok)[4], const double* Bs,
                                                 int k0, int kc, int g, int t) {
    double af[NTI], afn[NTI];
#pragma unroll
    for (int i = 0; i < NTI; ++i) af[i] = aok[i] ? Mb[aoff[i] + k0] : 0.0;
    const double* bs = Bs + g * KBC_LDB + t;
#pragma unroll 2
    for (int kk = 0; kk < kc; kk += 4) {
        if (kk + 4 < kc) {
#pragma unroll
            for (int i = 0; i < NTI; ++i) afn[i] = aok[i] ? Mb[aoff[i] + k0 + kk + 4] : 0.0;
        }
        double bf[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = bs[ni * 8 * KBC_LDB + kk];
#pragma unroll
        for (int ti = 0; ti < NTI; ++ti)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[ti][ni][0], acc[ti][ni][1], af[ti], bf[ni]);
#pragma unroll
        for (int i = 0; i < NTI; ++i) af[i] = afn[i];
    }
}

static __global__ void __launch_bounds__(KBC_NT, 2) kb_chol(const CholBatchedArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* D = smem;                       // [32][33]  diagonal block, becomes L_JJ
    double* DT = smem + 32 * KBC_LD + 2;    // [32][34]  DT[k][j] = L_JJ[j][k]  (+2 doubles: 16-byte aligned)
    double* Bs = DT + 32 * KBC_LDT;         // [32][68]  chunk of the J-block rows of L (B operand of the update)
    double* Ps = Bs + 32 * KBC_LDB;         // [m-32][33] rows below the diagonal block
    __shared__ double sh[32];
    __shared__ double dg[32];               // 1 / diagonal of L_JJ
    __shared__ double colb[32];             // column j of L_JJ while it is being folded into the rows
    __shared__ double s_maxdiag;
    __shared__ int s_nfix;
    const int lp = blockIdx.x;
    if (a.active && a.active[lp] == 0) return;
    const int m = a.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* Mb = a.M + (size_t)lp * a.strideM;
    const int64_t ldm = a.ldm;

    // max_i M_ii (threshold of the safeguard)
    {
        double v = red_identity<RED_MAX>();
        for (int i = tid; i < m; i += KBC_NT) v = fmax(v, Mb[(size_t)i * ldm + i]);
        v = block_red<RED_MAX>(v, sh);
        if (tid == 0) { s_maxdiag = v; s_nfix = 0; }
        __syncthreads();
    }
    const double thresh = a.tau * s_maxdiag;

    for (int j0 = 0; j0 < m; j0 += 32) {
        const int nb = (m - j0 < 32) ? (m - j0) : 32;
        const int rows = m - j0;                       // panel rows (diag block included)
        const int ntile = (rows + 7) >> 3;             // 8-row tiles
        // ---------------- 1. left-looking update on the tensor pipe: acc = -M[:,J] + sum_K L[:,K] L[J,K]^T,
        //                     the sign is flipped when the accumulators are spilled to the shared panel.
        double acc[4][4][2];
        bool aok[4];
        int aoff[4];                     // 32-bit element offsets (m*ldm <= 2^16 here)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int tile = warp + i * KBC_NW;
            const int ra = j0 + tile * 8 + g;
            aok[i] = (tile < ntile) && (ra < m);
            aoff[i] = (aok[i] ? ra : j0) * (int)ldm + t;
        }
#pragma unroll
        for (int ti = 0; ti < 4; ++ti) {
            const int r = j0 + (warp + ti * KBC_NW) * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = j0 + ni * 8 + 2 * t;
                double v0 = 0.0, v1 = 0.0;
                if (aok[ti]) {
                    if (c < m) v0 = -Mb[(size_t)r * ldm + c];
                    if (c + 1 < m) v1 = -Mb[(size_t)r * ldm + c + 1];
                }
                acc[ti][ni][0] = v0;
                acc[ti][ni][1] = v1;
            }
        }
        const int nti = (ntile > warp) ? (ntile - warp + KBC_NW - 1) / KBC_NW : 0;   // row tiles of this warp
        for (int k0 = 0; k0 < j0; k0 += KBC_KC) {
            const int kc = (j0 - k0 < KBC_KC) ? (j0 - k0) : KBC_KC;
            __syncthreads();                                   // previous chunk fully consumed
            for (int idx = tid; idx < 32 * kc; idx += KBC_NT) {
                const int r = idx / kc, k = idx - r * kc;
                Bs[r * KBC_LDB + k] = (j0 + r < m) ? Mb[(size_t)(j0 + r) * ldm + k0 + k] : 0.0;
            }
            __syncthreads();
            switch (nti) {
                case 1: kbc_update_chunk<1>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                case 2: kbc_update_chunk<2>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                case 3: kbc_update_chunk<3>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                case 4: kbc_update_chunk<4>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                default: break;
            }
        }
        // accumulators -> shared panel (sign restored)
#pragma unroll
        for (int ti = 0; ti < 4; ++ti) {
            const int tile = warp + ti * KBC_NW;
            const int pr = tile * 8 + g;                // row inside the panel
            if (tile >= ntile || pr >= rows) continue;
            double* dst = (pr < 32) ? (D + pr * KBC_LD) : (Ps + (size_t)(pr - 32) * KBC_LD);
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                dst[ni * 8 + 2 * t] = -acc[ti][ni][0];
                dst[ni * 8 + 2 * t + 1] = -acc[ti][ni][1];
            }
        }
        __syncthreads();
        // ---------------- 2. diagonal block: warp 0 only, lane i keeps row i in REGISTERS (right-looking).
        //                     Column j: the pivot travels by warp shuffle from lane j, every lane applies the
        //                     safeguard p <= tau*maxdiag or NaN -> 1e128 alike, scales its entry by 1/sqrt(p),
        //                     publishes it in a 32-entry shared column and folds the column into its own row.
        if (warp == 0) {
            double arow[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) arow[c] = D[lane * KBC_LD + c];
            int nfix = 0;
            double my_inv = 1.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < nb) {
                    double p = __shfl_sync(0xffffffffu, arow[j], j);
                    const bool bad = !(p > thresh);
                    if (bad) p = kPivotBig;
                    const double inv = rsqrt(p);
                    const double lij = (lane == j) ? p * inv : arow[j] * inv;
                    if (lane == j) { my_inv = inv; nfix += bad ? 1 : 0; }
                    arow[j] = lij;
                    colb[lane] = lij;
                    __syncwarp();
#pragma unroll
                    for (int k = j + 1; k < 32; ++k) arow[k] = fma(-lij, colb[k], arow[k]);   // entries k > lane are unused
                    __syncwarp();
                }
            }
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (c <= lane) D[lane * KBC_LD + c] = arow[c];
            dg[lane] = my_inv;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nfix += __shfl_xor_sync(0xffffffffu, nfix, o);
            if (lane == 0) s_nfix += nfix;               // lane j counted its own pivot
        }
        __syncthreads();
        // transposed copy of L_JJ so that the substitution below reads 8 consecutive entries per step
        for (int idx = tid; idx < 32 * 32; idx += KBC_NT) {
            const int jj = idx >> 5, kk = idx & 31;
            DT[kk * KBC_LDT + jj] = (jj < nb && kk <= jj) ? D[jj * KBC_LD + kk] : 0.0;
        }
        __syncthreads();
        // ---------------- 3. rows below the block: x L_JJ^T = a, one thread per row, 8 columns at a time
        const int below = rows - 32;
        for (int r = tid; r < below; r += KBC_NT) {
            double* pr = Ps + (size_t)r * KBC_LD;
#pragma unroll 1
            for (int jb = 0; jb < 32; jb += 8) {
                double x8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x8[q] = pr[jb + q];
#pragma unroll 4
                for (int k = 0; k < jb; ++k) {
                    const double xk = pr[k];
                    const double2* lp = reinterpret_cast<const double2*>(DT + k * KBC_LDT + jb);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 lv = lp[q];
                        x8[2 * q] -= xk * lv.x;
                        x8[2 * q + 1] -= xk * lv.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double* lrow = DT + (jb + q) * KBC_LDT + jb;
                    const double xv = x8[q] * dg[jb + q];
                    x8[q] = xv;
#pragma unroll
                    for (int q2 = q + 1; q2 < 8; ++q2) x8[q2] -= xv * lrow[q2];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) pr[jb + q] = x8[q];
            }
        }
        __syncthreads();
        // ---------------- 4. panel -> global
        for (int idx = tid; idx < rows * 32; idx += KBC_NT) {
            const int pr = idx >> 5, c = idx & 31;
            if (c >= nb) continue;
            if (pr < 32) {
                if (c <= pr) Mb[(size_t)(j0 + pr) * ldm + j0 + c] = D[pr * KBC_LD + c];
            } else {
                Mb[(size_t)(j0 + pr) * ldm + j0 + c] = Ps[(size_t)(pr - 32) * KBC_LD + c];
            }
        }
        __syncthreads();
    }
    if (tid == 0 && a.scal) {
        a.scal[(size_t)lp * a.strideScal + S_MAXDIAG] = s_maxdiag;
        a.scal[(size_t)lp * a.strideScal + S_NFIXED] = (double)s_nfix;
    }
}

inline int potrf_batched_fused(double* M, int64_t ldm, int64_t strideM, int m, int batch, double* scal,
                               int64_t strideScal, double tau, const int* active, cudaStream_t st) {
    static int configured_dev = -1;
    int dev = 0;
    IPM_CUDA_OK(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        IPM_CUDA_OK(cudaFuncSetAttribute(kb_chol, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kbc_smem_bytes(KBC_MAX_M)));
        configured_dev = dev;
    }
    CholBatchedArgs a;
    a.M = M; a.ldm = ldm; a.strideM = strideM; a.scal = scal; a.strideScal = strideScal; a.tau = tau; a.m = m;
    a.active = active;
    kb_chol<<<batch, KBC_NT, kbc_smem_bytes(m), st>>>(a);
    count_launch();
    return launch_check();
}
#endif

}  // namespace ipm
