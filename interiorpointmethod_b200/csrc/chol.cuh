// Blocked right-looking Cholesky with the tiny-pivot safeguard, and the triangular solves.
// Replaces the reference's linear-solve seam `solve_linear` (main.py:176-182) on the normal-equations
// matrix of main.py:223-224.  All matrices are row-major, lower triangle; every kernel takes a batch
// index in blockIdx.z so the same code serves one LP (batch = 1) and the batched workload.
//
// One panel step (width NB) = k_chol_diag (factor the NBxNB diagonal block in shared memory)
//                           + k_chol_trsm (rows below: X L_JJ^T = A_panel by forward substitution)
//                           + dmma_nt_kernel<EPI=1> (trailing update C -= X X^T on the FP64 tensor pipe).
#pragma once
#include "common.cuh"
#include "dmma_gemm.cuh"
#include "dmma_ws.cuh"

namespace ipm {

struct CholArgs {
    double* M;  int64_t ldm;  int64_t strideM;     // batch stride in doubles
    double* scal; int64_t strideScal;              // per-LP scalar block (S_MAXDIAG in, S_NFIXED accumulated)
    double tau;                                    // relative pivot threshold
    int m, j0, nb;                                 // matrix order, panel start, panel width (<= NB)
    const int* active;                             // nullable per-LP flag
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// max_i M_ii  ->  scal[S_MAXDIAG];  also zeroes scal[S_NFIXED].   grid (1,1,batch), 256 threads
static __global__ void k_maxdiag(const double* M, int64_t ldm, int64_t strideM, int m, double* scal, int64_t strideScal,
                          const int* active) {
    __shared__ double sh[32];
    const int bz = blockIdx.z;
    if (active && active[bz] == 0) return;
    const double* Mb = M + (size_t)bz * strideM;
    double v = red_identity<RED_MAX>();
    for (int i = threadIdx.x; i < m; i += blockDim.x) v = fmax(v, Mb[(size_t)i * ldm + i]);
    v = block_red<RED_MAX>(v, sh);
    if (threadIdx.x == 0) {
        scal[(size_t)bz * strideScal + S_MAXDIAG] = v;
        scal[(size_t)bz * strideScal + S_NFIXED] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Factor the diagonal block M[j0:j0+nb, j0:j0+nb] in shared memory.   grid (1,1,batch), NT threads.
// Column j: every thread reads the pivot (shared-memory broadcast), applies the safeguard
//   p <= tau*maxdiag or NaN  ->  p = 1e128            
// and scales its rows; then the trailing block gets the rank-1 update, warps over rows, lanes over columns.
template <int NB, int NT>
static __global__ void __launch_bounds__(NT, 1) k_chol_diag(const CholArgs a) {
    constexpr int LD = NB + 1;
    extern __shared__ __align__(16) double smem[];
    double* S = smem;                 // [NB][LD]
    double* diag = smem + NB * LD;    // [NB]
    const int bz = blockIdx.z;
    if (a.active && a.active[bz] == 0) return;
    double* Mb = a.M + (size_t)bz * a.strideM + (size_t)a.j0 * a.ldm + a.j0;
    const int nb = a.nb, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;

    for (int idx = tid; idx < nb * nb; idx += NT) {
        const int i = idx / nb, j = idx - i * nb;
        if (j <= i) S[i * LD + j] = Mb[(size_t)i * a.ldm + j];
    }
    __syncthreads();
    const double thresh = a.tau * a.scal[(size_t)bz * a.strideScal + S_MAXDIAG];
    int nfix = 0;
    for (int j = 0; j < nb; ++j) {
        double p = S[j * LD + j];
        const bool bad = !(p > thresh);
        if (bad) p = kPivotBig;
        const double l = sqrt(p);
        for (int i = j + 1 + tid; i < nb; i += NT) S[i * LD + j] = S[i * LD + j] / l;
        if (tid == 0) { diag[j] = l; nfix += bad ? 1 : 0; }
        __syncthreads();
        for (int i = j + 1 + warp; i < nb; i += NW) {
            const double lij = S[i * LD + j];
            for (int k = j + 1 + lane; k <= i; k += 32) S[i * LD + k] -= lij * S[k * LD + j];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nb * nb; idx += NT) {
        const int i = idx / nb, j = idx - i * nb;
        if (j < i) Mb[(size_t)i * a.ldm + j] = S[i * LD + j];
        else if (j == i) Mb[(size_t)i * a.ldm + j] = diag[i];
    }
    if (tid == 0 && nfix) a.scal[(size_t)bz * a.strideScal + S_NFIXED] += (double)nfix;
}
template <int NB>
constexpr size_t chol_diag_smem() { return (size_t)(NB * (NB + 1) + NB) * sizeof(double); }

// ------------------------------------------------------------------------------------------------
// Rows below the diagonal block: solve X L_JJ^T = A_panel, one thread per row, forward substitution in
// register blocks of 8 columns; L_JJ^T lives in shared memory (broadcast 128-bit reads), the row's earlier
// x values in a [k][row] shared array (conflict-free).   grid (ceil(rows/ROWS),1,batch), ROWS threads.
template <int NB, int ROWS>
static __global__ void __launch_bounds__(ROWS, 1) k_chol_trsm(const CholArgs a) {
    constexpr int LDT = NB + 2;
    extern __shared__ __align__(16) double smem[];
    double* LsT = smem;                    // [NB][LDT]   LsT[k][j] = L[j][k], j >= k
    double* Xs = smem + NB * LDT;          // [NB][ROWS]
    const int bz = blockIdx.z;
    if (a.active && a.active[bz] == 0) return;
    double* Mb = a.M + (size_t)bz * a.strideM;
    const int tid = threadIdx.x;
    const int j0 = a.j0, j1 = a.j0 + NB;
    const double* Ld = Mb + (size_t)j0 * a.ldm + j0;
    for (int idx = tid; idx < NB * NB; idx += ROWS) {
        const int j = idx / NB, k = idx - j * NB;
        if (k <= j) LsT[k * LDT + j] = Ld[(size_t)j * a.ldm + k];
    }
    __syncthreads();
    const int r = j1 + blockIdx.x * ROWS + tid;
    if (r >= a.m) return;
    double* row = Mb + (size_t)r * a.ldm + j0;
#pragma unroll 1
    for (int jb = 0; jb < NB; jb += 8) {
        double acc[8];
        {
            const double2* rp = reinterpret_cast<const double2*>(row + jb);
#pragma unroll
            for (int q = 0; q < 4; ++q) { double2 v = rp[q]; acc[2 * q] = v.x; acc[2 * q + 1] = v.y; }
        }
#pragma unroll 4
        for (int k = 0; k < jb; ++k) {
            const double xk = Xs[k * ROWS + tid];
            const double2* lp = reinterpret_cast<const double2*>(LsT + k * LDT + jb);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 l = lp[q];
                acc[2 * q] -= xk * l.x;
                acc[2 * q + 1] -= xk * l.y;
            }
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const double* lrow = LsT + (jb + jj) * LDT + jb;
            const double x = acc[jj] / lrow[jj];
            acc[jj] = x;
            Xs[(jb + jj) * ROWS + tid] = x;
#pragma unroll
            for (int j2 = jj + 1; j2 < 8; ++j2) acc[j2] -= x * lrow[j2];
        }
        {
            double2* wp = reinterpret_cast<double2*>(row + jb);
#pragma unroll
            for (int q = 0; q < 4; ++q) wp[q] = make_double2(acc[2 * q], acc[2 * q + 1]);
        }
    }
}
template <int NB, int ROWS>
constexpr size_t chol_trsm_smem() { return (size_t)(NB * (NB + 2) + NB * ROWS) * sizeof(double); }

// ------------------------------------------------------------------------------------------------
// Host driver: in-place factorisation of `batch` matrices of order m.
template <int NB, int NT_DIAG, int ROWS>
inline int potrf_blocked(double* M, int64_t ldm, int64_t strideM, int m, int batch, double* scal,
                         int64_t strideScal, double tau, const int* active, cudaStream_t st) {
    static int configured_dev = -1;
    int dev = 0;
    IPM_CUDA_OK(cudaGetDevice(&dev));
    auto kd = k_chol_diag<NB, NT_DIAG>;
    auto kt = k_chol_trsm<NB, ROWS>;
    if (configured_dev != dev) {
        IPM_CUDA_OK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_diag_smem<NB>()));
        IPM_CUDA_OK(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)chol_trsm_smem<NB, ROWS>()));
        configured_dev = dev;
    }
    k_maxdiag<<<dim3(1, 1, batch), 256, 0, st>>>(M, ldm, strideM, m, scal, strideScal, active);
    count_launch();
    CholArgs a;
    a.M = M; a.ldm = ldm; a.strideM = strideM; a.scal = scal; a.strideScal = strideScal; a.tau = tau;
    a.m = m; a.active = active;
    for (int j0 = 0; j0 < m; j0 += NB) {
        a.j0 = j0;
        a.nb = (m - j0 < NB) ? (m - j0) : NB;
        kd<<<dim3(1, 1, batch), NT_DIAG, chol_diag_smem<NB>(), st>>>(a);
        count_launch();
        const int below = m - (j0 + NB);
        if (below > 0) {
            kt<<<dim3(ceil_div(below, ROWS), 1, batch), ROWS, chol_trsm_smem<NB, ROWS>(), st>>>(a);
            count_launch();
            DmmaArgs g;
            const double* panel = M + (size_t)(j0 + NB) * ldm + j0;
            g.P = panel; g.ldp = ldm; g.strideP = strideM;
            g.Q = panel; g.ldq = ldm; g.strideQ = strideM;
            g.dvec = nullptr; g.strideD = 0;
            g.C = M + (size_t)(j0 + NB) * ldm + (j0 + NB); g.ldc = ldm; g.strideC = strideM;
            g.rowsP = below; g.rowsQ = below; g.K = NB; g.lower_only = 1; g.active = active;
            IPM_TRY((dmma_syrk_auto<1>(g, batch, st)));
        }
    }
    return launch_check();
}

// ------------------------------------------------------------------------------------------------
// Triangular solves  L z = r  then  L^T y = z  for ONE large matrix, one launch per 64-wide block.
// Right-looking in both sweeps so that every vector entry is owned by exactly one thread:
//   forward : z_J = L_JJ^-1 r_J ;  r_i -= sum_{k in J} L[i][k] z_k   for rows i below J   (thread per row)
//   backward: y_J = L_JJ^-T z_J ;  z_k -= sum_{i in J} L[i][k] y_i   for columns k left of J (thread per column)
// Every CTA re-solves the 64x64 diagonal system (warp 0, pivot broadcast by warp shuffles) and then
// applies its share of the update; CTA 0 stores the solved block to `out`.
constexpr int TRSV_NB = 64;
constexpr int TRSV_NT = 128;

struct TrsvArgs {
    const double* L; int64_t ldm;
    double* v;        // in/out work vector (r then z), length m
    double* out;      // solved blocks (z for forward, y for backward), length m
    int m, j0, nb;
};

__device__ __forceinline__ void trsv_diag_solve(const double* Ls /*[64][65]*/, double* zs /*[64]*/, int nb,
                                                bool transposed) {
    // executed by warp 0; lane owns entries lane and lane+32
    const int lane = threadIdx.x & 31;
    double v0 = (lane < nb) ? zs[lane] : 0.0;
    double v1 = (lane + 32 < nb) ? zs[lane + 32] : 0.0;
    const double i0 = (lane < nb) ? 1.0 / Ls[lane * 65 + lane] : 0.0;
    const double i1 = (lane + 32 < nb) ? 1.0 / Ls[(lane + 32) * 65 + lane + 32] : 0.0;
    if (!transposed) {
        for (int j = 0; j < nb; ++j) {
            const double cand = (j < 32) ? v0 * i0 : v1 * i1;
            const double zj = __shfl_sync(0xffffffffu, cand, j & 31);
            if (j < 32) {
                if (lane == j) v0 = zj;
                if (lane > j) v0 -= Ls[lane * 65 + j] * zj;
                if (lane + 32 < nb) v1 -= Ls[(lane + 32) * 65 + j] * zj;
            } else {
                if (lane + 32 == j) v1 = zj;
                if (lane + 32 > j && lane + 32 < nb) v1 -= Ls[(lane + 32) * 65 + j] * zj;
            }
        }
    } else {
        for (int j = nb - 1; j >= 0; --j) {
            const double cand = (j < 32) ? v0 * i0 : v1 * i1;
            const double yj = __shfl_sync(0xffffffffu, cand, j & 31);
            if (j >= 32) {
                if (lane + 32 == j) v1 = yj;
                if (lane + 32 < j) v1 -= Ls[j * 65 + lane + 32] * yj;
                v0 -= Ls[j * 65 + lane] * yj;
            } else {
                if (lane == j) v0 = yj;
                if (lane < j) v0 -= Ls[j * 65 + lane] * yj;
            }
        }
    }
    if (lane < nb) zs[lane] = v0;
    if (lane + 32 < nb) zs[lane + 32] = v1;
}

static __global__ void __launch_bounds__(TRSV_NT) k_trsv_fwd(const TrsvArgs a) {
    __shared__ double Ls[64 * 65];
    __shared__ double zs[64];
    const int tid = threadIdx.x, nb = a.nb, j0 = a.j0;
    const double* Ld = a.L + (size_t)j0 * a.ldm + j0;
    for (int idx = tid; idx < nb * nb; idx += TRSV_NT) {
        const int i = idx / nb, j = idx - i * nb;
        Ls[i * 65 + j] = (j <= i) ? Ld[(size_t)i * a.ldm + j] : 0.0;
    }
    if (tid < nb) zs[tid] = a.v[j0 + tid];
    __syncthreads();
    if (tid < 32) trsv_diag_solve(Ls, zs, nb, false);
    __syncthreads();
    if (blockIdx.x == 0 && tid < nb) a.out[j0 + tid] = zs[tid];
    const int r = j0 + nb + blockIdx.x * TRSV_NT + tid;
    if (r >= a.m) return;
    const double* row = a.L + (size_t)r * a.ldm + j0;
    double acc = 0.0;
    if (nb == TRSV_NB) {
        const double2* rp = reinterpret_cast<const double2*>(row);
#pragma unroll 8
        for (int q = 0; q < TRSV_NB / 2; ++q) {
            const double2 l = rp[q];
            acc += l.x * zs[2 * q];
            acc += l.y * zs[2 * q + 1];
        }
    } else {
        for (int k = 0; k < nb; ++k) acc += row[k] * zs[k];
    }
    a.v[r] -= acc;
}

static __global__ void __launch_bounds__(TRSV_NT) k_trsv_bwd(const TrsvArgs a) {
    __shared__ double Ls[64 * 65];
    __shared__ double zs[64];
    const int tid = threadIdx.x, nb = a.nb, j0 = a.j0;
    const double* Ld = a.L + (size_t)j0 * a.ldm + j0;
    for (int idx = tid; idx < nb * nb; idx += TRSV_NT) {
        const int i = idx / nb, j = idx - i * nb;
        Ls[i * 65 + j] = (j <= i) ? Ld[(size_t)i * a.ldm + j] : 0.0;
    }
    if (tid < nb) zs[tid] = a.v[j0 + tid];
    __syncthreads();
    if (tid < 32) trsv_diag_solve(Ls, zs, nb, true);
    __syncthreads();
    if (blockIdx.x == 0 && tid < nb) a.out[j0 + tid] = zs[tid];
    const int k = blockIdx.x * TRSV_NT + tid;
    if (k >= j0) return;
    const double* col = a.L + (size_t)j0 * a.ldm + k;
    double acc = 0.0;
#pragma unroll 8
    for (int i = 0; i < nb; ++i) acc += col[(size_t)i * a.ldm] * zs[i];
    a.v[k] -= acc;
}

// rhs (destroyed) -> sol.  L is the factor produced by potrf_blocked.
inline int potrs_single(const double* L, int64_t ldm, int m, double* rhs, double* tmp, double* sol, cudaStream_t st) {
    TrsvArgs a;
    a.L = L; a.ldm = ldm; a.m = m;
    a.v = rhs; a.out = tmp;
    for (int j0 = 0; j0 < m; j0 += TRSV_NB) {
        a.j0 = j0; a.nb = (m - j0 < TRSV_NB) ? (m - j0) : TRSV_NB;
        const int below = m - (j0 + a.nb);
        k_trsv_fwd<<<(below > 0 ? ceil_div(below, TRSV_NT) : 1), TRSV_NT, 0, st>>>(a);
        count_launch();
    }
    a.v = tmp; a.out = sol;
    const int nblk = ceil_div(m, TRSV_NB);
    for (int jb = nblk - 1; jb >= 0; --jb) {
        a.j0 = jb * TRSV_NB; a.nb = (m - a.j0 < TRSV_NB) ? (m - a.j0) : TRSV_NB;
        const int left = a.j0;
        k_trsv_bwd<<<(left > 0 ? ceil_div(left, TRSV_NT) : 1), TRSV_NT, 0, st>>>(a);
        count_launch();
    }
    return launch_check();
}

// ------------------------------------------------------------------------------------------------
// Batched solve: one light CTA per LP (only the vector lives in shared memory, so 8 CTAs fit on an SM and the
// serial 32-step diagonal solves of different LPs overlap).  Forward then backward sweep in one kernel, so L is
// read from L2 the second time.  Per 32-wide block:
//   forward : r_I -= L[I, 0:I) z      (warp per row, lanes over columns, coalesced)  ;  z_I = L_II^-1 r_I
//   backward: y_I = L_II^-T z_I       ;  z[0:I) -= L[I, 0:I)^T y_I                   (thread per column, coalesced)
// The diagonal solves run in warp 0: lane i keeps row i (forward) / column i (backward) of L_II in registers and
// the freshly solved component is broadcast with a warp shuffle.
struct TrsvBatchedArgs {
    const double* L; int64_t ldm; int64_t strideM;
    double* v; int64_t strideV;      // rhs in, solution out (length m per LP)
    int m;
    const int* active;
};
constexpr int TRSVB_NT = 256;
constexpr int TRSVB_NW = TRSVB_NT / 32;

static __global__ void __launch_bounds__(TRSVB_NT, 3) k_trsv_batched(const TrsvBatchedArgs a) {
    extern __shared__ __align__(16) double smem_tb[];
    double* Ls0 = smem_tb;                  // [2][32][33] diagonal blocks, double buffered
    double* vec = smem_tb + 2 * 32 * 33;    // [m]
    const int bz = blockIdx.x;
    if (a.active && a.active[bz] == 0) return;
    const double* L = a.L + (size_t)bz * a.strideM;
    double* v = a.v + (size_t)bz * a.strideV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = a.m;
    const int64_t ldm = a.ldm;
    for (int i = tid; i < m; i += TRSVB_NT) vec[i] = v[i];
    const int nblk = (m + 31) >> 5;
    // the 4 entries of a 32x32 diagonal block this thread stages (issued one block ahead of their use)
    auto fetch_block = [&](int I, double (&r)[4]) {
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * TRSVB_NT;
            const int i = idx >> 5, c = idx & 31;
            r[q] = (i < nb && c <= i) ? L[(size_t)(i0 + i) * ldm + i0 + c] : 0.0;
        }
    };
    auto put_block = [&](double* Ls, const double (&r)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * TRSVB_NT;
            Ls[(idx >> 5) * 33 + (idx & 31)] = r[q];
        }
    };
    double blk[4];
    fetch_block(0, blk);
    put_block(Ls0, blk);
    __syncthreads();
    // ---- forward
    for (int I = 0; I < nblk; ++I) {
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
        double* Ls = Ls0 + (I & 1) * 32 * 33;
        if (I + 1 < nblk) fetch_block(I + 1, blk);            // in flight during the GEMV and the solve
        if (i0 > 0) {
            // rows warp, warp+8, warp+16, warp+24 of the block at once: all loads issued before any reduction
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = lane; k < i0; k += 32) {
                const double zk = vec[k];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = warp + q * TRSVB_NW;
                    if (rr < nb) acc[q] += L[(size_t)(i0 + rr) * ldm + k] * zk;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double sum = warp_sum(acc[q]);
                const int rr = warp + q * TRSVB_NW;
                if (lane == 0 && rr < nb) vec[i0 + rr] -= sum;
            }
            __syncthreads();
        }
        if (warp == 0) {
            const bool ok = lane < nb;
            double x = ok ? vec[i0 + lane] : 0.0;
            const double inv = ok ? 1.0 / Ls[lane * 33 + lane] : 0.0;
#pragma unroll 8
            for (int j = 0; j < nb; ++j) {
                const double zj = __shfl_sync(0xffffffffu, x * inv, j);
                if (lane == j) x = zj;
                else if (lane > j) x -= Ls[lane * 33 + j] * zj;
            }
            if (ok) vec[i0 + lane] = x;
        }
        if (I + 1 < nblk) put_block(Ls0 + ((I + 1) & 1) * 32 * 33, blk);
        __syncthreads();
    }
    // ---- backward (the last diagonal block is still in its buffer)
    for (int I = nblk - 1; I >= 0; --I) {
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
        double* Ls = Ls0 + (I & 1) * 32 * 33;
        if (I > 0) fetch_block(I - 1, blk);
        if (warp == 0) {
            const bool ok = lane < nb;
            double x = ok ? vec[i0 + lane] : 0.0;
            const double inv = ok ? 1.0 / Ls[lane * 33 + lane] : 0.0;
#pragma unroll 8
            for (int i = nb - 1; i >= 0; --i) {
                const double yi = __shfl_sync(0xffffffffu, x * inv, i);
                if (lane == i) x = yi;
                else if (lane < i) x -= Ls[i * 33 + lane] * yi;
            }
            if (ok) vec[i0 + lane] = x;
        }
        if (I > 0) put_block(Ls0 + ((I - 1) & 1) * 32 * 33, blk);
        __syncthreads();
        for (int k = tid; k < i0; k += TRSVB_NT) {
            const double* col = L + (size_t)i0 * ldm + k;
            double acc = 0.0;
            if (nb == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += col[(size_t)i * ldm] * vec[i0 + i];
            } else {
                for (int i = 0; i < nb; ++i) acc += col[(size_t)i * ldm] * vec[i0 + i];
            }
            vec[k] -= acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < m; i += TRSVB_NT) v[i] = vec[i];
}
inline size_t trsv_batched_smem(int m) { return (size_t)(2 * 32 * 33 + m) * sizeof(double); }
#endif

}  // namespace ipm
