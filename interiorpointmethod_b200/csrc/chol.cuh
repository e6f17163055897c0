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
#include "chol_batched.cuh"

namespace ipm {

struct CholArgs {
    double* M;  int64_t ldm;  int64_t strideM;     // batch stride in doubles
    double* scal; int64_t strideScal;              // per-LP scalar block (S_MAXDIAG in, S_NFIXED accumulated)
    double tau;                                    // relative pivot threshold
    int m, j0, nb;                                 // matrix order, panel start, panel width (<= NB)
    const int* active;                             // nullable per-LP flag
    unsigned char* dep = nullptr;                  // nullable, m bytes (one matrix): dependent-row mask
    int dep_mode = 0;                              // 1: pivots of masked rows are replaced whatever their value
                                                   // 2: detect - the mask is WRITTEN (1 where the pivot was replaced)
};
// Dependent rows of a rank-deficient A (ipm_detect_dependent_rows, include/ipm_b200.h): threaded through the
// single-matrix drivers below.
struct DepMask {
    unsigned char* mask = nullptr;
    int mode = 0;
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
// Two-level right-looking inside the CTA, 32 columns at a time:
//   (1) warp 0 factors the 32x32 sub-block with lane i holding row i in registers; the pivot travels by warp
//       shuffle from lane j, every lane applies the safeguard  p <= tau*maxdiag or NaN -> 1e128
//       (SURVEY.md App. A.4), divides its entry by sqrt(p), publishes the column through shared memory and
//       folds it into its own row;
//   (2) the rows of the block below it are solved one thread per row against the transposed sub-block;
//   (3) the rest of the block gets the rank-32 update, warps over rows, lanes over columns.
template <int NB, int NT>
static __global__ void __launch_bounds__(NT, 1) k_chol_diag(const CholArgs a) {
    constexpr int LD = NB + 1;
    extern __shared__ __align__(16) double smem[];
    double* S = smem;                         // [NB][LD]
    double* DT = smem + NB * LD + (NB * LD & 1);   // [32][34] transposed sub-block, 16-byte aligned
    double* dinv = DT + 32 * 34;              // [32]
    double* colb = dinv + 32;                 // [32]
    __shared__ int s_nfix;
    const int bz = blockIdx.z;
    if (a.active && a.active[bz] == 0) return;
    double* Mb = a.M + (size_t)bz * a.strideM + (size_t)a.j0 * a.ldm + a.j0;
    const int nb = a.nb, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;

    if (nb == NB) {
        // full block: compile-time index math and 8 independent loads in flight per thread
#pragma unroll 8
        for (int idx = tid; idx < NB * NB; idx += NT) {
            const int i = idx / NB, j = idx % NB;
            if (j <= i) S[i * LD + j] = Mb[(size_t)i * a.ldm + j];
        }
    } else {
#pragma unroll 4
        for (int idx = tid; idx < nb * nb; idx += NT) {
            const int i = idx / nb, j = idx - i * nb;
            if (j <= i) S[i * LD + j] = Mb[(size_t)i * a.ldm + j];
        }
    }
    if (tid == 0) s_nfix = 0;
    __syncthreads();
    const double thresh = a.tau * a.scal[(size_t)bz * a.strideScal + S_MAXDIAG];
    for (int c0 = 0; c0 < nb; c0 += 32) {
        const int w = (nb - c0 < 32) ? (nb - c0) : 32;
        // ---- (1) 32x32 sub-block, warp 0
        if (warp == 0) {
            double arow[32];
            const bool ok = lane < w;
#pragma unroll
            for (int c = 0; c < 32; ++c) arow[c] = (ok && c <= lane) ? S[(c0 + lane) * LD + c0 + c] : 0.0;
            int nfix = 0;
            double my_inv = 1.0;
            const bool masked = a.dep != nullptr && a.dep_mode == 1;
            const unsigned forced_bits = masked ? __ballot_sync(0xffffffffu, ok && a.dep[a.j0 + c0 + lane] != 0) : 0u;
            bool my_bad = false;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < w) {
                    double p = __shfl_sync(0xffffffffu, arow[j], j);
                    const bool bad = !(p > thresh) || ((forced_bits >> j) & 1u);
                    if (bad) p = kPivotBig;
                    const double l = sqrt(p);
                    const double lij = (lane == j) ? l : arow[j] / l;
                    if (lane == j) { my_inv = 1.0 / l; nfix += bad ? 1 : 0; my_bad = bad; }
                    arow[j] = lij;
                    colb[lane] = lij;
                    __syncwarp();
#pragma unroll
                    for (int k = j + 1; k < 32; ++k) arow[k] = fma(-lij, colb[k], arow[k]);
                    __syncwarp();
                }
            }
            if (ok) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c <= lane) S[(c0 + lane) * LD + c0 + c] = arow[c];
                dinv[lane] = my_inv;
                if (a.dep != nullptr && a.dep_mode == 2) a.dep[a.j0 + c0 + lane] = my_bad ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nfix += __shfl_xor_sync(0xffffffffu, nfix, o);
            if (lane == 0) s_nfix += nfix;
        }
        __syncthreads();
        const int r0 = c0 + w;                    // first row below the sub-block
        if (r0 < nb) {
            for (int idx = tid; idx < 32 * 32; idx += NT) {
                const int jj = idx >> 5, kk = idx & 31;
                DT[kk * 34 + jj] = (jj < w && kk <= jj) ? S[(c0 + jj) * LD + c0 + kk] : 0.0;
            }
            __syncthreads();
            // ---- (2) rows below: x L^T = a, 8 columns at a time (w == 32 here: only the last sub-block is ragged)
            for (int r = r0 + tid; r < nb; r += NT) {
                double* pr = S + r * LD + c0;
#pragma unroll 1
                for (int jb = 0; jb < 32; jb += 8) {
                    double x8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) x8[q] = pr[jb + q];
#pragma unroll 4
                    for (int k = 0; k < jb; ++k) {
                        const double xk = pr[k];
                        const double2* lp = reinterpret_cast<const double2*>(DT + k * 34 + jb);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double2 lv = lp[q];
                            x8[2 * q] -= xk * lv.x;
                            x8[2 * q + 1] -= xk * lv.y;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double* lrow = DT + (jb + q) * 34 + jb;
                        const double xv = x8[q] / lrow[q];
                        x8[q] = xv;
#pragma unroll
                        for (int q2 = q + 1; q2 < 8; ++q2) x8[q2] -= xv * lrow[q2];
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) pr[jb + q] = x8[q];
                }
            }
            __syncthreads();
            // ---- (3) rank-32 update of the remaining lower triangle
            for (int i = r0 + warp; i < nb; i += NW) {
                const double* ri = S + i * LD + c0;
                for (int k = r0 + lane; k <= i; k += 32) {
                    const double* rk = S + k * LD + c0;
                    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
                    for (int c = 0; c < 32; c += 2) {
                        acc0 += ri[c] * rk[c];
                        acc1 += ri[c + 1] * rk[c + 1];
                    }
                    S[i * LD + k] -= acc0 + acc1;
                }
            }
            __syncthreads();
        }
    }
    for (int idx = tid; idx < nb * nb; idx += NT) {
        const int i = idx / nb, j = idx - i * nb;
        if (j <= i) Mb[(size_t)i * a.ldm + j] = S[i * LD + j];
    }
    if (tid == 0 && s_nfix) a.scal[(size_t)bz * a.strideScal + S_NFIXED] += (double)s_nfix;
}
template <int NB>
constexpr size_t chol_diag_smem() { return (size_t)(NB * (NB + 1) + 1 + 32 * 34 + 64) * sizeof(double); }

// ------------------------------------------------------------------------------------------------
// Rows below the diagonal block: solve X L_JJ^T = A_panel, one thread per row, forward substitution in
// register blocks of 8 columns; L_JJ^T lives in shared memory (broadcast 128-bit reads), the row's earlier
// x values in a [k][row] shared array (conflict-free).   grid (ceil(rows/ROWS),1,batch), ROWS threads.
constexpr int TRSM_NT = 256;      // all threads stage L_JJ^T, the first ROWS of them own one row each
template <int NB, int ROWS>
static __global__ void __launch_bounds__(TRSM_NT, 1) k_chol_trsm(const CholArgs a) {
    static_assert(ROWS <= TRSM_NT, "one thread per row");
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
#pragma unroll 8
    for (int idx = tid; idx < NB * NB; idx += TRSM_NT) {
        const int j = idx / NB, k = idx % NB;
        if (k <= j) LsT[k * LDT + j] = Ld[(size_t)j * a.ldm + k];
    }
    __syncthreads();
    const int r = j1 + blockIdx.x * ROWS + tid;
    if (tid >= ROWS || r >= a.m) return;
    double* row = Mb + (size_t)r * a.ldm + j0;
    // Register blocks of CB columns.  Every x[j] sees the same operations in the same order whatever CB is (columns
    // k = 0..j-1 in increasing order, then the division), so the block width is a pure scheduling choice: 32 gives the
    // rank-1 updates 32 independent FMAs per column instead of 8 (the kernel is latency-bound: one or two warps per
    // SM; measured 50 -> see profiles/ microseconds per 128-wide panel).
    constexpr int CB = (NB % 32 == 0) ? 32 : 8;
#pragma unroll 1
    for (int jb = 0; jb < NB; jb += CB) {
        double acc[CB];
        {
            const double2* rp = reinterpret_cast<const double2*>(row + jb);
#pragma unroll
            for (int q = 0; q < CB / 2; ++q) { double2 v = rp[q]; acc[2 * q] = v.x; acc[2 * q + 1] = v.y; }
        }
#pragma unroll 2
        for (int k = 0; k < jb; ++k) {
            const double xk = Xs[k * ROWS + tid];
            const double2* lp = reinterpret_cast<const double2*>(LsT + k * LDT + jb);
#pragma unroll
            for (int q = 0; q < CB / 2; ++q) {
                const double2 l = lp[q];
                acc[2 * q] -= xk * l.x;
                acc[2 * q + 1] -= xk * l.y;
            }
        }
#pragma unroll
        for (int jj = 0; jj < CB; ++jj) {
            const double* lrow = LsT + (jb + jj) * LDT + jb;
            const double x = acc[jj] / lrow[jj];
            acc[jj] = x;
            Xs[(jb + jj) * ROWS + tid] = x;
#pragma unroll
            for (int j2 = jj + 1; j2 < CB; ++j2) acc[j2] -= x * lrow[j2];
        }
        {
            double2* wp = reinterpret_cast<double2*>(row + jb);
#pragma unroll
            for (int q = 0; q < CB / 2; ++q) wp[q] = make_double2(acc[2 * q], acc[2 * q + 1]);
        }
    }
}
template <int NB, int ROWS>
constexpr size_t chol_trsm_smem() { return (size_t)(NB * (NB + 2) + NB * ROWS) * sizeof(double); }

// ------------------------------------------------------------------------------------------------
// side stream of the look-ahead (one per device AND host thread, created on first use: two threads factoring on
// one device through distinct handles never share the stream or re-record each other's events)
struct CholSide {
    cudaStream_t st = nullptr;
    cudaEvent_t strip_done = nullptr, rest_done = nullptr;
    int ensure() {
        if (st) return IPM_OK;
        IPM_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        IPM_CUDA_OK(cudaEventCreateWithFlags(&strip_done, cudaEventDisableTiming));
        IPM_CUDA_OK(cudaEventCreateWithFlags(&rest_done, cudaEventDisableTiming));
        return IPM_OK;
    }
};
static thread_local CholSide g_chol_side[16];
static bool g_chol_lookahead = true;
// one matrix: the diagonal blocks go through the fused kernel of the batched solver (32-wide sub-panels with the
// rank-32 updates on the tensor pipe and look-ahead inside the CTA) instead of k_chol_diag
inline std::atomic<int>& chol_fused_diag() {
    static std::atomic<int> on{1};
    return on;
}

// ------------------------------------------------------------------------------------------------
// Host driver: in-place factorisation of `batch` matrices of order m.
template <int NB, int NT_DIAG, int ROWS>
inline int potrf_blocked(double* M, int64_t ldm, int64_t strideM, int m, int batch, double* scal,
                         int64_t strideScal, double tau, const int* active, cudaStream_t st, DepMask dm = DepMask()) {
    auto kd = k_chol_diag<NB, NT_DIAG>;
    auto kt = k_chol_trsm<NB, ROWS>;
    IPM_TRY(ensure_dyn_smem(kd, chol_diag_smem<NB>()));
    IPM_TRY(ensure_dyn_smem(kt, chol_trsm_smem<NB, ROWS>()));
    k_maxdiag<<<dim3(1, 1, batch), 256, 0, st>>>(M, ldm, strideM, m, scal, strideScal, active);
    count_launch();
    CholArgs a;
    a.M = M; a.ldm = ldm; a.strideM = strideM; a.scal = scal; a.strideScal = strideScal; a.tau = tau;
    a.m = m; a.active = active;
    a.dep = (batch == 1) ? dm.mask : nullptr; a.dep_mode = (batch == 1) ? dm.mode : 0;
    // Look-ahead (one large matrix, 128-wide panels): the trailing update of panel j is split into the tile column
    // the next panel lives in ("strip", a few microseconds) and the rest, which runs on a side stream on all SMs
    // but one while the main stream already factors the next diagonal block (one CTA, latency-bound, about as
    // long as the rest of the update).  Same operations on the same entries in the same order: bitwise identical.
    DmmaArgs probe;
    probe.P = M; probe.Q = M; probe.C = M; probe.ldp = probe.ldq = probe.ldc = ldm;
    probe.strideP = probe.strideQ = probe.strideC = 0; probe.dvec = nullptr; probe.strideD = 0;
    probe.lower_only = 1; probe.rowsP = probe.rowsQ = m;
    const bool lookahead = (batch == 1) && (NB == WS_BM) && (m >= 8 * NB) && ws_eligible(probe) && g_chol_lookahead;
    CholSide* side = nullptr;
    if (lookahead) {
        int dev = 0;
        IPM_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 16) return IPM_ERR_ARG;
        side = &g_chol_side[dev];
        IPM_TRY(side->ensure());
    }
    bool rest_pending = false;
    for (int j0 = 0; j0 < m; j0 += NB) {
        a.j0 = j0;
        a.nb = (m - j0 < NB) ? (m - j0) : NB;
        // (not with a dependent-row mask: on rank-deficient LPs the opt-in path is sensitive to the last bits of the
        // factor - QAP12 converges in 112 iterations with k_chol_diag's sqrt/divide pivots and stalls 1e-5 short of the
        // optimum with the fused kernel's rsqrt/multiply ones - so that path keeps the kernel it was validated with)
        if (batch == 1 && NB <= KBC_MAX_M && scal != nullptr && chol_fused_diag().load() != 0 && (ldm % 2 == 0) &&
            ((reinterpret_cast<uintptr_t>(M) & 15) == 0) && a.dep == nullptr) {
            IPM_TRY(potrf_diag_block_fused(M, ldm, j0, a.nb, scal, tau, st, a.dep, a.dep_mode));
        } else {
            kd<<<dim3(1, 1, batch), NT_DIAG, chol_diag_smem<NB>(), st>>>(a);
            count_launch();
        }
        const int below = m - (j0 + NB);
        if (below > 0) {
            kt<<<dim3(ceil_div(below, ROWS), 1, batch), TRSM_NT, chol_trsm_smem<NB, ROWS>(), st>>>(a);
            count_launch();
            DmmaArgs g;
            const double* panel = M + (size_t)(j0 + NB) * ldm + j0;
            g.P = panel; g.ldp = ldm; g.strideP = strideM;
            g.Q = panel; g.ldq = ldm; g.strideQ = strideM;
            g.dvec = nullptr; g.strideD = 0;
            g.C = M + (size_t)(j0 + NB) * ldm + (j0 + NB); g.ldc = ldm; g.strideC = strideM;
            g.rowsP = below; g.rowsQ = below; g.K = NB; g.lower_only = 1; g.active = active;
            if (!lookahead) {
                IPM_TRY((dmma_syrk_auto<1>(g, batch, st)));
                continue;
            }
            // strip: tile column 0 of the update, after the rest of the previous panel has left those entries
            if (rest_pending) IPM_CUDA_OK(cudaStreamWaitEvent(st, side->rest_done, 0));
            g.col0_only = 1;
            IPM_TRY((dmma_ws_launch<1, false>(g, 1, st)));
            IPM_CUDA_OK(cudaEventRecord(side->strip_done, st));
            // rest: the same update one tile further down and right, on the side stream, one SM left free
            if (below > NB) {
                DmmaArgs r = g;
                r.col0_only = 0;
                r.P = panel + (size_t)NB * ldm; r.Q = r.P;
                r.C = g.C + (size_t)NB * ldm + NB;
                r.rowsP = r.rowsQ = below - NB;
                r.max_ctas = kNumSMs - 1;
                IPM_CUDA_OK(cudaStreamWaitEvent(side->st, side->strip_done, 0));
                IPM_TRY((dmma_ws_launch<1, false>(r, 1, side->st)));
                IPM_CUDA_OK(cudaEventRecord(side->rest_done, side->st));
                rest_pending = true;
            } else {
                rest_pending = false;
            }
        }
    }
    if (rest_pending) IPM_CUDA_OK(cudaStreamWaitEvent(st, side->rest_done, 0));
    return launch_check();
}

// ------------------------------------------------------------------------------------------------
// Triangular solves  L z = r  then  L^T y = z  for ONE large matrix.
// Right-looking in both sweeps so that every vector entry is owned by exactly one thread:
//   forward : z_J = L_JJ^-1 r_J ;  r_i -= sum_{k in J} L[i][k] z_k   for rows i below J   (thread per row)
//   backward: y_J = L_JJ^-T z_J ;  z_k -= sum_{i in J} L[i][k] y_i   for columns k left of J (thread per column)
// Every CTA re-solves the diagonal system (warp 0, pivot broadcast by warp shuffles) and then applies its share
// of the update; CTA 0 stores the solved block to `out`.
struct TrsvArgs {
    const double* L; int64_t ldm;
    double* v;        // in/out work vector (r then z), length m
    double* out;      // solved blocks (z for forward, y for backward), length m
    int m, j0, nb;
};

// One launch per 128 columns, 256 threads.  Every CTA solves the
// 128x128 triangular block in shared memory (4 sub-blocks of 32: warp 0 solves, all threads update) and then
// applies its share of the update to the rest of the vector.
constexpr int TRSV128_NB = 128;
constexpr int TRSV128_NT = 256;
constexpr int TRSV128_LD = TRSV128_NB + 1;
inline size_t trsv128_smem() { return (size_t)(TRSV128_NB * TRSV128_LD + TRSV128_NB) * sizeof(double); }

__device__ __forceinline__ void trsv128_load(const TrsvArgs& a, double* S, double* vec) {
    const int tid = threadIdx.x, nb = a.nb;
    const double* Ld = a.L + (size_t)a.j0 * a.ldm + a.j0;
    if (nb == TRSV128_NB) {
#pragma unroll 8
        for (int idx = tid; idx < TRSV128_NB * TRSV128_NB; idx += TRSV128_NT) {
            const int i = idx / TRSV128_NB, j = idx % TRSV128_NB;
            if (j <= i) S[i * TRSV128_LD + j] = Ld[(size_t)i * a.ldm + j];
        }
    } else {
#pragma unroll 4
        for (int idx = tid; idx < nb * nb; idx += TRSV128_NT) {
            const int i = idx / nb, j = idx - i * nb;
            if (j <= i) S[i * TRSV128_LD + j] = Ld[(size_t)i * a.ldm + j];
        }
    }
    if (tid < nb) vec[tid] = a.v[a.j0 + tid];
    __syncthreads();
}

static __global__ void __launch_bounds__(TRSV128_NT, 1) k_trsv_fwd128(const TrsvArgs a) {
    extern __shared__ __align__(16) double sm128[];
    double* S = sm128;
    double* vec = sm128 + TRSV128_NB * TRSV128_LD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = a.nb, j0 = a.j0;
    trsv128_load(a, S, vec);
    for (int c0 = 0; c0 < nb; c0 += 32) {
        const int w = (nb - c0 < 32) ? (nb - c0) : 32;
        if (warp == 0) {
            const bool ok = lane < w;
            double x = ok ? vec[c0 + lane] : 0.0;
            const double inv = ok ? 1.0 / S[(c0 + lane) * TRSV128_LD + c0 + lane] : 0.0;
#pragma unroll 8
            for (int j = 0; j < w; ++j) {
                const double zj = __shfl_sync(0xffffffffu, x * inv, j);
                if (lane == j) x = zj;
                else if (lane > j && ok) x -= S[(c0 + lane) * TRSV128_LD + c0 + j] * zj;
            }
            if (ok) vec[c0 + lane] = x;
        }
        __syncthreads();
        for (int r = c0 + w + tid; r < nb; r += TRSV128_NT) {
            const double* row = S + r * TRSV128_LD + c0;
            double acc = 0.0;
#pragma unroll 8
            for (int c = 0; c < w; ++c) acc += row[c] * vec[c0 + c];
            vec[r] -= acc;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid < nb) a.out[j0 + tid] = vec[tid];
    const int r = j0 + nb + blockIdx.x * TRSV128_NT + tid;
    if (r >= a.m) return;
    const double* row = a.L + (size_t)r * a.ldm + j0;
    double acc0 = 0.0, acc1 = 0.0;
    if (nb == TRSV128_NB) {
        const double2* rp = reinterpret_cast<const double2*>(row);
#pragma unroll 16
        for (int q = 0; q < TRSV128_NB / 2; ++q) {
            const double2 l = rp[q];
            acc0 += l.x * vec[2 * q];
            acc1 += l.y * vec[2 * q + 1];
        }
    } else {
        for (int k = 0; k < nb; ++k) acc0 += row[k] * vec[k];
    }
    a.v[r] -= acc0 + acc1;
}

static __global__ void __launch_bounds__(TRSV128_NT, 1) k_trsv_bwd128(const TrsvArgs a) {
    extern __shared__ __align__(16) double sm128[];
    double* S = sm128;
    double* vec = sm128 + TRSV128_NB * TRSV128_LD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = a.nb, j0 = a.j0;
    trsv128_load(a, S, vec);
    const int nsb = (nb + 31) >> 5;
    for (int sb = nsb - 1; sb >= 0; --sb) {
        const int c0 = sb << 5;
        const int w = (nb - c0 < 32) ? (nb - c0) : 32;
        if (warp == 0) {
            const bool ok = lane < w;
            double x = ok ? vec[c0 + lane] : 0.0;
            const double inv = ok ? 1.0 / S[(c0 + lane) * TRSV128_LD + c0 + lane] : 0.0;
#pragma unroll 8
            for (int i = w - 1; i >= 0; --i) {
                const double yi = __shfl_sync(0xffffffffu, x * inv, i);
                if (lane == i) x = yi;
                else if (lane < i) x -= S[(c0 + i) * TRSV128_LD + c0 + lane] * yi;
            }
            if (ok) vec[c0 + lane] = x;
        }
        __syncthreads();
        for (int k = tid; k < c0; k += TRSV128_NT) {
            double acc = 0.0;
#pragma unroll 8
            for (int i = 0; i < w; ++i) acc += S[(c0 + i) * TRSV128_LD + k] * vec[c0 + i];
            vec[k] -= acc;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid < nb) a.out[j0 + tid] = vec[tid];
    const int k = blockIdx.x * TRSV128_NT + tid;
    if (k >= j0) return;
    const double* col = a.L + (size_t)j0 * a.ldm + k;
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
    for (int i = 0; i + 1 < nb; i += 2) {
        acc0 += col[(size_t)i * a.ldm] * vec[i];
        acc1 += col[(size_t)(i + 1) * a.ldm] * vec[i + 1];
    }
    if (nb & 1) acc0 += col[(size_t)(nb - 1) * a.ldm] * vec[nb - 1];
    a.v[k] -= acc0 + acc1;
}

// rhs (destroyed) -> sol.  L is the factor produced by potrf_blocked.
inline int potrs_single_blocks(const double* L, int64_t ldm, int m, double* rhs, double* tmp, double* sol, cudaStream_t st) {
    IPM_TRY(ensure_dyn_smem(k_trsv_fwd128, trsv128_smem()));
    IPM_TRY(ensure_dyn_smem(k_trsv_bwd128, trsv128_smem()));
    TrsvArgs a;
    a.L = L; a.ldm = ldm; a.m = m;
    a.v = rhs; a.out = tmp;
    for (int j0 = 0; j0 < m; j0 += TRSV128_NB) {
        a.j0 = j0; a.nb = (m - j0 < TRSV128_NB) ? (m - j0) : TRSV128_NB;
        const int below = m - (j0 + a.nb);
        k_trsv_fwd128<<<(below > 0 ? ceil_div(below, TRSV128_NT) : 1), TRSV128_NT, trsv128_smem(), st>>>(a);
        count_launch();
    }
    a.v = tmp; a.out = sol;
    const int nblk = ceil_div(m, TRSV128_NB);
    for (int jb = nblk - 1; jb >= 0; --jb) {
        a.j0 = jb * TRSV128_NB; a.nb = (m - a.j0 < TRSV128_NB) ? (m - a.j0) : TRSV128_NB;
        const int left = a.j0;
        k_trsv_bwd128<<<(left > 0 ? ceil_div(left, TRSV128_NT) : 1), TRSV128_NT, trsv128_smem(), st>>>(a);
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
    double* out = nullptr;           // optional separate destination (same stride)
    int only_flag = 0;               // != 0: only LPs whose active flag equals it (corrector refinement: 3)
    int accumulate = 0;              // != 0: destination += solution
};
__device__ __forceinline__ bool trsv_batched_skip(const TrsvBatchedArgs& a, int bz) {
    if (!a.active) return false;
    const int f = a.active[bz];
    return a.only_flag ? (f != a.only_flag) : (f == 0);
}
constexpr int TRSVB_NT = 256;
constexpr int TRSVB_NW = TRSVB_NT / 32;

static __global__ void __launch_bounds__(TRSVB_NT, 3) k_trsv_batched(const TrsvBatchedArgs a) {
    extern __shared__ __align__(16) double smem_tb[];
    double* Ls0 = smem_tb;                  // [2][32][33] diagonal blocks, double buffered
    double* vec = smem_tb + 2 * 32 * 33;    // [m]
    const int bz = blockIdx.x;
    if (trsv_batched_skip(a, bz)) return;
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
#pragma unroll 4
            for (int k = lane; k < i0; k += 32) {       // unrolled: up to 16 independent row loads in flight
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
    double* dst = a.out ? a.out + (size_t)bz * a.strideV : v;
    if (a.accumulate) {
        for (int i = tid; i < m; i += TRSVB_NT) dst[i] = dst[i] + vec[i];
    } else {
        for (int i = tid; i < m; i += TRSVB_NT) dst[i] = vec[i];
    }
}
// ---- variant for m <= 256 (at most 8 diagonal blocks): the 32-step substitution chains of the diagonal blocks
// were the critical path (every other warp waited on warp 0), so each warp first INVERTS one 32x32 diagonal block
// (lane j solves L x = e_j in registers, no cross-lane traffic) and both sweeps apply the inverse as a 32x32
// mat-vec.  The off-diagonal part is unchanged (plain substitution by blocks).
constexpr int TRSVI_MAX_BLK = 8;
inline size_t trsv_batched_inv_smem(int m) { return (size_t)(TRSVI_MAX_BLK * 32 * 33 + m + 32 * TRSVB_NW) * sizeof(double); }

static __device__ __forceinline__ void d_trsv_batched_inv(const TrsvBatchedArgs a) {
    extern __shared__ __align__(16) double smem_ti[];
    double* Linv = smem_ti;                              // [8][32][33] block inverses (staging area first)
    double* vec = smem_ti + TRSVI_MAX_BLK * 32 * 33;     // [m]
    double* dinv = vec + a.m;                            // [8 warps][32]
    const int bz = blockIdx.x;
    if (trsv_batched_skip(a, bz)) return;
    const double* L = a.L + (size_t)bz * a.strideM;
    double* v = a.v + (size_t)bz * a.strideV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = a.m;
    const int64_t ldm = a.ldm;
    const int nblk = (m + 31) >> 5;
    // The forward sweep walks L block row by block row with a dependent chain of global round trips; with the lower
    // triangle on its way to L2 from the start they are L2 hits (the batched factors of 8192 LPs do not stay there).
    if ((ldm & 1) == 0 && (reinterpret_cast<uintptr_t>(L) & 15) == 0) {
        for (int r = tid; r < m; r += TRSVB_NT) {
            const uint32_t bytes = (uint32_t)(((r + 1) * 8 + 15) & ~15);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(L + (size_t)r * ldm), "r"(bytes) : "memory");
        }
    }
    for (int i = tid; i < m; i += TRSVB_NT) vec[i] = v[i];
    // ---- block inverses, one warp per block
    for (int blk = warp; blk < nblk; blk += TRSVB_NW) {
        const int i0 = blk << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
        double* Ls = Linv + blk * 32 * 33;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            // row i of the block, lane = column (coalesced); identity outside the matrix
            double val = (i == lane) ? 1.0 : 0.0;
            if (i < nb && lane <= i) val = L[(size_t)(i0 + i) * ldm + i0 + lane];
            Ls[i * 33 + lane] = val;
        }
        __syncwarp();
        dinv[warp * 32 + lane] = 1.0 / Ls[lane * 33 + lane];
        __syncwarp();
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) sacc = fma(-Ls[i * 33 + k], x[k], sacc);     // broadcast reads of row i
            x[i] = sacc * dinv[warp * 32 + i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) Ls[i * 33 + lane] = x[i];                        // column `lane` of the inverse
    }
    __syncthreads();
    // Both sweeps are chains of block steps, each: off-diagonal mat-vec, barrier, inverse applied by warp 0, barrier.
    // The entries of L a step multiplies do not depend on the solve, only the vector does: they are loaded into
    // registers one step AHEAD, right after the previous step has consumed its own, so that the L2 / HBM round trip
    // runs behind the reduction, the barriers and the inverse instead of in front of every step (ncu, round 2: 26 % of
    // the stall samples on the load of the forward mat-vec, 24 % on the barriers behind it).  Same products summed in
    // the same order: bitwise the results of the unpipelined sweeps.
    // ---- forward
    constexpr int FW_J = TRSVI_MAX_BLK - 1;                        // k-chunks of 32 a block row can have
    double lv[4][FW_J];
    auto load_rows = [&](int I) {                                  // block row I: rows i0 + warp + 8 q, columns lane + 32 j
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int rr = warp + q * TRSVB_NW;
            const double* row = L + (size_t)(i0 + rr) * ldm + lane;
#pragma unroll
            for (int j = 0; j < FW_J; ++j) lv[q][j] = (j < I && rr < nb) ? row[32 * j] : 0.0;
        }
    };
    if (nblk > 1) load_rows(1);
    for (int I = 0; I < nblk; ++I) {
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
        if (i0 > 0) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < FW_J; ++j) {
                if (j < I) {
                    const double zk = vec[lane + 32 * j];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int rr = warp + q * TRSVB_NW;
                        if (rr < nb) acc[q] += lv[q][j] * zk;
                    }
                }
            }
            if (I + 1 < nblk) load_rows(I + 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double sum = warp_sum(acc[q]);
                const int rr = warp + q * TRSVB_NW;
                if (lane == 0 && rr < nb) vec[i0 + rr] -= sum;
            }
            __syncthreads();
        }
        if (warp == 0) {
            const double* Li = Linv + I * 32 * 33 + lane * 33;          // row `lane` of the inverse
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                acc0 += Li[j] * ((j < nb) ? vec[i0 + j] : 0.0);
                acc1 += Li[j + 1] * ((j + 1 < nb) ? vec[i0 + j + 1] : 0.0);
            }
            __syncwarp();
            if (lane < nb) vec[i0 + lane] = acc0 + acc1;
        }
        __syncthreads();
    }
    // ---- backward (m <= 256: thread k owns column k of the block row being eliminated)
    double cv[32];
    auto load_cols = [&](int I) {                                  // block row I, column tid: 32 entries
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
        // (only threads tid < i0 use cv, and entries i >= nb are never read: no zero fill, no per-entry predicate on the
        // full blocks, one pointer increment per load - the ternary per entry was 55 % of the kernel's instructions)
        if (tid < i0) {
            const double* col = L + (size_t)i0 * ldm + tid;
            if (nb == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { cv[i] = *col; col += ldm; }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) { cv[i] = (i < nb) ? *col : 0.0; col += ldm; }
            }
        }
    };
    if (nblk > 1) load_cols(nblk - 1);
    for (int I = nblk - 1; I >= 0; --I) {
        const int i0 = I << 5;
        const int nb = (m - i0 < 32) ? (m - i0) : 32;
        if (warp == 0) {
            const double* Lc = Linv + I * 32 * 33 + lane;               // column `lane` of the inverse
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                acc0 += Lc[i * 33] * ((i < nb) ? vec[i0 + i] : 0.0);
                acc1 += Lc[(i + 1) * 33] * ((i + 1 < nb) ? vec[i0 + i + 1] : 0.0);
            }
            __syncwarp();
            if (lane < nb) vec[i0 + lane] = acc0 + acc1;
        }
        __syncthreads();
        if (tid < i0) {
            double acc = 0.0;
            if (nb == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += cv[i] * vec[i0 + i];
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nb) acc += cv[i] * vec[i0 + i];          // (static indices: cv stays in registers)
            }
            vec[tid] -= acc;
        }
        if (I - 1 >= 1) load_cols(I - 1);
        __syncthreads();
    }
    double* dst = a.out ? a.out + (size_t)bz * a.strideV : v;
    if (a.accumulate) {
        for (int i = tid; i < m; i += TRSVB_NT) dst[i] = dst[i] + vec[i];
    } else {
        for (int i = tid; i < m; i += TRSVB_NT) dst[i] = vec[i];
    }
}
// (two CTAs per SM: the prefetch buffers need the registers; three were no faster, IPM_TRSV_SMEM_KB A/B)
static __global__ void __launch_bounds__(TRSVB_NT, 2) k_trsv_batched_inv(const TrsvBatchedArgs a) { d_trsv_batched_inv(a); }

inline size_t trsv_batched_smem(int m) { return (size_t)(2 * 32 * 33 + m) * sizeof(double); }

constexpr int TRSV_ONE_CTA_MAX_M = 2048;
// rhs (destroyed) -> sol.  Orders up to 2048 run both sweeps in ONE CTA (no launch per block: the small and
// mid-size Netlib LPs are launch-bound); larger ones use one launch per 64-wide block over all SMs.
inline int potrs_single(const double* L, int64_t ldm, int m, double* rhs, double* tmp, double* sol, cudaStream_t st) {
    if (m <= TRSV_ONE_CTA_MAX_M) {
        TrsvBatchedArgs t;
        t.L = L; t.ldm = ldm; t.strideM = 0; t.v = rhs; t.strideV = 0; t.m = m; t.active = nullptr; t.out = sol;
        if (m <= 32 * TRSVI_MAX_BLK) {
            IPM_TRY(ensure_dyn_smem(k_trsv_batched_inv, trsv_batched_inv_smem(32 * TRSVI_MAX_BLK)));
            k_trsv_batched_inv<<<1, TRSVB_NT, trsv_batched_inv_smem(m), st>>>(t);
        } else
        k_trsv_batched<<<1, TRSVB_NT, trsv_batched_smem(m), st>>>(t);
        count_launch();
        return launch_check();
    }
    return potrs_single_blocks(L, ldm, m, rhs, tmp, sol, st);
}
#endif

}  // namespace ipm
