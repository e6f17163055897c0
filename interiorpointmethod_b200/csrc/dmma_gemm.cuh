// FP64 tensor-core (DMMA.8x8x4) "NT" product used for
//   * the dense normal-equations SYRK  M = A diag(d) A^T        (main.py:224)
//   * the trailing update of the blocked Cholesky  C -= P P^T   (replaces solve_linear, main.py:176-182)
//   * the batched versions of both (blockIdx.z = LP index).
//
//   C[i][j] (op)= sum_k P[i][k] * Q[j][k] * (dvec ? dvec[k] : 1)      i < rowsP, j < rowsQ, k < K
//
// Both operands are row-major with the contraction index contiguous, which is exactly how A (rows of the
// constraint matrix) and the Cholesky panel (rows of L) are stored, so no transposes are materialised.
// tcgen05 has no f64 kind (ptxas rejects kind::f64), so the tensor pipe is reached through warp-level
// mma.sync m8n8k4; operands are staged global -> registers -> shared (the diag(d) scaling is applied on the
// way into shared memory) with a two-stage register/shared double buffer.
#pragma once
#include "common.cuh"

namespace ipm {

struct DmmaArgs {
    const double* P;  int64_t ldp;  int64_t strideP;
    const double* Q;  int64_t ldq;  int64_t strideQ;
    const double* dvec;             int64_t strideD;   // nullable
    double* C;        int64_t ldc;  int64_t strideC;
    int rowsP, rowsQ, K;
    int lower_only;        // skip tiles strictly above the block diagonal (square tiles, P/Q share row space)
    const int* active;     // nullable: per-batch flag, 0 => skip this LP
    int col0_only = 0;     // warp-specialised kernel only: compute just the first tile column (tiles (bi, 0)) - the
                           // part of a Cholesky trailing update the next panel needs (look-ahead, chol.cuh)
    int max_ctas = 0;      // warp-specialised kernel only: cap of the persistent grid (0 = one CTA per SM)
    // warp-specialised kernel, batched IPM only (needs dvec): the diagonal tiles also form the predictor right-hand side
    //   rhs[z][r] = -rbvec[z][r] - sum_k P[r][k] * dvec[z][k] * vvec[z][k]          (main.py:225 with w = d * v)
    // from the operand slabs that are in shared memory anyway - one pass over A less per iteration.
    const double* vvec = nullptr;   int64_t strideV = 0;     // [batch][K]
    const double* rbvec = nullptr;  double* rhs = nullptr;  int64_t strideR = 0;   // [batch][rowsP]
};

constexpr int DMMA_BK = 16;
constexpr int DMMA_LD = DMMA_BK + 4;   // padded row (doubles): conflict-free 64-bit fragment loads

template <int BM, int BN, int WM, int WN>
constexpr size_t dmma_smem_bytes() { return (size_t)2 * (BM + BN) * DMMA_LD * sizeof(double); }

#ifdef __CUDACC__
// EPI: 0 -> C = acc ; 1 -> C = C - acc
template <int BM, int BN, int WM, int WN, int EPI>
__global__ void __launch_bounds__(WM * WN * 32, 1) dmma_nt_kernel(const DmmaArgs a) {
    constexpr int NT = WM * WN * 32;
    constexpr int TM = BM / WM, TN = BN / WN;     // warp tile
    constexpr int MI = TM / 8, NI = TN / 8;       // 8x8 sub-tiles per warp
    constexpr int LD = DMMA_LD;
    constexpr int PV = BM * DMMA_BK / 2 / NT;     // double2 loads per thread per tile (P)
    constexpr int QV = BN * DMMA_BK / 2 / NT;
    static_assert(BM * DMMA_BK / 2 % NT == 0 && BN * DMMA_BK / 2 % NT == 0, "tile/threads mismatch");
    static_assert(NT % 8 == 0, "k-pair index must be loop invariant");

    const int bi = blockIdx.y, bj = blockIdx.x, bz = blockIdx.z;
    if (a.lower_only && bj * BN > bi * BM + (BM - 1)) return;
    if (a.active && a.active[bz] == 0) return;

    extern __shared__ __align__(16) double smem[];
    double* Ps = smem;                         // [2][BM][LD]
    double* Qs = smem + 2 * BM * LD;           // [2][BN][LD]

    const double* __restrict__ P = a.P + (size_t)bz * a.strideP;
    const double* __restrict__ Q = a.Q + (size_t)bz * a.strideQ;
    const double* __restrict__ dv = a.dvec ? a.dvec + (size_t)bz * a.strideD : nullptr;
    double* __restrict__ C = a.C + (size_t)bz * a.strideC;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WN) * TM, wn0 = (warp % WN) * TN;
    const int row0 = bi * BM, col0 = bj * BN;
    const int K = a.K;
    const int nk = (K + DMMA_BK - 1) / DMMA_BK;

    // loader mapping: 8 threads cover one 16-double row segment (128 B)
    const int lrow = tid >> 3;            // + i * (NT/8)
    const int lk = (tid & 7) * 2;
    const bool vecP = ((a.ldp & 1) == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0);
    const bool vecQ = ((a.ldq & 1) == 0) && ((reinterpret_cast<uintptr_t>(Q) & 15) == 0);

    double2 pr[PV], qr[QV];
    double2 sc = make_double2(1.0, 1.0);   // diag(d) entries of the tile in flight

    auto load_tile = [&](int kt) {
        const int k = kt * DMMA_BK + lk;
#pragma unroll
        for (int i = 0; i < PV; ++i) {
            const int r = row0 + lrow + i * (NT / 8);
            double2 v = make_double2(0.0, 0.0);
            if (r < a.rowsP) {
                const double* p = P + (size_t)r * a.ldp + k;
                if (k + 1 < K) {
                    if (vecP) v = *reinterpret_cast<const double2*>(p);
                    else { v.x = p[0]; v.y = p[1]; }
                } else if (k < K) v.x = p[0];
            }
            pr[i] = v;
        }
        if (dv) {
            sc = make_double2(0.0, 0.0);
            if (k < K) sc.x = dv[k];
            if (k + 1 < K) sc.y = dv[k + 1];
        }
#pragma unroll
        for (int i = 0; i < QV; ++i) {
            const int r = col0 + lrow + i * (NT / 8);
            double2 v = make_double2(0.0, 0.0);
            if (r < a.rowsQ) {
                const double* p = Q + (size_t)r * a.ldq + k;
                if (k + 1 < K) {
                    if (vecQ) v = *reinterpret_cast<const double2*>(p);
                    else { v.x = p[0]; v.y = p[1]; }
                } else if (k < K) v.x = p[0];
            }
            qr[i] = v;
        }
    };
    auto store_tile = [&](int buf) {
        double* ps = Ps + buf * BM * LD;
        double* qs = Qs + buf * BN * LD;
#pragma unroll
        for (int i = 0; i < PV; ++i)
            *reinterpret_cast<double2*>(ps + (lrow + i * (NT / 8)) * LD + lk) = pr[i];
        // the diag(d) scaling is applied here, AFTER the MMAs of the previous tile, so the global loads issued
        // by load_tile stay in flight during the tensor work instead of being waited on immediately
#pragma unroll
        for (int i = 0; i < QV; ++i) {
            double2 v = qr[i];
            if (dv) { v.x *= sc.x; v.y *= sc.y; }
            *reinterpret_cast<double2*>(qs + (lrow + i * (NT / 8)) * LD + lk) = v;
        }
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    load_tile(0);
    store_tile(0);
    __syncthreads();

    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile(kt + 1);          // global loads in flight during the MMAs
        const double* ps = Ps + buf * BM * LD + (wm0 + g) * LD + t;
        const double* qs = Qs + buf * BN * LD + (wn0 + g) * LD + t;
#pragma unroll
        for (int kk = 0; kk < DMMA_BK; kk += 4) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = ps[i * 8 * LD + kk];
#pragma unroll
            for (int j = 0; j < NI; ++j) bf[j] = qs[j * 8 * LD + kk];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (kt + 1 < nk) store_tile(buf ^ 1);
        __syncthreads();
    }

    // epilogue
    const bool vecC = ((a.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = row0 + wm0 + i * 8 + g;
        if (r >= a.rowsP) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int c = col0 + wn0 + j * 8 + 2 * t;
            if (c >= a.rowsQ) continue;
            double* cp = C + (size_t)r * a.ldc + c;
            if (c + 1 < a.rowsQ && vecC) {
                double2 v;
                if (EPI == 1) {
                    v = *reinterpret_cast<double2*>(cp);
                    v.x -= acc[i][j][0];
                    v.y -= acc[i][j][1];
                } else {
                    v = make_double2(acc[i][j][0], acc[i][j][1]);
                }
                *reinterpret_cast<double2*>(cp) = v;
            } else {
                if (EPI == 1) {
                    cp[0] -= acc[i][j][0];
                    if (c + 1 < a.rowsQ) cp[1] -= acc[i][j][1];
                } else {
                    cp[0] = acc[i][j][0];
                    if (c + 1 < a.rowsQ) cp[1] = acc[i][j][1];
                }
            }
        }
    }
}

// Host launcher.  grid = (tiles over rowsQ, tiles over rowsP, batch).
template <int BM, int BN, int WM, int WN, int EPI>
inline int dmma_nt_launch(const DmmaArgs& a, int batch, cudaStream_t st) {
    constexpr size_t smem = dmma_smem_bytes<BM, BN, WM, WN>();
    auto kern = dmma_nt_kernel<BM, BN, WM, WN, EPI>;
    IPM_TRY(ensure_dyn_smem(kern, smem));
    if (a.rowsP <= 0 || a.rowsQ <= 0 || batch <= 0) return IPM_OK;
    dim3 grid(ceil_div(a.rowsQ, BN), ceil_div(a.rowsP, BM), batch);
    kern<<<grid, WM * WN * 32, smem, st>>>(a);
    count_launch();
    return launch_check();
}
#endif

}  // namespace ipm
