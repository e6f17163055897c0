// Pipelined triangular solves for one large factor (m > 256): replaces the back-substitution half of
// solve_linear (main.py:176-182) on the Cholesky factor of main.py:223-224.
//
// The per-block-launch version (k_trsv_fwd128 / k_trsv_bwd128, one launch per 128 columns) spends ~50 us per
// block on a chain of [launch, load the 128x128 diagonal block, 128 serial substitution steps, update]; on QAP15
// (m = 6330, 50 blocks, four sweeps per Newton iteration) that was 45 % of the iteration while the lower triangle
// (160 MB) streams in 27 us.  Here
//   * k_trinv128 inverts the 128x128 diagonal blocks once per factorisation (one CTA per block, all in parallel),
//     so a diagonal solve becomes a 128x128 mat-vec;
//   * k_trsv_pipe_fwd / k_trsv_pipe_bwd are single persistent launches, one CTA per block row (forward) / block
//     column (backward), at most one CTA per SM.  The owner of block I accumulates L[I,J] z_J for J = 0, 1, ... as
//     the z_J appear (one release/acquire flag per block in global memory), then applies the inverse and publishes
//     z_I.  The critical chain per block is two 128x128 mat-vecs and one flag hop instead of a kernel launch.
// All CTAs of a launch are co-resident (grid <= number of SMs, 1 CTA per SM), and every CTA walks its blocks in the
// order of the sweep, so the block with the smallest unfinished index can always make progress: no deadlock.
#pragma once
#include "common.cuh"

namespace ipm {

constexpr int TP_NB = 128;            // block size
constexpr int TP_NT = 256;            // threads per CTA: 8 warps x 16 rows, lanes over 4 columns each
constexpr int TP_PACK = TP_NB * (TP_NB + 1) / 2;

struct TrsvPipeWs {
    double* Linv = nullptr;           // [nblk][128][128] inverses of the diagonal blocks (lower, upper part zero)
    int* flags = nullptr;             // [2][nblk]: forward / backward "block solved" flags
    int nblk = 0;
};

inline size_t trinv_smem() { return (size_t)(2 * TP_PACK + TP_NB) * sizeof(double); }

#ifdef __CUDACC__
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Inverse of the 128x128 lower-triangular diagonal block `blockIdx.x` of L.  Thread j owns column j of the inverse:
// x_j = 1/L_jj, x_i = -(sum_{k=j}^{i-1} L_ik x_k) / L_ii.  Both triangles live packed in shared memory; the row of
// L is read at the same address by neighbouring threads (broadcast) and the column of X at consecutive addresses.
static __global__ void __launch_bounds__(TP_NB, 1) k_trinv128(const double* L, int64_t ldm, int m, double* Linv) {
    extern __shared__ __align__(16) double sm_ti[];
    double* Ls = sm_ti;                 // packed lower triangle of the block, row i at i(i+1)/2
    double* Xs = Ls + TP_PACK;          // packed lower triangle of the inverse
    double* dinv = Xs + TP_PACK;        // 1 / L_ii
    const int blk = blockIdx.x, j = threadIdx.x;
    const int i0 = blk * TP_NB;
    const int nb = (m - i0 < TP_NB) ? (m - i0) : TP_NB;
    const double* Ld = L + (size_t)i0 * ldm + i0;
    for (int idx = j; idx < TP_NB * TP_NB; idx += TP_NB) {
        const int r = idx >> 7, c = idx & (TP_NB - 1);
        if (c <= r) Ls[r * (r + 1) / 2 + c] = (r < nb) ? Ld[(size_t)r * ldm + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    dinv[j] = 1.0 / Ls[j * (j + 1) / 2 + j];
    __syncthreads();
    Xs[j * (j + 1) / 2 + j] = dinv[j];
    for (int i = j + 1; i < TP_NB; ++i) {
        const double* lrow = Ls + i * (i + 1) / 2;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = j;
        for (; k + 3 < i; k += 4) {
            a0 = fma(lrow[k], Xs[k * (k + 1) / 2 + j], a0);
            a1 = fma(lrow[k + 1], Xs[(k + 1) * (k + 2) / 2 + j], a1);
            a2 = fma(lrow[k + 2], Xs[(k + 2) * (k + 3) / 2 + j], a2);
            a3 = fma(lrow[k + 3], Xs[(k + 3) * (k + 4) / 2 + j], a3);
        }
        for (; k < i; ++k) a0 = fma(lrow[k], Xs[k * (k + 1) / 2 + j], a0);
        Xs[i * (i + 1) / 2 + j] = -((a0 + a1) + (a2 + a3)) * dinv[i];
    }
    __syncthreads();
    double* out = Linv + (size_t)blk * TP_NB * TP_NB;
    for (int idx = j; idx < TP_NB * TP_NB; idx += TP_NB) {
        const int r = idx >> 7, c = idx & (TP_NB - 1);
        out[idx] = (c <= r) ? Xs[r * (r + 1) / 2 + c] : 0.0;
    }
}

struct TrsvPipeArgs {
    const double* L; int64_t ldm; int m;
    const double* Linv;
    const double* rhs;        // right-hand side of this sweep
    double* out;              // solution of this sweep (read by the other CTAs once the flag is up)
    int* flags;               // [nblk], zero on entry
    int nblk;
};

// sum over the 32 lanes of 16 per-lane partials: lane 0 ends up with all 16 sums in v[]
__device__ __forceinline__ void tp_reduce16(double (&v)[16]) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[r] += __shfl_xor_sync(0xffffffffu, v[r], o);
    }
}

// y[row] (16 rows of this warp) += B[row][0:128] . x[0:128] for a 128x128 row-major block B with leading dimension ld;
// lane l covers columns 4l..4l+3 (x4 holds x there), rows beyond `nrows` and columns beyond `ncols` do not exist.
__device__ __forceinline__ void tp_rows_accumulate(double (&acc)[16], const double* B, int64_t ld, int row0, int nrows,
                                                   int ncols, int lane, const double (&x4)[4]) {
    const int c = 4 * lane;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double2 v0[8], v1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = row0 + h * 8 + q;
            const bool ok = r < nrows;
            const double* p = B + (size_t)(ok ? r : 0) * ld + c;
            v0[q] = (ok && c + 1 < ncols) ? *reinterpret_cast<const double2*>(p)
                                          : make_double2((ok && c < ncols) ? p[0] : 0.0, 0.0);
            v1[q] = (ok && c + 3 < ncols) ? *reinterpret_cast<const double2*>(p + 2)
                                          : make_double2((ok && c + 2 < ncols) ? p[2] : 0.0, 0.0);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double s = acc[h * 8 + q];
            s = fma(v0[q].x, x4[0], s);
            s = fma(v0[q].y, x4[1], s);
            s = fma(v1[q].x, x4[2], s);
            s = fma(v1[q].y, x4[3], s);
            acc[h * 8 + q] = s;
        }
    }
}

static __global__ void __launch_bounds__(TP_NT, 1) k_trsv_pipe_fwd(const TrsvPipeArgs a) {
    __shared__ double rs[TP_NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = warp * 16;
    for (int I = blockIdx.x; I < a.nblk; I += gridDim.x) {
        const int i0 = I * TP_NB;
        const int nrows = (a.m - i0 < TP_NB) ? (a.m - i0) : TP_NB;
        double acc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] = 0.0;
        // ---- r_I - sum_{J<I} L[I,J] z_J, consuming the z_J as they are published
        for (int J = 0; J < I; ++J) {
            while (ld_acquire(a.flags + J) == 0) {}
            double x4[4];
            const double2 z0 = __ldcg(reinterpret_cast<const double2*>(a.out + J * TP_NB + 4 * lane));      // L2: written by another SM
            const double2 z1 = __ldcg(reinterpret_cast<const double2*>(a.out + J * TP_NB + 4 * lane + 2));
            x4[0] = z0.x; x4[1] = z0.y; x4[2] = z1.x; x4[3] = z1.y;
            tp_rows_accumulate(acc, a.L + (size_t)i0 * a.ldm + (size_t)J * TP_NB, a.ldm, row0, nrows, TP_NB, lane, x4);
        }
        tp_reduce16(acc);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int row = row0 + r;
                rs[row] = (row < nrows) ? a.rhs[i0 + row] - acc[r] : 0.0;
            }
        }
        __syncthreads();
        // ---- z_I = Linv_I r_I
        {
            double x4[4], acc2[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) x4[q] = rs[4 * lane + q];
#pragma unroll
            for (int r = 0; r < 16; ++r) acc2[r] = 0.0;
            tp_rows_accumulate(acc2, a.Linv + (size_t)I * TP_NB * TP_NB, TP_NB, row0, TP_NB, TP_NB, lane, x4);
            tp_reduce16(acc2);
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (row0 + r < nrows) a.out[i0 + row0 + r] = acc2[r];
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(a.flags + I, 1);
    }
}

// y_I = Linv_I^T (z_I - sum_{J>I} L[J,I]^T y_J): column sums, lane l covers columns 4l..4l+3 of block column I, the
// eight warps split the 128 rows of every L[J,I] and are combined through shared memory at the end.
static __global__ void __launch_bounds__(TP_NT, 1) k_trsv_pipe_bwd(const TrsvPipeArgs a) {
    __shared__ double part[TP_NT / 32][TP_NB];
    __shared__ double ts[TP_NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = warp * 16;
    const int last = a.nblk - 1;
    for (int I = last - (int)blockIdx.x; I >= 0; I -= gridDim.x) {
        const int i0 = I * TP_NB;
        const int ncols = (a.m - i0 < TP_NB) ? (a.m - i0) : TP_NB;
        const int c = 4 * lane;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        auto cols_accumulate = [&](const double* B, int64_t ld, int nrows, const volatile double* yv) {
            // acc[0..3] += sum over this warp's 16 rows of B[row][c..c+3] * y[row]
            double yr = (lane < 16 && row0 + lane < nrows) ? yv[row0 + lane] : 0.0;      // volatile: see the call sites
            double2 v0[16], v1[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int r = row0 + q;
                const bool ok = r < nrows;
                const double* p = B + (size_t)(ok ? r : 0) * ld + c;
                v0[q] = (ok && c + 1 < ncols) ? *reinterpret_cast<const double2*>(p)
                                              : make_double2((ok && c < ncols) ? p[0] : 0.0, 0.0);
                v1[q] = (ok && c + 3 < ncols) ? *reinterpret_cast<const double2*>(p + 2)
                                              : make_double2((ok && c + 2 < ncols) ? p[2] : 0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const double y = __shfl_sync(0xffffffffu, yr, q);
                acc[0] = fma(v0[q].x, y, acc[0]);
                acc[1] = fma(v0[q].y, y, acc[1]);
                acc[2] = fma(v1[q].x, y, acc[2]);
                acc[3] = fma(v1[q].y, y, acc[3]);
            }
        };
        for (int J = last; J > I; --J) {
            while (ld_acquire(a.flags + J) == 0) {}
            const int j0 = J * TP_NB;
            const int nrows = (a.m - j0 < TP_NB) ? (a.m - j0) : TP_NB;
            cols_accumulate(a.L + (size_t)j0 * a.ldm + i0, a.ldm, nrows, a.out + j0);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) part[warp][c + q] = acc[q];
        __syncthreads();
        if (tid < TP_NB) {
            double u = 0.0;
#pragma unroll
            for (int w = 0; w < TP_NT / 32; ++w) u += part[w][tid];
            ts[tid] = (tid < ncols) ? a.rhs[i0 + tid] - u : 0.0;
        }
        __syncthreads();
        // ---- y_I = Linv_I^T t
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = 0.0;
        cols_accumulate(a.Linv + (size_t)I * TP_NB * TP_NB, TP_NB, TP_NB, ts);
        __syncthreads();                       // part is reused
#pragma unroll
        for (int q = 0; q < 4; ++q) part[warp][c + q] = acc[q];
        __syncthreads();
        if (tid < ncols) {
            double y = 0.0;
#pragma unroll
            for (int w = 0; w < TP_NT / 32; ++w) y += part[w][tid];
            a.out[i0 + tid] = y;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(a.flags + I, 1);
    }
}

// inverses of all diagonal blocks of the factor (after potrf)
inline int trinv_blocks(const double* L, int64_t ldm, int m, const TrsvPipeWs& ws, cudaStream_t st) {
    IPM_TRY(ensure_dyn_smem(k_trinv128, trinv_smem()));
    k_trinv128<<<ws.nblk, TP_NB, trinv_smem(), st>>>(L, ldm, m, ws.Linv);
    count_launch();
    return launch_check();
}

// rhs -> sol (tmp holds the forward solution); needs trinv_blocks for the current factor
inline int potrs_pipe(const double* L, int64_t ldm, int m, const TrsvPipeWs& ws, const double* rhs, double* tmp,
                      double* sol, cudaStream_t st) {
    IPM_CUDA_OK(cudaMemsetAsync(ws.flags, 0, (size_t)2 * ws.nblk * sizeof(int), st));
    TrsvPipeArgs a;
    a.L = L; a.ldm = ldm; a.m = m; a.Linv = ws.Linv; a.nblk = ws.nblk;
    const int grid = ws.nblk < kNumSMs ? ws.nblk : kNumSMs;
    a.rhs = rhs; a.out = tmp; a.flags = ws.flags;
    k_trsv_pipe_fwd<<<grid, TP_NT, 0, st>>>(a);
    a.rhs = tmp; a.out = sol; a.flags = ws.flags + ws.nblk;
    k_trsv_pipe_bwd<<<grid, TP_NT, 0, st>>>(a);
    count_launch(2);
    return launch_check();
}
#endif

}  // namespace ipm
