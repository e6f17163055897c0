// Dense predictor-corrector iteration on the AUGMENTED system, one CTA per LP: the reference's dense path
// (`create_matrix` main.py:13-21 + `np.linalg.solve` main.py:178 = LAPACK dgesv on the unreduced (m+2n) KKT matrix)
// with ds eliminated exactly (ds = -s dx/x - rcomp/x), i.e. order n + m:
//        [ -D^-1  A^T ] [dx]   [ -(rc - rcomp/x) ]
//        [   A     0  ] [dy] = [       -rb        ]          D = diag(x/s)
// factored by LU with PARTIAL PIVOTING like dgesv.  It never pivots on the tiny s_j/x_j of the basic variables, so it
// keeps the primal block row A dx = -rb to working precision where the normal equations M = A D A^T cannot: on a
// (nearly) degenerate vertex M is numerically singular once max d / min d passes 1e19, the safeguarded Cholesky
// drops a row that is NOT dependent, and an LP that has not met check_optimality (main.py:169-173) by then never
// will (generator LPs 16893, 31186, 54456 ...: thousands of iterations, the reference needs 17-18).
//
// Used (a) by the batched solver as the hand-off target for exactly those LPs (ipm_batched.cu: the corrector pass
// detects |(-rb - A dx)| > |rb| after a refinement step and parks the LP; this kernel continues it from its current
// iterate), about one LP in a thousand, and (b) stand-alone through ipm_solve_dense_kkt (parity tests against the
// reference's goldens: same Newton system, same pivoting rule => same iteration counts).
// The tests compare it with a CPU restatement of the same formulas (`direction_augmented` in the test infrastructure).
//
// One thread-block CLUSTER of KA_CL CTAs (1024 threads each) per LP; the matrix K (order N = n + m) lives in global
// memory (L2 resident: 4.7 MB at 256 x 512), panels of 16 columns are factored in shared memory.  Every CTA of the
// cluster runs the WHOLE iteration redundantly - residuals, panel factorisations, triangular solves, ratio tests are
// cheap and their results bitwise identical, so no data has to be exchanged for them - and only the work that scales
// with N^2 per panel is divided: rows swaps, the U12 solve and the trailing update are split by 32-column chunks
// over the CTAs (round-robin), with one cluster barrier per panel.  CTA 0 alone writes the LP's state.  It is a
// robustness path, not a throughput path: (2/3) N^3 flop per iteration on KA_CL SMs (5 ms per iteration at 256 x 512 on
// one SM - measured, round 2 - which is what the cluster is for: an LP parked in the last iterations of a solve
// finishes after everybody else).
#pragma once
#include <cstring>

#include "common.cuh"

namespace ipm {

// Measured (tools/kkt_profile.py, 256 x 512, round 2): the first version - 1024 threads, panel rows 16 doubles apart,
// 8-row batches - spent 3.7 ms per iteration in the panel's column steps (every access to one panel column by
// consecutive rows was a 32-way bank conflict) and 6.1 ms in the trailing update (under the 64-register cap of a
// 1024-thread CTA the compiler sank the batched loads next to their uses: one exposed L2 round trip per row).  Hence:
// an odd row stride for the panel, 512 threads (128 registers) and 16-row batches.
constexpr int KA_NT = 512;
constexpr int KA_NW = KA_NT / 32;
constexpr int KA_PW = 16;                      // panel width
constexpr int KA_LDP = KA_PW + 1;              // panel row stride in shared memory (odd: conflict-free columns)
constexpr int KA_RB = 16;                      // rows per batch of the trailing update
constexpr int KA_CL_MAX = 8;                   // largest cluster (portable limit)
constexpr int KA_MAX_N = 1600;                 // order n + m the panel buffer admits (218 KB)

struct KktArgs {
    const double* A;        // [B][m][n]
    const double* b;        // [B][m]
    const double* c;        // [B][n]
    double *x, *s;          // [B][n]   iterate, continued in place
    double* y;              // [B][m]
    double* scal;           // [B][S_COUNT]  S_NB, S_NC in; norms, objective, S_CONT out
    int* iters;             // [B]           continued
    const int* list;        // [count] LP indices to process (nullptr: LP = blockIdx.x)
    double* work;           // per CTA: ka_work_doubles(m, n)
    int m, n;
    double tol, eta;
    int max_iter;
    long long* prof = nullptr;   // optional [16]: cycles per phase of cluster 0's CTA 0 (ipm_kkt_last_profile)
};

#ifdef __CUDACC__
#define KA_HD __host__ __device__
#else
#define KA_HD
#endif
KA_HD inline int64_t ka_ldk(int m, int n) { return ((int64_t)m + n + 1) / 2 * 2; }
KA_HD inline int64_t ka_work_doubles(int m, int n) {
    const int64_t N = (int64_t)m + n;
    // K once per LP; piv, rb and six n-vectors once per CTA of the cluster (each CTA keeps its own redundant copy)
    const int64_t v = N * ka_ldk(m, n) + (int64_t)KA_CL_MAX * (N + (int64_t)m + 6 * (int64_t)n);
    return (v + 15) / 16 * 16;
}
inline size_t ka_smem_bytes(int m, int n) {
    const size_t N = (size_t)m + n;
    return (N * KA_LDP + 32 * 33 + 64) * sizeof(double);
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned ka_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned ka_cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// barrier over all threads of all CTAs of the cluster; release/acquire at cluster scope orders their global writes
__device__ __forceinline__ void ka_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// argmax |v| with the lowest index on ties (LAPACK idamax): result valid in ALL threads.
__device__ __forceinline__ void ka_block_argmax(double v, int idx, double* shv, int* shi, double& vout, int& iout) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { shv[w] = v; shi[w] = idx; }
    __syncthreads();
    v = (lane < KA_NW) ? shv[lane] : -1.0;
    idx = (lane < KA_NW) ? shi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    vout = v; iout = idx;
}

// block sum / min, result in ALL threads
template <int OP>
__device__ __forceinline__ double ka_block_all(double v, double* sh, double* bc) {
    v = block_red<OP>(v, sh);
    if (threadIdx.x == 0) *bc = v;
    __syncthreads();
    v = *bc;
    __syncthreads();
    return v;
}

// In-place LU with partial pivoting of K (N x N, row-major, leading dimension ld), blocked by KA_PW columns.
// piv[k] = row swapped with row k at step k (LAPACK ipiv, 0-based).
// crank / csize: this CTA's rank in the cluster and the cluster's size (see the header comment).
#define KA_T(slot) do { if (prof && threadIdx.x == 0) { const long long _t = clock64(); prof[slot] += _t - t_last; t_last = _t; } } while (0)
__device__ void ka_lu_factor(double* K, int N, int64_t ld, int* piv, double* Psm, double* shv, int* shi, int crank,
                             int csize, long long* prof) {
    const int tid = threadIdx.x;
    long long t_last = clock64();
    auto mine = [&](int cc) { return ((cc >> 5) % csize) == crank; };      // 32-column chunks, round-robin
    for (int k0 = 0; k0 < N; k0 += KA_PW) {
        const int pw = (N - k0 < KA_PW) ? (N - k0) : KA_PW;
        const int rows = N - k0;
        // (1) panel -> shared memory
        for (int idx = tid; idx < rows * KA_PW; idx += KA_NT) {
            const int r = idx / KA_PW, cc = idx - r * KA_PW;
            Psm[r * KA_LDP + cc] = (cc < pw) ? K[(size_t)(k0 + r) * ld + k0 + cc] : 0.0;
        }
        __syncthreads();
        KA_T(2);
        // (2) unblocked LU of the panel
        for (int j = 0; j < pw; ++j) {
            double best = -1.0;
            int bi = 0x7fffffff;
            for (int r = j + tid; r < rows; r += KA_NT) {
                const double av = fabs(Psm[r * KA_LDP + j]);
                if (av > best) { best = av; bi = r; }       // first maximum, like idamax; NaN never compares greater
            }
            double bv; int br;
            ka_block_argmax(best, bi, shv, shi, bv, br);
            if (br == 0x7fffffff) br = j;                   // all NaN / empty
            if (tid == 0) piv[k0 + j] = k0 + br;
            if (br != j && tid < KA_PW) {
                const double t0 = Psm[j * KA_LDP + tid];
                Psm[j * KA_LDP + tid] = Psm[br * KA_LDP + tid];
                Psm[br * KA_LDP + tid] = t0;
            }
            __syncthreads();
            const double p = Psm[j * KA_LDP + j];
            for (int r = j + 1 + tid; r < rows; r += KA_NT) {
                double* pr = Psm + r * KA_LDP;
                const double l = pr[j] / p;
                pr[j] = l;
                for (int cc = j + 1; cc < pw; ++cc) pr[cc] = fma(-l, Psm[j * KA_LDP + cc], pr[cc]);
            }
            __syncthreads();
        }
        KA_T(3);
        // (3) panel back to K (every CTA holds the same panel: CTA 0 writes it)
        if (crank == 0) {
            for (int idx = tid; idx < rows * KA_PW; idx += KA_NT) {
                const int r = idx / KA_PW, cc = idx - r * KA_PW;
                if (cc < pw) K[(size_t)(k0 + r) * ld + k0 + cc] = Psm[r * KA_LDP + cc];
            }
        }
        KA_T(4);
        // (4) the panel's row interchanges on the columns outside the panel, in order
        for (int j = 0; j < pw; ++j) {
            const int r1 = k0 + j, r2 = piv[k0 + j];        // written by thread 0 before a barrier above
            if (r1 != r2) {
                for (int cc = tid; cc < N; cc += KA_NT) {
                    if ((cc >= k0 && cc < k0 + pw) || !mine(cc)) continue;
                    const double t0 = K[(size_t)r1 * ld + cc];
                    K[(size_t)r1 * ld + cc] = K[(size_t)r2 * ld + cc];
                    K[(size_t)r2 * ld + cc] = t0;
                }
            }
            __syncthreads();
        }
        KA_T(5);
        // (5) U12 = L11^-1 K12 and the trailing update K22 -= L21 U12.  A work item is (column, row slab): the column's 16
        // entries of U12 live in registers, the rows of the slab are walked in batches of KA_RB whose loads are issued
        // together.  This CTA owns the 32-column chunks cq with cq % csize == crank; when it has fewer columns than
        // threads (late panels, or a cluster), the rows of a column are cut into slabs so that the idle threads take a
        // share of every column's dependent chain of L2 round trips.
        const int c0 = k0 + pw;
        const int cq0 = c0 >> 5;                                         // first chunk that still has columns >= c0
        const int cqf = cq0 + ((crank - cq0 % csize) + csize) % csize;   // first such chunk owned by this CTA
        const int cqn = (N + 31) >> 5;
        const int nchunk = (cqf < cqn) ? (cqn - cqf + csize - 1) / csize : 0;
        const int ncol = 32 * nchunk;
        int nslab = (ncol > 0 && ncol < KA_NT) ? KA_NT / ncol : 1;
        if (nslab > 8) nslab = 8;
        const int nrow = N - c0;
        const int slab_rows = ((nrow + nslab - 1) / nslab + KA_RB - 1) / KA_RB * KA_RB;
        auto load_u = [&](int cc, double (&u)[KA_PW]) {                  // u = L11^-1 K[k0 .. k0+pw)[cc], K untouched
#pragma unroll
            for (int r = 0; r < KA_PW; ++r) u[r] = (r < pw) ? K[(size_t)(k0 + r) * ld + cc] : 0.0;
#pragma unroll
            for (int r = 1; r < KA_PW; ++r) {
                if (r < pw) {
                    double v = u[r];
#pragma unroll
                    for (int q = 0; q < r; ++q) v = fma(-Psm[r * KA_LDP + q], u[q], v);
                    u[r] = v;
                }
            }
        };
        auto store_u = [&](int cc, const double (&u)[KA_PW]) {
#pragma unroll
            for (int r = 0; r < KA_PW; ++r)
                if (r < pw) K[(size_t)(k0 + r) * ld + cc] = u[r];
        };
        auto update_rows = [&](int cc, const double (&u)[KA_PW], int i_begin, int i_end) {
            for (int i = i_begin; i < i_end; i += KA_RB) {
                double kv[KA_RB];
#pragma unroll
                for (int q = 0; q < KA_RB; ++q) kv[q] = (i + q < i_end) ? K[(size_t)(i + q) * ld + cc] : 0.0;
#pragma unroll
                for (int q = 0; q < KA_RB; ++q) {
                    const double* lr = Psm + (size_t)(i + q - k0) * KA_LDP;      // same address in every lane: broadcast
                    if (i + q < i_end) {
#pragma unroll
                        for (int r = 0; r < KA_PW; ++r) kv[q] = fma(-lr[r], u[r], kv[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < KA_RB; ++q)
                    if (i + q < i_end) K[(size_t)(i + q) * ld + cc] = kv[q];
            }
        };
        if (nslab == 1) {
            for (int w = tid; w < ncol; w += KA_NT) {
                const int cc = (cqf + (w >> 5) * csize) * 32 + (w & 31);
                if (cc < c0 || cc >= N) continue;
                double u[KA_PW];
                load_u(cc, u);
                store_u(cc, u);
                update_rows(cc, u, c0, N);
            }
        } else {
            // at most one work item per thread; U12 is written back only after every slab has read the raw column
            const int ci = (ncol > 0) ? tid % ncol : 0, slab = (ncol > 0) ? tid / ncol : nslab;
            const int cc = (cqf + (ci >> 5) * csize) * 32 + (ci & 31);
            const bool have = ncol > 0 && slab < nslab && cc >= c0 && cc < N;
            double u[KA_PW];
            if (have) load_u(cc, u);
            __syncthreads();
            if (have) {
                if (slab == 0) store_u(cc, u);
                const int ib = c0 + slab * slab_rows;
                const int ie = (ib + slab_rows < N) ? ib + slab_rows : N;
                update_rows(cc, u, ib, ie);
            }
        }
        // the next panel's columns were updated by whichever CTA owns their chunk
        if (csize > 1) ka_cluster_sync();
        else __syncthreads();
        KA_T(6);
    }
}

// Solve with the factors: vec (N, shared memory) <- K^-1 vec.  blk = 32 x 33 staging of one diagonal block.
__device__ void ka_lu_solve(const double* K, int N, int64_t ld, const int* piv, double* vec, double* blk) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int k = 0; k < N; ++k) {
            const int p = piv[k];
            if (p != k) { const double t0 = vec[k]; vec[k] = vec[p]; vec[p] = t0; }
        }
    }
    __syncthreads();
    const int nblk = (N + 31) >> 5;
    // forward, unit lower: the warps share the 32 rows of a block
    for (int I = 0; I < nblk; ++I) {
        const int i0 = I << 5;
        for (int w = warp; w < 32; w += KA_NW) {
            const int i = i0 + w;
            if (i < N) {
                const double* row = K + (size_t)i * ld;
                double acc = 0.0;
                for (int k = lane; k < i0; k += 32) acc = fma(row[k], vec[k], acc);
                acc = warp_sum(acc);
                blk[w * 33 + lane] = (i0 + lane < i) ? row[i0 + lane] : 0.0;      // strictly lower part of the block
                if (lane == 0) vec[i] -= acc;
            } else {
                blk[w * 33 + lane] = 0.0;
            }
        }
        __syncthreads();
        if (warp == 0) {
            double v = (i0 + lane < N) ? vec[i0 + lane] : 0.0;
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                const double vq = __shfl_sync(0xffffffffu, v, q);
                if (lane > q) v = fma(-blk[lane * 33 + q], vq, v);
            }
            if (i0 + lane < N) vec[i0 + lane] = v;
        }
        __syncthreads();
    }
    // backward, upper with diagonal
    for (int I = nblk - 1; I >= 0; --I) {
        const int i0 = I << 5;
        const int k1 = (i0 + 32 < N) ? i0 + 32 : N;
        for (int w = warp; w < 32; w += KA_NW) {
            const int i = i0 + w;
            if (i < N) {
                const double* row = K + (size_t)i * ld;
                double acc = 0.0;
                for (int k = k1 + lane; k < N; k += 32) acc = fma(row[k], vec[k], acc);
                acc = warp_sum(acc);
                blk[w * 33 + lane] = (i0 + lane >= i && i0 + lane < N) ? row[i0 + lane] : 0.0;   // upper part incl. diagonal
                if (lane == 0) vec[i] -= acc;
            } else {
                blk[w * 33 + lane] = (lane == w) ? 1.0 : 0.0;
            }
        }
        __syncthreads();
        if (warp == 0) {
            double v = (i0 + lane < N) ? vec[i0 + lane] : 0.0;
            const double dg = blk[lane * 33 + lane];
#pragma unroll 8
            for (int q = 31; q >= 0; --q) {
                const double vq = __shfl_sync(0xffffffffu, v / dg, q);     // lane q's final value
                if (lane == q) v = vq;
                else if (lane < q) v = fma(-blk[lane * 33 + q], vq, v);
            }
            if (i0 + lane < N) vec[i0 + lane] = v;
        }
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(KA_NT, 1) ka_solve(const KktArgs a) {
    extern __shared__ __align__(16) double smem_ka[];
    double* Psm = smem_ka;                                  // [N][KA_LDP] panel; the solves keep their vector here
    double* blk = smem_ka + (size_t)(a.m + a.n) * KA_LDP;   // [32][33]
    __shared__ double sh[32];
    __shared__ double shv[32];
    __shared__ int shi[32];
    __shared__ double bc;
    const int crank = (int)ka_cluster_rank(), csize = (int)ka_cluster_size();
    const int slot = blockIdx.x / csize;                    // one LP (and one K) per cluster
    const int lp = a.list ? a.list[slot] : slot;
    const int m = a.m, n = a.n, N = m + n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ld = ka_ldk(m, n);
    const double* A = a.A + (size_t)lp * m * n;
    const double* b = a.b + (size_t)lp * m;
    const double* c = a.c + (size_t)lp * n;
    double* x = a.x + (size_t)lp * n;
    double* s = a.s + (size_t)lp * n;
    double* y = a.y + (size_t)lp * m;
    double* scal = a.scal + (size_t)lp * S_COUNT;
    double* W = a.work + (size_t)slot * ka_work_doubles(m, n);
    double* K = W;                          W += (size_t)N * ld;
    W += (size_t)crank * (N + m + 6 * (size_t)n);            // this CTA's own copy of the small arrays
    int* piv = reinterpret_cast<int*>(W);   W += N;
    double* rb = W;                         W += m;
    double* rc = W;                         W += n;
    double* dxa = W;                        W += n;
    double* dsa = W;                        W += n;
    double* dx = W;                         W += n;
    double* ds = W;                         W += n;
    double* tt = W;                         // [n] rcomp / x of the current right-hand side
    double* vec = Psm;
    const double nb = scal[S_NB], nc = scal[S_NC];
    int it = a.iters[lp];
    long long* prof = (a.prof && blockIdx.x == 0) ? a.prof : nullptr;
    long long t_last = clock64();

    for (;;) {
        // ---- residuals of the current iterate, from scratch (main.py:67-70), check_optimality (main.py:169-173)
        double nrb2 = 0.0;
        for (int i = warp; i < m; i += KA_NW) {
            const double* row = A + (size_t)i * n;
            double acc = 0.0;
            for (int k = lane; k < n; k += 32) acc = fma(row[k], x[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const double r = acc - b[i];
                rb[i] = r;
                nrb2 += r * r;
            }
        }
        double nrc2 = 0.0, xs = 0.0, obj = 0.0;
        for (int j = tid; j < n; j += KA_NT) {
            double acc = 0.0;
            for (int i = 0; i < m; ++i) acc = fma(A[(size_t)i * n + j], y[i], acc);
            const double r = acc + s[j] - c[j];
            rc[j] = r;
            nrc2 += r * r;
            xs += x[j] * s[j];
            obj += x[j] * c[j];
        }
        nrb2 = ka_block_all<RED_SUM>(nrb2, sh, &bc);
        nrc2 = ka_block_all<RED_SUM>(nrc2, sh, &bc);
        xs = ka_block_all<RED_SUM>(xs, sh, &bc);
        obj = ka_block_all<RED_SUM>(obj, sh, &bc);
        const double nrb = sqrt(nrb2), nrc = sqrt(nrc2);
        const bool cont = (a.tol * (1.0 + nb) < nrb) || (a.tol * (1.0 + nc) < nrc) || (a.tol < xs);
        if (tid == 0 && crank == 0) {
            scal[S_NRB2] = nrb2; scal[S_NRB] = nrb; scal[S_NRC2] = nrc2; scal[S_NRC] = nrc;
            scal[S_XS] = xs; scal[S_OBJ] = obj; scal[S_CONT] = cont ? 1.0 : 0.0;
        }
        KA_T(0);
        if (!cont || it >= a.max_iter) break;

        // ---- K = [[-D^-1, A^T], [A, 0]]: the CTAs of the cluster share the work (interleaved slices)
        const int64_t gstride = (int64_t)KA_NT * csize, g0 = (int64_t)crank * KA_NT + tid;
        for (int64_t idx = g0; idx < (int64_t)N * ld; idx += gstride) K[idx] = 0.0;
        if (csize > 1) ka_cluster_sync();
        else __syncthreads();
        for (int64_t j = g0; j < n; j += gstride) K[(size_t)j * ld + j] = -(s[j] / x[j]);
        for (int64_t idx = g0; idx < (int64_t)m * n; idx += gstride) {
            const int i = (int)(idx / n), j = (int)(idx - (int64_t)i * n);
            const double v = A[idx];
            K[(size_t)(n + i) * ld + j] = v;
            K[(size_t)j * ld + n + i] = v;
        }
        if (csize > 1) ka_cluster_sync();
        else __syncthreads();
        KA_T(1);
        ka_lu_factor(K, N, ld, piv, Psm, shv, shi, crank, csize, prof);
        t_last = clock64();

        // ---- predictor (main.py:66-76, 101-109): rcomp = x s
        for (int j = tid; j < n; j += KA_NT) {
            const double q = (x[j] * s[j]) / x[j];
            tt[j] = q;
            vec[j] = -(rc[j] - q);
        }
        for (int i = tid; i < m; i += KA_NT) vec[n + i] = -rb[i];
        __syncthreads();
        KA_T(8);
        ka_lu_solve(K, N, ld, piv, vec, blk);
        KA_T(7);
        double minp = 1.0, mind = 1.0;
        for (int j = tid; j < n; j += KA_NT) {
            const double dxi = vec[j];
            const double dsi = (-s[j] * dxi / x[j]) - tt[j];
            dxa[j] = dxi; dsa[j] = dsi;
            if (dxi < 0.0) minp = fmin(minp, -x[j] / dxi);
            if (dsi < 0.0) mind = fmin(mind, -s[j] / dsi);
        }
        const double apa = ka_block_all<RED_MIN>(minp, sh, &bc);
        const double ada = ka_block_all<RED_MIN>(mind, sh, &bc);
        double part = 0.0;
        for (int j = tid; j < n; j += KA_NT) part += (x[j] + apa * dxa[j]) * (s[j] + ada * dsa[j]);
        part = ka_block_all<RED_SUM>(part, sh, &bc);
        const double mu_aff = part / (double)n, mu = xs / (double)n;
        const double ratio = mu_aff / mu, sigma = ratio * ratio * ratio;       // main.py:598-600, unclamped
        const double sigma_mu = sigma * mu;

        // ---- corrector (main.py:142-159): rcomp = x s + dxa dsa - sigma mu
        for (int j = tid; j < n; j += KA_NT) {
            const double rcomp = x[j] * s[j] + dxa[j] * dsa[j] - sigma_mu;
            const double q = rcomp / x[j];
            tt[j] = q;
            vec[j] = -(rc[j] - q);
        }
        for (int i = tid; i < m; i += KA_NT) vec[n + i] = -rb[i];
        __syncthreads();
        KA_T(8);
        ka_lu_solve(K, N, ld, piv, vec, blk);
        KA_T(7);
        minp = 1.0; mind = 1.0;
        for (int j = tid; j < n; j += KA_NT) {
            const double dxi = vec[j];
            const double dsi = (-s[j] * dxi / x[j]) - tt[j];
            dx[j] = dxi; ds[j] = dsi;
            if (dxi < 0.0) minp = fmin(minp, -x[j] / dxi);
            if (dsi < 0.0) mind = fmin(mind, -s[j] / dsi);
        }
        double ap = ka_block_all<RED_MIN>(minp, sh, &bc);
        double ad = ka_block_all<RED_MIN>(mind, sh, &bc);
        ap = fmin(1.0, a.eta * ap);                                            // main.py:616-623
        ad = fmin(1.0, a.eta * ad);
        // every CTA has read the old iterate for the last time (ratio tests above) before CTA 0 overwrites it
        if (csize > 1) ka_cluster_sync();
        if (crank == 0) {
            for (int j = tid; j < n; j += KA_NT) {
                x[j] = x[j] + ap * dx[j];
                s[j] = s[j] + ad * ds[j];
            }
            for (int i = tid; i < m; i += KA_NT) y[i] = y[i] + ad * vec[n + i];
            if (tid == 0) {
                scal[S_AP_AFF] = apa; scal[S_AD_AFF] = ada; scal[S_MU_AFF] = mu_aff; scal[S_MU] = mu;
                scal[S_SIGMA] = sigma; scal[S_SIGMA_MU] = sigma_mu; scal[S_AP] = ap; scal[S_AD] = ad;
            }
        }
        ++it;
        if (csize > 1) ka_cluster_sync();           // the new iterate is visible to the whole cluster
        else __syncthreads();
        KA_T(8);
        if (prof && tid == 0) prof[9] += 1;
    }
    if (tid == 0 && crank == 0) a.iters[lp] = it;
}

// count LPs (list[0..count), or LPs 0..count-1 when list is null), work_d: count * ka_work_doubles(m, n) doubles
inline std::atomic<int>& ka_cluster_ctas() {        // CTAs per LP: 1, 2, 4 (default) or 8 (ipm_set_kkt_cluster)
    static std::atomic<int> v{4};
    return v;
}
inline int ka_launch(const KktArgs& a, int count, cudaStream_t st) {
    if (a.m + a.n > KA_MAX_N) {
        g_last_error = "augmented-system kernel: n + m exceeds " + std::to_string(KA_MAX_N);
        return IPM_ERR_SHAPE;
    }
    IPM_TRY(ensure_dyn_smem(ka_solve, (size_t)(KA_MAX_N * KA_LDP + 32 * 33 + 64) * sizeof(double)));
    const int cl = ka_cluster_ctas().load();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(count * cl), 1, 1);
    cfg.blockDim = dim3(KA_NT, 1, 1);
    cfg.dynamicSmemBytes = ka_smem_bytes(a.m, a.n);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IPM_CUDA_OK(cudaLaunchKernelEx(&cfg, ka_solve, a));
    count_launch();
    return launch_check();
}
#endif

}  // namespace ipm
