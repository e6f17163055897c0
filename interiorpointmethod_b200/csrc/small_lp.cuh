// Whole solve of ONE small sparse LP in ONE launch of ONE CTA: the predictor-corrector loop of main.py:776-815 with every
// kernel of the single-LP path (ipm_single.cu: assemble, factor, two directions, sigma, update, residual check) called
// as a device function, a block barrier where the stream order used to be.  A small Netlib LP is launch-bound: AFIRO
// (27 x 51) ran 21 graph nodes and one host round trip per iteration, 87 us, for a few microseconds of arithmetic.
//
// The iteration's vectors and scalars live in shared memory for the whole solve (the phases are dependent chains of
// short loads); structure arrays, the scaled values and M stay in global memory.
//
// Eligible: sparse A, n <= 512 (every vector kernel is ONE block of the multi-kernel path too, so its reductions see
// the same partial sums in the same order), m <= 256 (fused Cholesky and the inverse-block triangular solves), no
// dependent-row mask, no refinement.  The device functions are the bodies of the stand-alone kernels, executed with
// the same block size: the iterates are bitwise what the multi-kernel path computes (tests/test_gpu_small_lp.py).
#pragma once
#include "chol.cuh"
#include "chol_batched.cuh"
#include "sparse.cuh"
#include "vec.cuh"

namespace ipm {

constexpr int SMALL_MAX_N = 2 * VEC_NT;      // vec_grid(n) == 1
constexpr int SMALL_MAX_M = KBC_MAX_M;       // 256
constexpr int64_t SMALL_MAX_TERMS = 10000;   // products of the SpGEMM pattern

struct SmallArgs {
    int m, n;
    int64_t nnz, nent, ldm;
    const int32_t *rowptr, *colind, *t_rowptr, *t_colind;
    const double *val, *t_val;
    double* ad;
    const int64_t *out_idx, *prod_ptr;
    const int32_t *pa, *pb;
    const double *b, *c;
    double *x, *y, *s, *rb, *rc, *d, *w, *rcx, *dxa, *dya, *dsa, *dx, *dy, *ds, *tm, *tn, *rhs;
    double* M;
    double *scal, *partials;
    unsigned* counter;
    double tol, eta, tau;
    int max_iter;
    int* k_out;
    long long* prof;      // nullable: 16 phase counters (clock64 cycles summed over the iterations; IPM_SMALL_PROF=1)
    int vec_off;          // doubles: where the vector slab starts in dynamic shared memory (after the factor / solve area)
};

inline std::atomic<int>& small_lp_fused() {          // ipm_set_small_lp_fused: 1 (default) = use the one-launch solve
    static std::atomic<int> on{1};
    return on;
}

// Dynamic shared memory: [work area of the fused Cholesky / the triangular solves][vector slab].  Every vector of the
// iteration lives in the slab for the whole solve (13 of length n, 8 of length m, the scalar block): the phases are
// dependent chains of short loads, 30 cycles each from shared memory instead of 600 from L2.
inline int small_slab_doubles(int m, int n) {
    const int np = (n + 1) & ~1, mp = (m + 1) & ~1;
    return 13 * np + 8 * mp + S_COUNT + 8;
}
inline size_t small_work_bytes(int m) {
    const size_t a = kbc_smem_bytes(m), b = trsv_batched_inv_smem(m);
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}
inline size_t small_smem_bytes(int m, int n) { return small_work_bytes(m) + (size_t)small_slab_doubles(m, n) * sizeof(double); }
inline size_t small_smem_bytes_max() { return small_smem_bytes(SMALL_MAX_M, SMALL_MAX_N); }

#ifdef __CUDACC__
__device__ __forceinline__ void small_direction(const SmallArgs& a, int kind, double* dxo, double* dyo, double* dso,
                                                long long* prof, long long& t_last) {
#define SMALL_TD(slot) do { if (prof && threadIdx.x == 0) { const long long _t = clock64(); prof[slot] += _t - t_last; t_last = _t; } } while (0)
    d_make_w(kind, a.x, a.s, a.rc, a.d, a.dxa, a.dsa, a.scal, a.rcx, a.w, a.n);
    __syncthreads();
    SMALL_TD(2);
    d_spmv_csr(a.m, a.rowptr, a.colind, a.val, a.w, a.tm);
    __syncthreads();
    d_make_rhs(a.rb, a.tm, a.rhs, a.m);
    __syncthreads();
    SMALL_TD(3);
    TrsvBatchedArgs t;
    t.L = a.M; t.ldm = a.ldm; t.strideM = 0; t.v = a.rhs; t.strideV = 0; t.m = a.m; t.active = nullptr; t.out = dyo;
    d_trsv_batched_inv(t);
    __syncthreads();
    SMALL_TD(4);
    d_spmv_csr(a.n, a.t_rowptr, a.t_colind, a.t_val, dyo, a.tn);
    __syncthreads();
    SMALL_TD(5);
    d_direction(kind, a.tn, a.d, a.w, a.rcx, a.x, a.s, dxo, dso, a.n, a.eta, a.scal, a.partials, a.counter);
    __syncthreads();
    SMALL_TD(6);
}

// Precondition: the residual check of the starting point has run (scal[S_CONT] is valid), like before the host loop of
// ipm_solve.  On exit *k_out = iterations taken, scal holds the residual check of the last iterate.
static __global__ void __launch_bounds__(VEC_NT, 1) k_small_solve(const SmallArgs g) {
    extern __shared__ __align__(16) double small_smem[];
    // ---- the iteration's vectors move to shared memory; `a` is the argument block with the pointers redirected
    SmallArgs a = g;
    {
        const int np = (g.n + 1) & ~1, mp = (g.m + 1) & ~1;
        double* p = small_smem + g.vec_off;
        double* nb = p;             p += 13 * np;
        double* mb = p;             p += 8 * mp;
        a.scal = p;                 p += S_COUNT;
        a.partials = p;
        double* bs = mb + 6 * mp; double* cs = nb + 12 * np;
        a.x = nb; a.s = nb + np; a.rc = nb + 2 * np; a.d = nb + 3 * np; a.w = nb + 4 * np; a.rcx = nb + 5 * np;
        a.dxa = nb + 6 * np; a.dsa = nb + 7 * np; a.dx = nb + 8 * np; a.ds = nb + 9 * np; a.tn = nb + 10 * np;
        a.y = mb; a.rb = mb + mp; a.dya = mb + 2 * mp; a.dy = mb + 3 * mp; a.tm = mb + 4 * mp; a.rhs = mb + 5 * mp;
        for (int i = threadIdx.x; i < g.n; i += VEC_NT) {
            a.x[i] = g.x[i]; a.s[i] = g.s[i]; a.rc[i] = g.rc[i]; a.d[i] = g.d[i]; cs[i] = g.c[i];
        }
        for (int i = threadIdx.x; i < g.m; i += VEC_NT) { a.y[i] = g.y[i]; a.rb[i] = g.rb[i]; bs[i] = g.b[i]; }
        for (int i = threadIdx.x; i < S_COUNT; i += VEC_NT) a.scal[i] = g.scal[i];
        a.b = bs; a.c = cs;
    }
    __syncthreads();
    long long t_last = clock64();
#define SMALL_T(slot) do { if (g.prof && threadIdx.x == 0) { const long long _t = clock64(); g.prof[slot] += _t - t_last; t_last = _t; } } while (0)
    int k = 0;
    while (k < a.max_iter) {
        if (!(a.scal[S_CONT] > 0.5)) break;                   // uniform: written before the last barrier
        // ---- assemble M = A diag(x/s) A^T on the fixed pattern (main.py:223-224)
        for (int64_t i = threadIdx.x; i < (int64_t)a.m * a.ldm; i += VEC_NT) a.M[i] = 0.0;
        d_scale_vals(a.nnz, a.colind, a.val, a.d, a.ad);
        __syncthreads();
        d_spgemm_numeric(a.nent, a.out_idx, a.prod_ptr, a.pa, a.pb, a.ad, a.val, a.M);
        __syncthreads();
        SMALL_T(0);
        // ---- factor (main.py:176-182)
        {
            CholBatchedArgs ca;
            ca.M = a.M; ca.ldm = a.ldm; ca.strideM = 0; ca.scal = a.scal; ca.strideScal = 0; ca.tau = a.tau; ca.m = a.m;
            ca.active = nullptr;
            d_kb_chol<KBC_NT>(ca);
        }
        __syncthreads();
        SMALL_T(1);
        small_direction(a, 0, a.dxa, a.dya, a.dsa, g.prof, t_last);           // main.py:783
        d_sigma(a.x, a.s, a.dxa, a.dsa, a.n, a.scal, a.partials, a.counter);      // main.py:795
        __syncthreads();
        SMALL_T(7);
        small_direction(a, 1, a.dx, a.dy, a.ds, g.prof, t_last);              // main.py:799
        d_update(a.x, a.y, a.s, a.dx, a.dy, a.ds, a.m, a.n, a.scal, -1.0, -1.0);  // main.py:803
        __syncthreads();
        SMALL_T(8);
        // ---- check_optimality of the new iterate (main.py:780)
        d_spmv_csr(a.m, a.rowptr, a.colind, a.val, a.x, a.tm);
        __syncthreads();
        d_resid_primal(a.tm, a.b, a.rb, a.m, a.scal, a.partials, a.counter);
        __syncthreads();
        d_spmv_csr(a.n, a.t_rowptr, a.t_colind, a.t_val, a.y, a.tn);
        __syncthreads();
        d_resid_dual(a.tn, a.s, a.c, a.x, a.rc, a.d, a.n, a.tol, a.scal, a.partials, a.counter);
        __syncthreads();
        SMALL_T(9);
        ++k;
    }
    // ---- the state the handle's other entry points expect in global memory: iterate, residuals, d, the scalar block
    // and the last directions (ipm_get_direction after a solve)
    for (int i = threadIdx.x; i < g.n; i += VEC_NT) {
        g.x[i] = a.x[i]; g.s[i] = a.s[i]; g.rc[i] = a.rc[i]; g.d[i] = a.d[i];
        g.w[i] = a.w[i]; g.rcx[i] = a.rcx[i]; g.dxa[i] = a.dxa[i]; g.dsa[i] = a.dsa[i]; g.dx[i] = a.dx[i]; g.ds[i] = a.ds[i];
    }
    for (int i = threadIdx.x; i < g.m; i += VEC_NT) {
        g.y[i] = a.y[i]; g.rb[i] = a.rb[i]; g.dya[i] = a.dya[i]; g.dy[i] = a.dy[i];
    }
    for (int i = threadIdx.x; i < S_COUNT; i += VEC_NT) g.scal[i] = a.scal[i];
    if (threadIdx.x == 0) *g.k_out = k;
}
#endif

}  // namespace ipm
