// C ABI (include/ipm_b200.h), single-LP part: handle, problem upload, op-level entry points that mirror the
// reference's Python seams (main.py:162-322, 562-697) and the device-resident predictor-corrector loop.
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "chol.cuh"
#include <cstdio>
#include "potrf_auto.cuh"
#include "small_lp.cuh"
#include "common.cuh"
#include "dense.cuh"
#include "dmma_gemm.cuh"
#include "dmma_ws.cuh"
#include "ingest.cuh"
#include "sparse.cuh"
#include "trsv_pipe.cuh"
#include "vec.cuh"

namespace ipm {
std::atomic<int64_t> g_launches{0};
thread_local std::string g_last_error;
}  // namespace ipm

using namespace ipm;

// above this order the triangular solves run as pipelined persistent kernels (one CTA per 128-block, k_trinv128 once
// per factorisation); at or below it one CTA does both sweeps (k_trsv_batched).  Measured crossover between m = 444
// and m = 821 (round 2: SCSD8 m = 397 +7 %, BANDM m = 305 +17 % with one CTA; 25FV47 m = 821 -12 %, TRUSS m = 1000
// -18 %).  IPM_PIPE_MIN_M overrides it for A/B runs.
static int pipe_min_m() {
    static const int v = [] { const char* e = getenv("IPM_PIPE_MIN_M"); return e ? atoi(e) : 512; }();
    return v;
}

struct ipm_handle {
    int dev = 0;
    cudaStream_t st = nullptr;
    int m = 0, n = 0;
    bool loaded = false, dense = false;
    // sparse A (CSR) and A^T (CSR)
    int64_t nnz = 0;
    // structure (both orientations + SpGEMM pattern): owned by `pat`, shared through the pattern cache (ingest.cuh)
    std::shared_ptr<DevPattern> pat;
    bool pat_hit = false;
    double load_ms = 0.0;
    int32_t *rowptr = nullptr, *colind = nullptr, *t_rowptr = nullptr, *t_colind = nullptr;
    double* vslab = nullptr;      // val | t_val | ad
    double *val = nullptr, *t_val = nullptr, *ad = nullptr;
    // SpGEMM pattern
    int64_t nent = 0;
    int64_t *out_idx = nullptr, *prod_ptr = nullptr;
    int32_t *pa = nullptr, *pb = nullptr;
    // dense A
    const double* A = nullptr;
    int64_t lda = 0;
    double* A_own = nullptr;
    double* gemv_partial = nullptr;
    int nchunks = 0;
    // vectors (one slab)
    double* slab = nullptr;
    double *b, *c, *x, *y, *s, *rb, *rc, *d, *w, *rcx, *dxa, *dya, *dsa, *dx, *dy, *ds, *tm, *tn, *rhs, *tmp_m;
    double* M = nullptr;
    int64_t ldm = 0;
    TrsvPipeWs pipe;              // m > pipe_min_m(): inverses of the diagonal blocks + flags of the pipelined solves
    unsigned char* dep = nullptr; // m bytes: dependent rows of A (ipm_detect_dependent_rows), nullptr = none
    int n_dep = 0;
    double* scal = nullptr;
    double* partials = nullptr;
    unsigned* counter = nullptr;
    double* h_scal = nullptr;     // pinned mirror of scal
    double tol = 1e-8;
    double eta = 0.91;            // main.py:607
    double tau = 1e-30;
    double refine_thresh = -1.0;  // >= 0: conditional refinement of the corrector (ipm_set_refinement); off by default
    // one predictor-corrector iteration captured as a CUDA graph (the small Netlib LPs are launch-bound)
    cudaGraphExec_t gexec = nullptr;
    double g_tol = 0.0, g_tau = 0.0;
    bool g_dep = false;
    double g_refine = -1.0;
    bool use_graph = true;
    int64_t launches_per_graph = 0;
    bool have_resid = false, have_M = false, have_factor = false, have_pred = false, have_sigma = false,
         have_corr = false;
    std::string err;
};

namespace {

int fail(ipm_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    g_last_error = msg;
    return code;
}
int cuda_fail(ipm_handle* h) {
    if (h) h->err = g_last_error;
    return IPM_ERR_CUDA;
}
#define H_CUDA(expr)                                   \
    do {                                               \
        int _r = [&]() -> int { IPM_CUDA_OK(expr); return IPM_OK; }(); \
        if (_r != IPM_OK) return cuda_fail(h);         \
    } while (0)
#define H_TRY(expr)                                    \
    do {                                               \
        int _r = (expr);                               \
        if (_r == IPM_ERR_CUDA) return cuda_fail(h);   \
        if (_r != IPM_OK) return _r;                   \
    } while (0)

void free_problem(ipm_handle* h) {
    cudaSetDevice(h->dev);
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    void* ptrs[] = {h->vslab, h->A_own, h->gemv_partial, h->slab, h->M, h->pipe.Linv, h->pipe.flags, h->dep};
    h->dep = nullptr; h->n_dep = 0;
    h->pipe = TrsvPipeWs();
    for (void* p : ptrs)
        if (p) cudaFree(p);
    h->pat.reset();               // the structure arrays live on in the cache (or die with the last user)
    h->vslab = nullptr;
    h->rowptr = h->colind = h->t_rowptr = h->t_colind = nullptr;
    h->val = h->t_val = h->ad = nullptr;
    h->out_idx = h->prod_ptr = nullptr;
    h->pa = h->pb = nullptr;
    h->A_own = nullptr; h->A = nullptr; h->gemv_partial = nullptr; h->slab = nullptr; h->M = nullptr;
    h->loaded = false;
    h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
}

int alloc_common(ipm_handle* h, int m, int n) {
    h->m = m; h->n = n;
    const int64_t pm = round_up(m, 16), pn = round_up(n, 16);
    const int64_t total = 6 * pm /*b y rb dya dy tm*/ + 2 * pm /*rhs tmp_m*/ + 12 * pn;
    H_CUDA(cudaMalloc(&h->slab, (size_t)total * sizeof(double)));
    H_CUDA(cudaMemsetAsync(h->slab, 0, (size_t)total * sizeof(double), h->st));
    double* p = h->slab;
    auto take = [&](int64_t len) { double* q = p; p += len; return q; };
    h->b = take(pm); h->y = take(pm); h->rb = take(pm); h->dya = take(pm); h->dy = take(pm); h->tm = take(pm);
    h->rhs = take(pm); h->tmp_m = take(pm);
    h->c = take(pn); h->x = take(pn); h->s = take(pn); h->rc = take(pn); h->d = take(pn); h->w = take(pn);
    h->rcx = take(pn); h->dxa = take(pn); h->dsa = take(pn); h->dx = take(pn); h->ds = take(pn); h->tn = take(pn);
    h->ldm = pm;
    H_CUDA(cudaMalloc(&h->M, (size_t)m * h->ldm * sizeof(double)));
    if (m > pipe_min_m() && m > 256) {
        h->pipe.nblk = ceil_div(m, TP_NB);
        H_CUDA(cudaMalloc(&h->pipe.Linv, (size_t)h->pipe.nblk * TP_NB * TP_NB * sizeof(double)));
        H_CUDA(cudaMalloc(&h->pipe.flags, (size_t)2 * h->pipe.nblk * sizeof(int)));
    }
    return IPM_OK;
}

// triangular solves with the current factor: pipelined persistent kernels for large m, one CTA otherwise
int solve_factored(ipm_handle* h, double* rhs, double* sol) {
    if (h->pipe.Linv) return potrs_pipe(h->M, h->ldm, h->m, h->pipe, rhs, h->tmp_m, sol, h->st);
    return potrs_single(h->M, h->ldm, h->m, rhs, h->tmp_m, sol, h->st);
}

int finish_load(ipm_handle* h) {
    // |b|, |c| once per problem (main.py:169-170 recomputes them every iteration)
    k_norm2<<<vec_grid(h->m), VEC_NT, 0, h->st>>>(h->b, h->m, h->scal + S_NB, h->partials, h->counter);
    k_norm2<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(h->c, h->n, h->scal + S_NC, h->partials, h->counter);
    count_launch(2);
    H_TRY(launch_check());
    H_CUDA(cudaStreamSynchronize(h->st));
    h->loaded = true;
    return IPM_OK;
}

int matvec_A(ipm_handle* h, const double* v, double* out) {
    if (h->dense) {
        const int blocks = std::min<int64_t>(ceil_div((int64_t)h->m * 32, GEMV_NT), 8 * kNumSMs * 4);
        k_gemv_rows<<<blocks, GEMV_NT, 0, h->st>>>(h->m, h->n, h->A, h->lda, v, out);
    } else {
        const int blocks = std::min<int64_t>(ceil_div((int64_t)h->m * 8, 256), 8 * kNumSMs);
        k_spmv_csr<<<blocks, 256, 0, h->st>>>(h->m, h->rowptr, h->colind, h->val, v, out);
    }
    count_launch();
    return launch_check();
}

int matvec_AT(ipm_handle* h, const double* u, double* out) {
    if (h->dense) {
        dim3 grid(ceil_div(h->n, GEMV_NT), h->nchunks);
        double* partial = (h->nchunks == 1) ? out : h->gemv_partial;
        k_gemv_cols_partial<<<grid, GEMV_NT, 0, h->st>>>(h->m, h->n, h->A, h->lda, u, partial);
        count_launch();
        if (h->nchunks > 1) {
            k_gemv_cols_combine<<<ceil_div(h->n, 256), 256, 0, h->st>>>(h->n, h->nchunks, partial, out);
            count_launch();
        }
    } else {
        const int blocks = std::min<int64_t>(ceil_div((int64_t)h->n * 8, 256), 8 * kNumSMs);
        k_spmv_csr<<<blocks, 256, 0, h->st>>>(h->n, h->t_rowptr, h->t_colind, h->t_val, u, out);
        count_launch();
    }
    return launch_check();
}

int residual_step(ipm_handle* h) {
    H_TRY(matvec_A(h, h->x, h->tm));
    k_resid_primal<<<vec_grid(h->m), VEC_NT, 0, h->st>>>(h->tm, h->b, h->rb, h->m, h->scal, h->partials, h->counter);
    count_launch();
    H_TRY(matvec_AT(h, h->y, h->tn));
    k_resid_dual<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(h->tn, h->s, h->c, h->x, h->rc, h->d, h->n, h->tol, h->scal,
                                                       h->partials, h->counter);
    count_launch();
    H_TRY(launch_check());
    h->have_resid = true;
    h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    return IPM_OK;
}

int assemble_step(ipm_handle* h) {
    if (h->dense) {
        DmmaArgs g;
        g.P = h->A; g.ldp = h->lda; g.strideP = 0;
        g.Q = h->A; g.ldq = h->lda; g.strideQ = 0;
        g.dvec = h->d; g.strideD = 0;
        g.C = h->M; g.ldc = h->ldm; g.strideC = 0;
        g.rowsP = h->m; g.rowsQ = h->m; g.K = h->n; g.lower_only = 1; g.active = nullptr;
        H_TRY((dmma_syrk_auto<0>(g, 1, h->st)));
    } else {
        H_CUDA(cudaMemsetAsync(h->M, 0, (size_t)h->m * h->ldm * sizeof(double), h->st));
        k_scale_vals<<<std::max(1, std::min<int>(ceil_div(h->nnz, 256), 8 * kNumSMs)), 256, 0, h->st>>>(
            h->nnz, h->colind, h->val, h->d, h->ad);
        k_spgemm_numeric<<<std::max(1, std::min<int>(ceil_div(h->nent, 256), 8 * kNumSMs)), 256, 0, h->st>>>(
            h->nent, h->out_idx, h->prod_ptr, h->pa, h->pb, h->ad, h->val, h->M);
        count_launch(2);
        H_TRY(launch_check());
    }
    h->have_M = true;
    h->have_factor = false;
    return IPM_OK;
}

int factor_step(ipm_handle* h, double tau, int dep_mode = 1) {
    DepMask dm;
    dm.mask = h->dep; dm.mode = h->dep ? dep_mode : 0;
    H_TRY((potrf_single_auto(h->M, h->ldm, h->m, h->scal, tau, h->st, dm)));
    if (h->pipe.Linv) H_TRY(trinv_blocks(h->M, h->ldm, h->m, h->pipe, h->st));
    h->have_factor = true;
    h->have_M = false;
    return IPM_OK;
}

int direction_step(ipm_handle* h, int kind) {
    double* dxo = kind ? h->dx : h->dxa;
    double* dyo = kind ? h->dy : h->dya;
    double* dso = kind ? h->ds : h->dsa;
    k_make_w<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(kind, h->x, h->s, h->rc, h->d, h->dxa, h->dsa, h->scal, h->rcx,
                                                   h->w, h->n);
    count_launch();
    H_TRY(matvec_A(h, h->w, h->tm));
    k_make_rhs<<<vec_grid(h->m), VEC_NT, 0, h->st>>>(h->rb, h->tm, h->rhs, h->m);
    count_launch();
    H_TRY(solve_factored(h, h->rhs, dyo));
    H_TRY(matvec_AT(h, dyo, h->tn));
    k_direction<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(kind, h->tn, h->d, h->w, h->rcx, h->x, h->s, dxo, dso, h->n,
                                                      h->eta, h->scal, h->partials, h->counter);
    count_launch();
    H_TRY(launch_check());
    if (kind == 1 && h->refine_thresh >= 0.0) {
        // delta = -rb - A dx -> rhs; M ddy = delta -> tm; dy += ddy when |delta| > thresh |rb|; dx, ds again
        H_TRY(matvec_A(h, dxo, h->tm));
        k_refine_delta<<<vec_grid(h->m), VEC_NT, 0, h->st>>>(h->rb, h->tm, h->rhs, h->m, h->refine_thresh, h->scal,
                                                             h->partials, h->counter);
        count_launch();
        H_TRY(solve_factored(h, h->rhs, h->tm));
        k_add_if<<<vec_grid(h->m), VEC_NT, 0, h->st>>>(dyo, h->tm, h->m, h->scal + S_REFINE_FLAG);
        count_launch();
        H_TRY(matvec_AT(h, dyo, h->tn));
        k_direction<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(kind, h->tn, h->d, h->w, h->rcx, h->x, h->s, dxo, dso, h->n,
                                                          h->eta, h->scal, h->partials, h->counter);
        count_launch();
        H_TRY(launch_check());
    }
    if (kind == 0) { h->have_pred = true; h->have_sigma = false; h->have_corr = false; }
    else h->have_corr = true;
    return IPM_OK;
}

int sigma_step(ipm_handle* h) {
    k_sigma<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(h->x, h->s, h->dxa, h->dsa, h->n, h->scal, h->partials, h->counter);
    count_launch();
    H_TRY(launch_check());
    h->have_sigma = true;
    return IPM_OK;
}

int update_step(ipm_handle* h, double ap, double ad) {
    const int len = std::max(h->m, h->n);
    k_update<<<vec_grid(len), VEC_NT, 0, h->st>>>(h->x, h->y, h->s, h->dx, h->dy, h->ds, h->m, h->n, h->scal, ap, ad);
    count_launch();
    H_TRY(launch_check());
    h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    return IPM_OK;
}

int fetch_scal(ipm_handle* h) {
    H_CUDA(cudaMemcpyAsync(h->h_scal, h->scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int check_handle(ipm_handle* h, bool need_loaded = true) {
    if (!h) return IPM_ERR_ARG;
    if (cudaSetDevice(h->dev) != cudaSuccess) return fail(h, IPM_ERR_CUDA, "cudaSetDevice failed");
    if (need_loaded && !h->loaded) return fail(h, IPM_ERR_STATE, "no problem loaded");
    return IPM_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* ipm_version(void) { return "interiorpointmethod_b200 0.1 (sm_100a)"; }
int64_t ipm_launch_count(void) { return g_launches.load(); }

const char* ipm_last_error(const ipm_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int ipm_create(ipm_handle** out, int device_ordinal) {
    if (!out) return IPM_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_last_error = "no CUDA device available (this library has no CPU fallback)";
        return IPM_ERR_CUDA;
    }
    if (device_ordinal < 0 || device_ordinal >= count) {
        g_last_error = "device ordinal out of range";
        return IPM_ERR_ARG;
    }
    ipm_handle* h = new ipm_handle();
    h->dev = device_ordinal;
    h->use_graph = (getenv("IPM_NO_GRAPH") == nullptr);
    auto init = [&]() -> int {
        IPM_CUDA_OK(cudaSetDevice(h->dev));
        IPM_CUDA_OK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
        IPM_CUDA_OK(cudaMalloc(&h->scal, S_COUNT * sizeof(double)));
        IPM_CUDA_OK(cudaMemset(h->scal, 0, S_COUNT * sizeof(double)));
        IPM_CUDA_OK(cudaMalloc(&h->partials, (size_t)VEC_MAX_BLOCKS * 4 * sizeof(double)));
        IPM_CUDA_OK(cudaMalloc(&h->counter, 4 * sizeof(unsigned)));      // [1]: iterations of k_small_solve
        IPM_CUDA_OK(cudaMemset(h->counter, 0, 4 * sizeof(unsigned)));
        IPM_CUDA_OK(cudaMallocHost(&h->h_scal, S_COUNT * sizeof(double)));
        return IPM_OK;
    };
    if (init() != IPM_OK) {
        ipm_destroy(h);
        return IPM_ERR_CUDA;
    }
    *out = h;
    return IPM_OK;
}

void ipm_destroy(ipm_handle* h) {
    if (!h) return;
    cudaSetDevice(h->dev);
    if (h->st) cudaStreamSynchronize(h->st);
    free_problem(h);
    if (h->scal) cudaFree(h->scal);
    if (h->partials) cudaFree(h->partials);
    if (h->counter) cudaFree(h->counter);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

int ipm_set_pivot_threshold(ipm_handle* h, double pivot_rel_thresh) {
    if (!h) return IPM_ERR_ARG;
    if (!(pivot_rel_thresh >= 0.0)) return fail(h, IPM_ERR_ARG, "threshold must be >= 0");
    h->tau = pivot_rel_thresh;
    return IPM_OK;
}

// Sparse A in either compressed orientation (from_csc: ptr = column pointers, idx = row indices).  Structure work
// (other orientation, symbolic SpGEMM) happens on the device and is cached per structure (ingest.cuh).
static int load_sparse(ipm_handle* h, int m, int n, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                       const double* val, const double* b, const double* c, bool from_csc) {
    const auto t_begin = std::chrono::steady_clock::now();
    H_TRY(check_handle(h, false));
    if (!ptr || !idx || !val || !b || !c) return fail(h, IPM_ERR_ARG, "null pointer");
    const int nseg = from_csc ? n : m, nother = from_csc ? m : n;
    if (m <= 0 || n <= 0 || nnz < 0 || nnz > INT32_MAX || ptr[0] != 0 || ptr[nseg] != nnz)
        return fail(h, IPM_ERR_SHAPE, "bad compressed-sparse header");
    for (int i = 0; i < nseg; ++i)         // the whole pointer array first: idx is only read inside [0, nnz)
        if (ptr[i + 1] < ptr[i] || ptr[i + 1] > nnz) return fail(h, IPM_ERR_SHAPE, "pointer array not monotone within [0, nnz]");
    for (int i = 0; i < nseg; ++i) {
        for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) {
            if (idx[p] < 0 || idx[p] >= nother) return fail(h, IPM_ERR_SHAPE, "index out of range");
            if (p > ptr[i] && idx[p] <= idx[p - 1])
                return fail(h, IPM_ERR_SHAPE, "indices must be strictly ascending inside a row/column");
        }
    }
    free_problem(h);
    h->dense = false;
    h->nnz = nnz;
    H_TRY(alloc_common(h, m, n));
    {
        int r = acquire_pattern(h->dev, m, n, nnz, h->ldm, from_csc, ptr, idx, h->st, h->pat, h->pat_hit);
        if (r == IPM_ERR_CUDA) return cuda_fail(h);
        if (r != IPM_OK) return r;
    }
    const DevPattern& P = *h->pat;
    h->rowptr = P.rowptr; h->colind = P.colind; h->t_rowptr = P.t_rowptr; h->t_colind = P.t_colind;
    h->out_idx = P.out_idx; h->prod_ptr = P.prod_ptr; h->pa = P.pa; h->pb = P.pb;
    h->nent = P.nent;
    const size_t nz = (size_t)round_up(std::max<int64_t>(nnz, 1), 2);
    H_CUDA(cudaMalloc(&h->vslab, 3 * nz * sizeof(double)));
    h->val = h->vslab; h->t_val = h->val + nz; h->ad = h->t_val + nz;
    double* given = from_csc ? h->t_val : h->val;       // A in CSC order = A^T in CSR order
    double* derived = from_csc ? h->val : h->t_val;
    H_CUDA(cudaMemcpyAsync(given, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, h->st));
    k_ing_gather<<<std::max(1, std::min<int>(ceil_div(nnz, 256), 8 * kNumSMs)), 256, 0, h->st>>>(nnz, given, P.perm,
                                                                                                 derived);
    count_launch();
    H_CUDA(cudaMemcpyAsync(h->b, b, m * sizeof(double), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->c, c, n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    H_TRY(finish_load(h));
    h->load_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return IPM_OK;
}

int ipm_load_csr(ipm_handle* h, int m, int n, int64_t nnz, const int32_t* rowptr, const int32_t* colind,
                 const double* val, const double* b, const double* c) {
    if (!h) return IPM_ERR_ARG;
    return load_sparse(h, m, n, nnz, rowptr, colind, val, b, c, false);
}

int ipm_load_csc(ipm_handle* h, int m, int n, int64_t nnz, const int32_t* colptr, const int32_t* rowind,
                 const double* val, const double* b, const double* c) {
    if (!h) return IPM_ERR_ARG;
    return load_sparse(h, m, n, nnz, colptr, rowind, val, b, c, true);
}

int ipm_set_ingest_mode(int host_symbolic, int use_cache) {
    ingest_mode().store((host_symbolic ? 1 : 0) | (use_cache ? 0 : 2));
    return IPM_OK;
}

int ipm_pattern_info(ipm_handle* h, int64_t out[6]) {
    H_TRY(check_handle(h));
    if (!out) return fail(h, IPM_ERR_ARG, "null pointer");
    if (h->dense || !h->pat) return fail(h, IPM_ERR_STATE, "no sparse problem loaded");
    out[0] = h->pat->nent; out[1] = h->pat->nterms; out[2] = h->pat_hit ? 1 : 0;
    out[3] = h->pat->device_built ? 1 : 0;
    out[4] = (int64_t)(h->pat->build_ms * 1e3); out[5] = (int64_t)(h->load_ms * 1e3);
    return IPM_OK;
}

int ipm_get_pattern(ipm_handle* h, int32_t* rowptr, int32_t* colind, int32_t* t_rowptr, int32_t* t_colind,
                    int64_t* out_idx, int64_t* prod_ptr, int32_t* pa, int32_t* pb) {
    H_TRY(check_handle(h));
    if (h->dense || !h->pat) return fail(h, IPM_ERR_STATE, "no sparse problem loaded");
    const DevPattern& P = *h->pat;
    auto get = [&](void* dst, const void* src, size_t bytes) -> int {
        if (dst && bytes) IPM_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st));
        return IPM_OK;
    };
    H_TRY(get(rowptr, P.rowptr, (size_t)(P.m + 1) * sizeof(int32_t)));
    H_TRY(get(colind, P.colind, (size_t)P.nnz * sizeof(int32_t)));
    H_TRY(get(t_rowptr, P.t_rowptr, (size_t)(P.n + 1) * sizeof(int32_t)));
    H_TRY(get(t_colind, P.t_colind, (size_t)P.nnz * sizeof(int32_t)));
    H_TRY(get(out_idx, P.out_idx, (size_t)P.nent * sizeof(int64_t)));
    H_TRY(get(prod_ptr, P.prod_ptr, (size_t)(P.nent + 1) * sizeof(int64_t)));
    H_TRY(get(pa, P.pa, (size_t)P.nterms * sizeof(int32_t)));
    H_TRY(get(pb, P.pb, (size_t)P.nterms * sizeof(int32_t)));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_get_values(ipm_handle* h, double* val_csr, double* val_csc) {
    H_TRY(check_handle(h));
    if (h->dense || !h->pat) return fail(h, IPM_ERR_STATE, "no sparse problem loaded");
    if (val_csr) H_CUDA(cudaMemcpyAsync(val_csr, h->val, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (val_csc) H_CUDA(cudaMemcpyAsync(val_csc, h->t_val, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_pattern_cache_stats(int64_t out[4]) {
    if (!out) return IPM_ERR_ARG;
    PatternCache& C = pattern_cache();
    std::lock_guard<std::mutex> g(C.mu);
    size_t bytes = 0;
    for (auto& p : C.items) bytes += p->bytes;
    out[0] = (int64_t)C.items.size(); out[1] = C.hits; out[2] = C.misses; out[3] = (int64_t)bytes;
    return IPM_OK;
}

int ipm_pattern_cache_clear(void) {
    pattern_cache_clear();
    return IPM_OK;
}

static int load_dense_common(ipm_handle* h, int m, int n) {
    h->dense = true;
    H_TRY(alloc_common(h, m, n));
    h->nchunks = ceil_div(m, GEMVC_ROWS);
    if (h->nchunks > 1) H_CUDA(cudaMalloc(&h->gemv_partial, (size_t)h->nchunks * n * sizeof(double)));
    return IPM_OK;
}

int ipm_load_dense(ipm_handle* h, int m, int n, const double* A, int64_t lda, const double* b, const double* c) {
    H_TRY(check_handle(h, false));
    if (!A || !b || !c) return fail(h, IPM_ERR_ARG, "null pointer");
    if (m <= 0 || n <= 0 || lda < n) return fail(h, IPM_ERR_SHAPE, "bad dense shape");
    free_problem(h);
    H_TRY(load_dense_common(h, m, n));
    h->lda = round_up(n, 16);
    H_CUDA(cudaMalloc(&h->A_own, (size_t)m * h->lda * sizeof(double)));
    H_CUDA(cudaMemsetAsync(h->A_own, 0, (size_t)m * h->lda * sizeof(double), h->st));
    H_CUDA(cudaMemcpy2DAsync(h->A_own, h->lda * sizeof(double), A, lda * sizeof(double), n * sizeof(double), m,
                             cudaMemcpyHostToDevice, h->st));
    h->A = h->A_own;
    H_CUDA(cudaMemcpyAsync(h->b, b, m * sizeof(double), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->c, c, n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    return finish_load(h);
}

int ipm_load_dense_d(ipm_handle* h, int m, int n, const double* A_d, int64_t lda, const double* b_d,
                     const double* c_d) {
    H_TRY(check_handle(h, false));
    if (!A_d || !b_d || !c_d) return fail(h, IPM_ERR_ARG, "null pointer");
    if (m <= 0 || n <= 0 || lda < n) return fail(h, IPM_ERR_SHAPE, "bad dense shape");
    free_problem(h);
    H_TRY(load_dense_common(h, m, n));
    h->A = A_d;
    h->lda = lda;
    H_CUDA(cudaMemcpyAsync(h->b, b_d, m * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->c, c_d, n * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    return finish_load(h);
}

int ipm_init_state(ipm_handle* h, int y0_is_one) {
    H_TRY(check_handle(h));
    k_fill<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(h->x, h->n, 1.0);
    k_fill<<<vec_grid(h->n), VEC_NT, 0, h->st>>>(h->s, h->n, 1.0);
    k_fill<<<vec_grid(h->m), VEC_NT, 0, h->st>>>(h->y, h->m, y0_is_one ? 1.0 : 0.0);
    count_launch(3);
    H_TRY(launch_check());
    h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    return IPM_OK;
}

int ipm_start_mehrotra(ipm_handle* h) {
    H_TRY(check_handle(h));
    const int m = h->m, n = h->n;
    // d = 1: M = A A^T, factored once with the same safeguarded Cholesky the iterations use
    k_fill<<<vec_grid(n), VEC_NT, 0, h->st>>>(h->d, n, 1.0);
    count_launch();
    H_TRY(assemble_step(h));
    H_TRY(factor_step(h, h->tau));
    // x = A^T (A A^T)^-1 b
    H_CUDA(cudaMemcpyAsync(h->rhs, h->b, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    H_TRY(solve_factored(h, h->rhs, h->dy));
    H_TRY(matvec_AT(h, h->dy, h->x));
    // y = (A A^T)^-1 A c,  s = c - A^T y
    H_TRY(matvec_A(h, h->c, h->rhs));
    H_TRY(solve_factored(h, h->rhs, h->y));
    H_TRY(matvec_AT(h, h->y, h->tn));
    k_sub<<<vec_grid(n), VEC_NT, 0, h->st>>>(h->c, h->tn, h->s, n);
    k_mehrotra_shift<<<1, 1024, 0, h->st>>>(h->x, h->s, n);
    count_launch(2);
    H_TRY(launch_check());
    h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    return IPM_OK;
}

int ipm_set_refinement(ipm_handle* h, double thresh) {
    if (!h) return IPM_ERR_ARG;
    h->refine_thresh = (thresh >= 0.0) ? thresh : -1.0;
    return IPM_OK;
}

int ipm_detect_dependent_rows(ipm_handle* h, double rel_tol, int* n_dependent) {
    H_TRY(check_handle(h));
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }      // the captured iteration holds the mask pointer
    if (h->dep) { H_CUDA(cudaFree(h->dep)); h->dep = nullptr; h->n_dep = 0; }
    if (n_dependent) *n_dependent = 0;
    if (!(rel_tol > 0.0)) return IPM_OK;                  // off
    const int m = h->m, n = h->n;
    H_CUDA(cudaMalloc(&h->dep, (size_t)m));
    H_CUDA(cudaMemsetAsync(h->dep, 0, (size_t)m, h->st));
    // M = A A^T (d = 1): as well scaled as M will ever be, so a pivot at round-off level means "this row is a
    // combination of the rows before it"
    k_fill<<<vec_grid(n), VEC_NT, 0, h->st>>>(h->d, n, 1.0);
    count_launch();
    H_TRY(assemble_step(h));
    H_TRY(factor_step(h, rel_tol, 2));
    std::vector<unsigned char> host((size_t)m);
    H_CUDA(cudaMemcpyAsync(host.data(), h->dep, (size_t)m, cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    int cnt = 0;
    for (unsigned char v : host) cnt += v ? 1 : 0;
    h->n_dep = cnt;
    if (cnt == 0) { H_CUDA(cudaFree(h->dep)); h->dep = nullptr; }      // full row rank: nothing to carry around
    if (n_dependent) *n_dependent = cnt;
    h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    return IPM_OK;
}

int ipm_set_state(ipm_handle* h, const double* x, const double* y, const double* s) {
    H_TRY(check_handle(h));
    if (!x || !y || !s) return fail(h, IPM_ERR_ARG, "null pointer");
    H_CUDA(cudaMemcpyAsync(h->x, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->y, y, h->m * sizeof(double), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaMemcpyAsync(h->s, s, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    return IPM_OK;
}

int ipm_get_state(ipm_handle* h, double* x, double* y, double* s) {
    H_TRY(check_handle(h));
    if (x) H_CUDA(cudaMemcpyAsync(x, h->x, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (y) H_CUDA(cudaMemcpyAsync(y, h->y, h->m * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (s) H_CUDA(cudaMemcpyAsync(s, h->s, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_residual_norms(ipm_handle* h, double out[5]) {
    H_TRY(check_handle(h));
    H_TRY(residual_step(h));
    H_TRY(fetch_scal(h));
    if (out) {
        out[0] = h->h_scal[S_NRB]; out[1] = h->h_scal[S_NRC]; out[2] = h->h_scal[S_XS];
        out[3] = h->h_scal[S_NB]; out[4] = h->h_scal[S_NC];
    }
    return IPM_OK;
}

int ipm_get_residuals(ipm_handle* h, double* rb, double* rc) {
    H_TRY(check_handle(h));
    if (!h->have_resid) return fail(h, IPM_ERR_STATE, "call ipm_residual_norms first");
    if (rb) H_CUDA(cudaMemcpyAsync(rb, h->rb, h->m * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (rc) H_CUDA(cudaMemcpyAsync(rc, h->rc, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_assemble_normal(ipm_handle* h) {
    H_TRY(check_handle(h));
    if (!h->have_resid) return fail(h, IPM_ERR_STATE, "call ipm_residual_norms first (it forms d = x/s)");
    return assemble_step(h);
}

int ipm_get_M(ipm_handle* h, double* M_rowmajor) {
    H_TRY(check_handle(h));
    if (!M_rowmajor) return fail(h, IPM_ERR_ARG, "null pointer");
    if (!h->have_M && !h->have_factor) return fail(h, IPM_ERR_STATE, "M not assembled");
    H_CUDA(cudaMemcpy2DAsync(M_rowmajor, (size_t)h->m * sizeof(double), h->M, h->ldm * sizeof(double),
                             (size_t)h->m * sizeof(double), h->m, cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_factor(ipm_handle* h, double pivot_rel_thresh, int* n_fixed) {
    H_TRY(check_handle(h));
    if (!h->have_M) return fail(h, IPM_ERR_STATE, "call ipm_assemble_normal first");
    H_TRY(factor_step(h, pivot_rel_thresh));
    if (n_fixed) {
        H_TRY(fetch_scal(h));
        *n_fixed = (int)h->h_scal[S_NFIXED];
    }
    return IPM_OK;
}

int ipm_direction(ipm_handle* h, int kind, double* dx, double* dy, double* ds) {
    H_TRY(check_handle(h));
    if (kind != 0 && kind != 1) return fail(h, IPM_ERR_ARG, "kind must be 0 or 1");
    if (!h->have_resid || !h->have_factor) return fail(h, IPM_ERR_STATE, "need residuals and a factorisation");
    if (kind == 1 && (!h->have_pred || !h->have_sigma))
        return fail(h, IPM_ERR_STATE, "corrector needs the predictor direction and ipm_sigma");
    H_TRY(direction_step(h, kind));
    const double* sx = kind ? h->dx : h->dxa;
    const double* sy = kind ? h->dy : h->dya;
    const double* ss = kind ? h->ds : h->dsa;
    if (dx) H_CUDA(cudaMemcpyAsync(dx, sx, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (dy) H_CUDA(cudaMemcpyAsync(dy, sy, h->m * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (ds) H_CUDA(cudaMemcpyAsync(ds, ss, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_ratio_test(ipm_handle* h, int kind, double eta, double alpha[2]) {
    H_TRY(check_handle(h));
    if (kind != 0 && kind != 1) return fail(h, IPM_ERR_ARG, "kind must be 0 or 1");
    if ((kind == 0 && !h->have_pred) || (kind == 1 && !h->have_corr))
        return fail(h, IPM_ERR_STATE, "direction not computed");
    H_TRY(fetch_scal(h));
    double ap = h->h_scal[kind ? S_AP : S_AP_AFF], ad = h->h_scal[kind ? S_AD : S_AD_AFF];
    if (kind == 1 && eta != h->eta) {
        // the fused kernel applied the handle's eta (0.91, main.py:607); redo the scalar step with the caller's
        ap = std::fmin(1.0, eta * h->h_scal[S_RAW_P]);
        ad = std::fmin(1.0, eta * h->h_scal[S_RAW_D]);
        const double v[2] = {ap, ad};
        H_CUDA(cudaMemcpyAsync(h->scal + S_AP, v, 2 * sizeof(double), cudaMemcpyHostToDevice, h->st));
        H_CUDA(cudaStreamSynchronize(h->st));
    }
    if (alpha) { alpha[0] = ap; alpha[1] = ad; }
    return IPM_OK;
}

int ipm_sigma(ipm_handle* h, double out[3]) {
    H_TRY(check_handle(h));
    if (!h->have_pred) return fail(h, IPM_ERR_STATE, "predictor direction not computed");
    H_TRY(sigma_step(h));
    H_TRY(fetch_scal(h));
    if (out) { out[0] = h->h_scal[S_MU_AFF]; out[1] = h->h_scal[S_MU]; out[2] = h->h_scal[S_SIGMA]; }
    return IPM_OK;
}

int ipm_update(ipm_handle* h, double alpha_p, double alpha_d) {
    H_TRY(check_handle(h));
    if (!h->have_corr) return fail(h, IPM_ERR_STATE, "corrector direction not computed");
    if (!(alpha_p >= 0.0) || !(alpha_d >= 0.0)) return fail(h, IPM_ERR_ARG, "step lengths must be >= 0");
    H_TRY(update_step(h, alpha_p, alpha_d));
    H_CUDA(cudaStreamSynchronize(h->st));
    return IPM_OK;
}

int ipm_solve(ipm_handle* h, double tol, int max_iter, int y0_is_one, double* x, double* y, double* s, double* obj,
              int* iters, int* status, double resid[5]) {
    H_TRY(check_handle(h));
    if (max_iter < 0) return fail(h, IPM_ERR_ARG, "max_iter < 0");
    h->tol = tol;
    if (y0_is_one == IPM_START_KEEP) {            // iterate set by ipm_set_state / ipm_start_mehrotra
        h->have_resid = h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    } else if (y0_is_one == IPM_START_MEHROTRA) {
        H_TRY(ipm_start_mehrotra(h));
    } else {
        H_TRY(ipm_init_state(h, y0_is_one));
    }
    int k = 0;
    const double tau = h->tau;
    auto body = [&]() -> int {
        H_TRY(assemble_step(h));                 // main.py:223-224
        H_TRY(factor_step(h, tau));              // main.py:176-182 (one factorisation per iteration)
        H_TRY(direction_step(h, 0));             // main.py:783
        H_TRY(sigma_step(h));                    // main.py:795
        H_TRY(direction_step(h, 1));             // main.py:799
        H_TRY(update_step(h, -1.0, -1.0));       // main.py:803
        H_TRY(residual_step(h));                 // check_optimality of the new iterate, main.py:780
        H_CUDA(cudaMemcpyAsync(h->h_scal, h->scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->st));
        return IPM_OK;
    };
    if (h->gexec && (h->g_tol != tol || h->g_tau != tau || h->g_dep != (h->dep != nullptr) ||
                     h->g_refine != h->refine_thresh)) {
        cudaGraphExecDestroy(h->gexec);
        h->gexec = nullptr;
    }
    H_TRY(residual_step(h));                     // check_optimality, main.py:780
    H_TRY(fetch_scal(h));
    // Small sparse LPs: the whole loop in one launch of one CTA (small_lp.cuh) - same device functions, same block
    // size, bitwise the same iterates as the loop below, without 21 kernel launches and a host round trip per iteration.
    // (one CTA also forms M: beyond ~10^4 products the many-CTA SpGEMM of the loop below wins - E226, 17.7 k terms:
    // 228 against 215 us per iteration; BOEING2, 8.3 k: 213 against 234)
    if (small_lp_fused().load() != 0 && !h->dense && h->n <= SMALL_MAX_N && h->m <= SMALL_MAX_M && h->dep == nullptr &&
        h->pat && h->pat->nterms <= SMALL_MAX_TERMS &&
        h->refine_thresh < 0.0 && !h->pipe.Linv && (h->ldm % 2 == 0) && h->h_scal[S_CONT] > 0.5 && max_iter > 0) {
        H_TRY(ensure_dyn_smem(k_small_solve, small_smem_bytes_max()));
        SmallArgs sa;
        sa.m = h->m; sa.n = h->n; sa.nnz = h->nnz; sa.nent = h->nent; sa.ldm = h->ldm;
        sa.rowptr = h->rowptr; sa.colind = h->colind; sa.t_rowptr = h->t_rowptr; sa.t_colind = h->t_colind;
        sa.val = h->val; sa.t_val = h->t_val; sa.ad = h->ad;
        sa.out_idx = h->out_idx; sa.prod_ptr = h->prod_ptr; sa.pa = h->pa; sa.pb = h->pb;
        sa.b = h->b; sa.c = h->c; sa.x = h->x; sa.y = h->y; sa.s = h->s; sa.rb = h->rb; sa.rc = h->rc; sa.d = h->d;
        sa.w = h->w; sa.rcx = h->rcx; sa.dxa = h->dxa; sa.dya = h->dya; sa.dsa = h->dsa; sa.dx = h->dx; sa.dy = h->dy;
        sa.ds = h->ds; sa.tm = h->tm; sa.tn = h->tn; sa.rhs = h->rhs; sa.M = h->M;
        sa.scal = h->scal; sa.partials = h->partials; sa.counter = h->counter;
        sa.tol = tol; sa.eta = h->eta; sa.tau = tau; sa.max_iter = max_iter;
        sa.k_out = reinterpret_cast<int*>(h->counter + 1);
        sa.vec_off = (int)(small_work_bytes(h->m) / sizeof(double));
        static const bool want_prof = getenv("IPM_SMALL_PROF") != nullptr;      // tools/small_lp_rate.py: phase cycles
        long long* prof_d = nullptr;
        if (want_prof) {
            H_CUDA(cudaMalloc(&prof_d, 16 * sizeof(long long)));
            H_CUDA(cudaMemsetAsync(prof_d, 0, 16 * sizeof(long long), h->st));
        }
        sa.prof = prof_d;
        k_small_solve<<<1, VEC_NT, small_smem_bytes(h->m, h->n), h->st>>>(sa);
        count_launch();
        H_TRY(launch_check());
        if (prof_d) {
            long long hp[16];
            H_CUDA(cudaMemcpyAsync(hp, prof_d, sizeof(hp), cudaMemcpyDeviceToHost, h->st));
            H_CUDA(cudaStreamSynchronize(h->st));
            cudaFree(prof_d);
            static const char* nm[10] = {"assemble", "cholesky", "make_w", "A w + rhs", "solves", "A^T dy", "direction", "sigma",
                                         "update", "residual check"};
            fprintf(stderr, "k_small_solve m=%d n=%d cycles:", h->m, h->n);
            for (int i = 0; i < 10; ++i) fprintf(stderr, " %s %lld |", nm[i], hp[i]);
            fprintf(stderr, "\n");
        }
        H_CUDA(cudaMemcpyAsync(h->h_scal, h->scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->st));
        H_CUDA(cudaMemcpyAsync(&k, h->counter + 1, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        H_CUDA(cudaStreamSynchronize(h->st));
    }
    while (h->h_scal[S_CONT] > 0.5 && k < max_iter) {
        if (h->use_graph && k >= 1) {
            // iteration 0 ran eagerly (it also sets the kernels' shared-memory attributes); from here on the
            // whole iteration, next residual check and the read-back of the scalars are one graph launch
            if (!h->gexec) {
                cudaGraph_t graph = nullptr;
                H_CUDA(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
                const int64_t launches_before = g_launches.load();
                const int rc_body = body();
                h->launches_per_graph = g_launches.load() - launches_before;
                cudaError_t ce = cudaStreamEndCapture(h->st, &graph);
                if (rc_body != IPM_OK) { if (graph) cudaGraphDestroy(graph); return rc_body; }
                if (ce != cudaSuccess) return fail(h, IPM_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
                ce = cudaGraphInstantiate(&h->gexec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) return fail(h, IPM_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
                h->g_tol = tol; h->g_tau = tau; h->g_dep = h->dep != nullptr; h->g_refine = h->refine_thresh;
            } else {
                count_launch(h->launches_per_graph);
            }
            H_CUDA(cudaGraphLaunch(h->gexec, h->st));
            H_CUDA(cudaStreamSynchronize(h->st));
        } else {
            H_TRY(body());
            H_CUDA(cudaStreamSynchronize(h->st));
        }
        ++k;
    }
    h->have_resid = true;
    h->have_M = h->have_factor = h->have_pred = h->have_sigma = h->have_corr = false;
    const double* hs = h->h_scal;
    const bool finite = std::isfinite(hs[S_NRB]) && std::isfinite(hs[S_NRC]) && std::isfinite(hs[S_XS]) &&
                        std::isfinite(hs[S_OBJ]);
    int st = IPM_STATUS_CONVERGED;
    if (!finite) st = IPM_STATUS_NAN;
    else if (hs[S_CONT] > 0.5) st = IPM_STATUS_MAX_ITER;
    if (obj) *obj = hs[S_OBJ];
    if (iters) *iters = k;
    if (status) *status = st;
    if (resid) {
        resid[0] = hs[S_NRB]; resid[1] = hs[S_NRC]; resid[2] = hs[S_XS]; resid[3] = hs[S_NB]; resid[4] = hs[S_NC];
    }
    return ipm_get_state(h, x, y, s);
}

// ------------------------------------------------------------------------------------------------- stand-alone kernels
int ipm_syrk_d(int device_ordinal, int m, int n, const double* A_d, int64_t lda, const double* d_d, double* M_d,
               int64_t ldm) {
    if (!A_d || !M_d) return IPM_ERR_ARG;
    if (m <= 0 || n <= 0 || lda < n || ldm < m) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    DmmaArgs g;
    g.P = A_d; g.ldp = lda; g.strideP = 0;
    g.Q = A_d; g.ldq = lda; g.strideQ = 0;
    g.dvec = d_d; g.strideD = 0;
    g.C = M_d; g.ldc = ldm; g.strideC = 0;
    g.rowsP = m; g.rowsQ = m; g.K = n; g.lower_only = 1; g.active = nullptr;
    IPM_TRY((dmma_syrk_auto<0>(g, 1, 0)));
    return IPM_OK;
}

int ipm_set_syrk_stage_width(int columns) {
    if (columns != 16 && columns != 32) return IPM_ERR_ARG;
    ws_stage_width().store(columns);
    return IPM_OK;
}

int ipm_set_syrk_consumers(int warps) {
    if (warps != 8 && warps != 16) return IPM_ERR_ARG;
    ws_consumer_warps().store(warps);
    return IPM_OK;
}

int ipm_set_small_lp_fused(int on) {
    small_lp_fused().store(on != 0);
    return IPM_OK;
}

int ipm_set_chol_fused_diag(int on) {
    chol_fused_diag().store(on != 0);
    return IPM_OK;
}

int ipm_potrf_d(int device_ordinal, int m, double* M_d, int64_t ldm, double pivot_rel_thresh, int* n_fixed) {
    if (!M_d) return IPM_ERR_ARG;
    if (m <= 0 || ldm < m || (ldm & 1) || (reinterpret_cast<uintptr_t>(M_d) & 15)) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    double* scal = nullptr;
    IPM_CUDA_OK(cudaMalloc(&scal, S_COUNT * sizeof(double)));
    int rc = potrf_single_auto(M_d, ldm, m, scal, pivot_rel_thresh, 0);
    double hs[S_COUNT];
    if (rc == IPM_OK) {
        cudaError_t e = cudaMemcpy(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); rc = IPM_ERR_CUDA; }
        else if (n_fixed) *n_fixed = (int)hs[S_NFIXED];
    }
    cudaFree(scal);
    return rc;
}

}  // extern "C"
