/*
 * ipm_b200.h — C ABI of the B200-native Newton-step hot path.
 *
 * Drop-in boundary for the predictor-corrector interior-point LP solver
 * payakorn/InteriorPointMethod.  The reference has no FFI of its own: its seams are the
 * plain Python functions of `main.py` listed beside each entry point below (file:line are
 * into the reference tree).  The reference-side binding a maintainer would add (a ctypes
 * stub that rebinds those names) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative IPM_ERR_* code; nothing throws
 *     across the ABI.  Numerical breakdown is reported through `status`, never as an
 *     error (the reference returns NaN, it does not raise: main.py:180, 780, 812).
 *   - all pointers are HOST pointers unless the name ends in `_d` (device pointers in
 *     the handle's device).  Host buffers are borrowed for the duration of the call.
 *   - all reals are IEEE float64, all indices int32; matrices are ROW-major.
 *   - a handle binds one GPU and one stream and is not thread-safe; distinct handles
 *     may be used from distinct threads.
 *   - there is no CPU fallback: without a usable CUDA device ipm_create fails with
 *     IPM_ERR_CUDA.
 */
#ifndef IPM_B200_H
#define IPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipm_handle ipm_handle;

enum {
    IPM_OK = 0,
    IPM_ERR_CUDA = -1,      /* CUDA runtime failure; text in ipm_last_error */
    IPM_ERR_ARG = -2,       /* null pointer / bad enum */
    IPM_ERR_SHAPE = -3,     /* inconsistent m, n, nnz, lda, indices out of range */
    IPM_ERR_STATE = -4,     /* call order violated (e.g. ipm_direction before ipm_factor) */
    IPM_ERR_NOMEM = -5
};

/* status values written by the solve-level calls (reference behaviour in parentheses) */
enum {
    IPM_STATUS_CONVERGED = 0,   /* check_optimality turned False (main.py:169-173) */
    IPM_STATUS_MAX_ITER = 1,    /* hit the cap: 5000 in interior_sparse (main.py:780), 50000 in interior (main.py:725) */
    IPM_STATUS_NAN = 2          /* a non-finite iterate: the reference loop exits because every
                                   comparison with NaN is False (main.py:170-173, 780) */
};

/* ---------------------------------------------------------------- lifetime */
int  ipm_create(ipm_handle **h, int device_ordinal);
void ipm_destroy(ipm_handle *h);
const char *ipm_last_error(const ipm_handle *h);   /* h may be NULL: last error of a failed ipm_create / batched call */
const char *ipm_version(void);

/* Kernels launched by this library in this process since load (bench.py "gpu_launches"). */
int64_t ipm_launch_count(void);

/* Relative pivot threshold of the safeguarded Cholesky used by ipm_solve (default 1e-30). */
int ipm_set_pivot_threshold(ipm_handle *h, double pivot_rel_thresh);

/* ---------------------------------------------------------------- problem data
 * Replaces what create_problem_from_mps hands to the drivers (sparse_interior.py:211-216):
 * A (scipy CSC there; ipm_load_csc takes it as it is, ipm_load_csr takes the row-compressed form), b (m), c (n). */
int ipm_load_csr(ipm_handle *h, int m, int n, int64_t nnz,
                 const int32_t *rowptr, const int32_t *colind, const double *val,
                 const double *b, const double *c);
/* The same problem handed over in the orientation the reference's loader holds it in: `create_problem_from_mps`
 * returns the scipy csc_matrix that loadmat read from benchmarks/<name>.mat (sparse_interior.py:139-167, 211-216):
 * colptr (n+1), rowind (nnz, strictly ascending inside a column), val (nnz).  No host-side conversion is needed.
 *
 * Both loaders do all structure-dependent work ON THE DEVICE: the other orientation by a
 * stable transposition, and the symbolic pattern of the lower triangle of M = A diag(d) A^T (entry list + term
 * lists ordered by the shared column index, the summation order of scipy's csr_matmat for main.py:224).  The
 * result is cached per (device, orientation, m, n, structure arrays) in a process-wide LRU cache: loading the same
 * structure again - the same LP, or new values / right-hand sides on the same pattern - only uploads val, b, c. */
int ipm_load_csc(ipm_handle *h, int m, int n, int64_t nnz,
                 const int32_t *colptr, const int32_t *rowind, const double *val,
                 const double *b, const double *c);
/* host_symbolic != 0: build the SpGEMM pattern with the host routine instead (cross-check of the device pass; it
 * is also what orders m beyond the 224 KB shared-memory marker, m > 57344, fall back to).  use_cache == 0: bypass
 * the pattern cache.  Process-wide; default (0, 1). */
int ipm_set_ingest_mode(int host_symbolic, int use_cache);
/* out = { entries of tril(M), terms (scalar products) of the numeric SpGEMM, 1 if the last load hit the cache,
 *         1 if the pattern was built on the device, build time of the pattern in us, time of the last load in us } */
int ipm_pattern_info(ipm_handle *h, int64_t out[6]);
/* Debug/parity: copies of the device-resident structure (any pointer may be NULL): CSR of A (m+1, nnz), CSR of
 * A^T = CSC of A (n+1, nnz), and the pattern (out_idx[nent] = i*ldm + j with ldm = 16*ceil(m/16),
 * prod_ptr[nent+1], pa/pb[nterms] = positions in the CSR value array of the two factors of every term). */
int ipm_get_pattern(ipm_handle *h, int32_t *rowptr, int32_t *colind, int32_t *t_rowptr, int32_t *t_colind,
                    int64_t *out_idx, int64_t *prod_ptr, int32_t *pa, int32_t *pb);
/* Debug/parity: the value arrays in CSR order and in CSC order as the device holds them. */
int ipm_get_values(ipm_handle *h, double *val_csr, double *val_csc);
/* out = { cached structures, hits, misses, device bytes held by the cache } */
int ipm_pattern_cache_stats(int64_t out[4]);
int ipm_pattern_cache_clear(void);
/* Dense A, row-major with leading dimension lda >= n (the `interior` caller, main.py:707-757). */
int ipm_load_dense(ipm_handle *h, int m, int n, const double *A, int64_t lda,
                   const double *b, const double *c);
/* Dense A already resident in this handle's device (borrowed, must outlive the handle's use). */
int ipm_load_dense_d(ipm_handle *h, int m, int n, const double *A_d, int64_t lda,
                     const double *b_d, const double *c_d);

/* ---------------------------------------------------------------- iterate */
/* x = s = 1; y = 1 (initial_vector_sparse, sparse_interior.py:193-200) or y = 0 (initial_vector, main.py:287-302). */
int ipm_init_state(ipm_handle *h, int y0_is_one);
/* NOT in the reference (opt-in; iteration parity does not apply, objective parity does): dependent rows of a
 * rank-deficient A.  The reference's loader hands over A as it is in the file; QAP8/12/15 are 13-20 % rank deficient
 * (SURVEY App. C.3), and on such a matrix the pivots of the dependent rows of M = A D A^T are round-off - sometimes
 * below the 1e-30 threshold of the safeguard, sometimes not - so the iteration crawls (QAP8: 196 iterations) or
 * stalls short of the optimum (QAP15: gap 1e-3 after 500).  This call factors A A^T (d = 1, where M is as well
 * scaled as it will ever be) with the relative pivot threshold rel_tol (1e-10 is the tested value): a row whose
 * pivot falls below it is a combination of the rows before it.  The mask stays in the handle until the next load,
 * and every later factorisation replaces the pivots of the masked rows by 1e128, which removes the row from the
 * normal equations (dy_i = 0) - the LIPSOL / PCx treatment.  rel_tol <= 0 clears the mask.  Call after a load and
 * before ipm_solve / ipm_start_mehrotra.  QAP8 then needs 17-19 iterations, QAP15 reaches the Netlib optimum. */
int ipm_detect_dependent_rows(ipm_handle *h, double rel_tol, int *n_dependent);

/* NOT in the reference (opt-in): conditional refinement of the corrector for ONE LP, the rule the batched solver
 * applies by default (IPM_BOPT_REFINE above): delta = -rb - A dx; when |delta| > thresh |rb| one step of iterative
 * refinement on the same factor (M ddy = delta, dy += ddy, dx and ds re-formed).  thresh < 0 switches it off
 * (default: the 26 LPs the reference converges on are reproduced without it).  thresh = 1 is the batched
 * solver's rule; smaller values refine more often (0 = always). */
int ipm_set_refinement(ipm_handle *h, double thresh);

/* NOT in the reference (opt-in): Mehrotra's starting point, x = A^T (A A^T)^-1 b and
 * s = c - A^T y with y = (A A^T)^-1 A c, both shifted into the positive orthant (SIAM J. Optim. 2 (1992) sec. 7),
 * computed on the device with the solver's own SYRK/SpGEMM, Cholesky and triangular-solve kernels.  It changes the
 * iteration count (AFIRO 93 -> 15) and lets the LPs converge on which the reference's start x = s = 1 fails
 * (25FV47: NaN after 331 iterations -> Netlib optimum in 33), so iteration parity with the reference does not
 * apply to it. */
int ipm_start_mehrotra(ipm_handle *h);
int ipm_set_state(ipm_handle *h, const double *x, const double *y, const double *s);
int ipm_get_state(ipm_handle *h, double *x, double *y, double *s);

/* ---------------------------------------------------------------- op level (parity tests mirror the Python seams) */
/* out = { |A x - b|_2, |A^T y + s - c|_2, x^T s, |b|_2, |c|_2 } — the quantities check_optimality
 * compares (main.py:169-172).  Also refreshes the residual vectors rb, rc kept in the handle. */
int ipm_residual_norms(ipm_handle *h, double out[5]);
int ipm_get_residuals(ipm_handle *h, double *rb, double *rc);
/* M = A diag(x/s) A^T (main.py:223-224): SpGEMM (CSR input) or DMMA SYRK (dense input). */
int ipm_assemble_normal(ipm_handle *h);
/* Debug/parity: full m x m row-major copy; lower triangle is meaningful (M before ipm_factor, L after). */
int ipm_get_M(ipm_handle *h, double *M_rowmajor);
/* Safeguarded blocked Cholesky of M in place, replacing solve_linear (main.py:176-182):
 * pivot <= pivot_rel_thresh * max diag(M) (or NaN) -> 1e128. */
int ipm_factor(ipm_handle *h, double pivot_rel_thresh, int *n_fixed);
/* kind 0: predictor, rhs [-rc;-rb;-x*s]         (direction_predicted_sparse, main.py:197-229)
 * kind 1: corrector, rhs [-rc;-rb;-(x*s+dxa*dsa-sigma*mu)] using the predictor direction and the
 *         sigma, mu stored by ipm_sigma           (direction_corrected_sparse, main.py:247-277; create_rhs_corrected main.py:142-159)
 * Needs ipm_residual_norms + ipm_assemble_normal + ipm_factor for the current iterate.
 * dx, dy, ds may be NULL (results stay on the device). */
int ipm_direction(ipm_handle *h, int kind, double *dx, double *dy, double *ds);
/* kind 0: predicted_stepsize (main.py:305-322) on the predictor direction, eta ignored.
 * kind 1: full_stepsize (main.py:604-626) on the corrector direction: min(1, eta*min(...)).
 * alpha = { alpha_primal, alpha_dual }. */
int ipm_ratio_test(ipm_handle *h, int kind, double eta, double alpha[2]);
/* out = { mu_aff, mu, sigma } (duality_gap, main.py:588-601; predicted, main.py:562-585). */
int ipm_sigma(ipm_handle *h, double out[3]);
/* x += alpha_p dx; y += alpha_d dy; s += alpha_d ds with the corrector direction (corrected, main.py:694-696). */
int ipm_update(ipm_handle *h, double alpha_p, double alpha_d);

/* ---------------------------------------------------------------- stateless op level (HOST vectors)
 * Pure-function seams of the reference that take no matrix.  Each call uploads its vectors, runs the same
 * kernels the solver uses and copies the result back: for parity tests and reference-style driver loops.
 * ipm_op_ratio_test: eta <= 0 -> predicted_stepsize (main.py:305-322); eta > 0 -> full_stepsize (main.py:604-626).
 * ipm_op_sigma: out = { mu_aff, mu, sigma } (predicted + duality_gap, main.py:562-601).
 * ipm_op_update: x += ap dx; y += ad dy; s += ad ds in place (corrected, main.py:694-696).
 * ipm_solve_spd: z = M^-1 rhs by the safeguarded Cholesky (solve_linear on main.py:226's matrix, main.py:176-182). */
int ipm_op_ratio_test(int device_ordinal, int n, const double *x, const double *dx, const double *s,
                      const double *ds, double eta, double alpha[2]);
/* step_size (main.py:325-547) = predicted_stepsize_lb_ub (main.py:550-559) / full_stepsize_lb_ub (main.py:629-660): the
 * ratio test with simple bounds lb <= x <= ub kept implicit (either may be NULL: x >= 0 only on that side, the four
 * cases of the reference).  eta <= 0: predictor step lengths; eta > 0 (0.91 in the reference): corrector,
 * min(eta * ratio, 1).  The reference's quirks are kept: an empty index set gives 1 BEFORE the eta scaling where the
 * reference does so, and the corrector without bounds returns alpha_dual = 1 (main.py:449-454 reads an unbound name
 * and swallows the error).  alpha = { alpha_primal, alpha_dual }. */
int ipm_op_step_size_bounded(int device_ordinal, int n, const double *x, const double *dx, const double *s,
                             const double *ds, const double *lb, const double *ub, double eta, double alpha[2]);
int ipm_op_sigma(int device_ordinal, int n, const double *x, const double *s, const double *dx_aff,
                 const double *ds_aff, double out[3]);
int ipm_op_update(int device_ordinal, int m, int n, double *x, double *y, double *s, const double *dx,
                  const double *dy, const double *ds, double alpha_p, double alpha_d);
int ipm_solve_spd(int device_ordinal, int m, const double *M_rowmajor, const double *rhs, double pivot_rel_thresh,
                  double *z, int *n_fixed);

/* ---------------------------------------------------------------- solve level
 * Whole predictor-corrector loop on the device (interior_sparse main.py:760-815 when the problem
 * was loaded with ipm_load_csr, interior main.py:707-757 when loaded dense).  Starts from
 * ipm_init_state(y0_is_one) for y0_is_one = 0 / 1 (the reference's two drivers); IPM_START_KEEP and
 * IPM_START_MEHROTRA select the other starting points.  e1 = e2 = e3 = tol as in both drivers.
 * Outputs (each may be NULL): x (n), y (m), s (n), obj = c^T x (the caller subtracts cTlb),
 * iters, status, resid = { |rb|, |rc|, x^T s, |b|, |c| } at exit. */
enum {
    IPM_START_Y0 = 0,        /* x = s = 1, y = 0   (initial_vector, main.py:287-302) */
    IPM_START_Y1 = 1,        /* x = s = 1, y = 1   (initial_vector_sparse, sparse_interior.py:193-200) */
    IPM_START_KEEP = 2,      /* the iterate already in the handle (ipm_set_state / ipm_start_mehrotra) */
    IPM_START_MEHROTRA = 3   /* ipm_start_mehrotra first; not in the reference */
};
int ipm_solve(ipm_handle *h, double tol, int max_iter, int y0_is_one,
              double *x, double *y, double *s, double *obj, int *iters, int *status, double resid[5]);

/* Batch of B independent dense LPs of one shape, contiguous A[B][m][n], b[B][m], c[B][n].
 * HOST buffers: the call stages chunks host->device on a copy stream overlapped with the solve
 * of the previous chunk.  x (B*n) may be NULL.  Start y = 0 (dense driver convention). */
int ipm_solve_batched_dense(int device_ordinal, int B, int m, int n,
                            const double *A, const double *b, const double *c,
                            double tol, int max_iter,
                            double *obj, int *iters, int *status, double *x);
/* Frees what the batched entry points keep between calls: the device/pinned staging buffers of ipm_solve_batched_dense
 * and the per-device pool of loop resources (a pinned page, events, the streams of the hand-off kernel) both batched
 * entry points draw from.  Not to be called while a batched solve is running on another thread. */
int ipm_release_cached(void);
/* Same with everything resident on `device_ordinal`; stream 0 of that device; synchronises before return.
 * work_d: scratch of ipm_batched_workspace_bytes(B,m,n) bytes or NULL (allocated internally). */
int ipm_solve_batched_dense_d(int device_ordinal, int B, int m, int n,
                              const double *A_d, const double *b_d, const double *c_d,
                              double tol, int max_iter,
                              double *obj_d, int *iters_d, int *status_d, double *x_d,
                              void *work_d, int *iterations_run);
int64_t ipm_batched_workspace_bytes(int B, int m, int n);
/* Iteration variant of the batched solver.
 * three_pass = 1 (default), m <= 256: the critical path of an iteration reads A three times (SYRK, predictor
 *   direction, corrector direction): the corrector right-hand side comes from the predictor's by linearity of
 *   main.py:150-152 and the residuals of the new point from the recurrences rb += ap A dx, rc += ad (A^T dy + ds).
 *   check_optimality (main.py:169-173) is re-evaluated from scratch before an LP is declared finished and in
 *   every refresh_every-th iteration (default 12; 0 = only at the end).  The recurrences drift from the true
 *   residuals in the ill-conditioned last iterations (round 1, before an LP's residuals were REPLACED by the
 *   from-scratch values whenever it was checked: LP 7466 of the benchmark batch needed 60 iterations instead of 16
 *   without a periodic refresh).  Measured on all 65536 generator LPs of the frozen table with periods 6, 9, 12 and
 *   18: every LP converges, 65535 in exactly the table's iteration count and one in one more, objectives within
 *   4.1e-9 - the same as with the period of 3 that was the default until late in round 2 and cost four from-scratch
 *   passes over A more per solve (tools/scan_batch_gpu.py --refresh=N, profiles/r2_scan_65536_refresh*.txt).  12 keeps
 *   one refresh just before the ill-conditioned phase of a 15-19 iteration solve.
 * three_pass = 0: six passes, every residual from scratch in every iteration, exactly as main.py:725-751 orders it.
 * Process-wide; for A/B measurements and parity tests. */
int ipm_batched_set_variant(int three_pass, int refresh_every);
/* Further process-wide options of the batched solver (read at the start of every solve).
 * IPM_BOPT_REFINE (default 1): conditional refinement of the corrector.  The reference solves the unreduced
 *   Newton system (main.py:13-21), whose second block row is A dx = -rb; the normal equations (main.py:221-229)
 *   satisfy it only as well as the factor of M = A D A^T allows, and once max(x/s)/min(x/s) passes 1e19 that is
 *   not well enough for |rb| to keep falling: an LP that has not met check_optimality (main.py:169-173) by then
 *   stays trapped at the boundary for thousands of iterations (generator LPs 16893 and 31186; the reference needs
 *   18 iterations on both), and one trapped LP keeps the whole lockstep loop alive.  The corrector pass forms
 *   delta = -rb - A dx anyway (A dx carries the residual recurrence); when |delta| > 0.1 |rb| (and above 1e-3 of
 *   the stopping threshold of |rb|) the LP takes ONE step of iterative refinement on the same factor, applied
 *   incrementally (M ddy = delta, dy += ddy, dx += D A^T ddy, ds re-formed from dx) before it is updated.  Taken
 *   0.06 times per LP on the benchmark generator (3.5 % of its LPs); against the UNMODIFIED reference on 517 of them
 *   the CPU restatement of the rule gives equal iteration counts and objectives within 1.6e-10 relative
 *   (tests/golden/batch_256x512_*.{npz,json}).  0 switches it off (A/B, and to document the trap).
 * IPM_BOPT_STRIP_TMA (default 1): the four-pass direction kernels read the column strips of A straight from the
 *   caller's row-major array through a 3-D tensor map (cp.async.bulk.tensor); 0 = from a strip-major copy of A
 *   made once per solve (costs a second copy of A in the workspace: set it BEFORE ipm_batched_workspace_bytes).
 * IPM_BOPT_HANDOFF (default 1, needs IPM_BOPT_REFINE and n + m <= 1600): when the refined corrector STILL has
 *   |delta| > |rb|, the normal equations have broken down for this LP (the safeguarded Cholesky dropped a row that
 *   is not dependent: numerically degenerate vertex); the LP leaves the lockstep loop and is finished from its
 *   current iterate by ipm_solve_dense_kkt's kernel (augmented system, LU with partial pivoting - the reference's
 *   own dgesv route, main.py:178).  About one generator LP in a thousand; without it such an LP can iterate for
 *   thousands of iterations at |rb| just above the threshold (LPs 16893, 31186, 54456), the reference needs 17-18.
 * Value 2 of IPM_BOPT_REFINE / IPM_BOPT_HANDOFF is a test hook: every corrector takes the refinement step / every LP
 * is handed off after it (tests/test_zz_gpu_refinement.py exercises both paths on whole blocks with them).
 * IPM_BOPT_SYRK_RHS (default 1, four-pass iteration only): the product A w of the predictor right-hand side
 *   (main.py:225, rhs = -rb - A d (rc - rcomp/x)) is formed by the diagonal tiles of the SYRK kernel from the operand
 *   slabs it has in shared memory anyway, instead of by a pass of its own over A; 0 = separate pass (A/B). */
enum { IPM_BOPT_REFINE = 1, IPM_BOPT_STRIP_TMA = 2, IPM_BOPT_HANDOFF = 3, IPM_BOPT_SYRK_RHS = 4 };
int ipm_batched_set_option(int option, int value);
/* LPs the most recent batched solve on this process handed to the augmented-system kernel. */
int ipm_batched_last_handoffs(void);

/* One dense LP by the reference's dense route (`interior`, main.py:707-757: `create_matrix` main.py:13-21 +
 * np.linalg.solve main.py:178) on the GPU: predictor-corrector iteration on the augmented system
 * [[-D^-1, A^T], [A, 0]] (the unreduced KKT matrix with ds eliminated exactly), LU with partial pivoting, one CTA.
 * Start x = s = 1, y = 0 (main.py:287-302).  Host buffers; A row-major m x n, n + m <= 1600.  A robustness path
 * (same Newton system and pivoting rule as the reference => same iteration counts), not a throughput path. */
int ipm_solve_dense_kkt(int device_ordinal, int m, int n, const double *A, const double *b, const double *c,
                        double tol, int max_iter, double *x, double *y, double *s, double *obj, int *iters,
                        int *status);
/* CTAs (= SMs) per LP of that kernel: a thread-block cluster of 1, 2, 4 (default) or 8 CTAs shares the row swaps and the
 * trailing update of every panel; results are bitwise the same for every cluster size.  Process-wide. */
int ipm_set_kkt_cluster(int ctas);
/* Cycle counts of the last ipm_solve_dense_kkt call, by phase (CTA 0): [0] residuals + check, [1] building K, [2] panel
 * loads, [3] panel column steps (pivot search, swap, scale, rank-1 update in shared memory), [4] panel write-back,
 * [5] row interchanges outside the panel, [6] U12 + trailing update + cluster barrier, [7] the two solves,
 * [8] elementwise work, ratio tests, update, [9] iterations.  Diagnostic. */
int ipm_kkt_last_profile(int64_t out[16]);

/* Phase timing of the batched solver (bench.py roofline): CUDA events on the solve stream around the four
 * phases of every lockstep iteration.  ms/calls index: 0 residual pass, 1 SYRK (dmma_nt_kernel, one launch
 * per call), 2 Cholesky, 3 both solves (rhs, triangular solves, direction/update).  lp_iterations = sum over
 * lockstep iterations of the number of LPs still active.  ipm_profile_enable resets the accumulators. */
int ipm_profile_enable(int on);
int ipm_profile_read(double ms[4], int64_t calls[4], int64_t *lp_iterations);
/* Intervals of the most recent batched solve in launch order (phase index as above); returns their number. */
int ipm_profile_last(double *ms, int *phase, int cap);

/* Issue-rate ceiling of DMMA.8x8x4 on this device in TFLOP/s (register-only loop, ~10 ms): the FP64
 * tensor-core peak the SYRK/Cholesky roofline fractions are quoted against. */
double ipm_measure_dmma_peak(int device_ordinal);

/* ---------------------------------------------------------------- stand-alone kernels (roofline benches, parity)
 * C_lower = A diag(d) A^T for a dense row-major device matrix (the SYRK of main.py:224). */
int ipm_syrk_d(int device_ordinal, int m, int n, const double *A_d, int64_t lda, const double *d_d,
               double *M_d, int64_t ldm);
/* Stage width of the SYRK / trailing-update kernel's operand ring: 16 columns x 5 stages (default) or 32 x 3 (half as
 * many stage boundaries per tile).  Process-wide, for A/B measurements; results are bitwise the same. */
int ipm_set_syrk_stage_width(int columns);
/* Consumer warps of that kernel: 8 of 32 x 64 accumulator blocks (default) or 16 of 16 x 64 (four MMA-issuing warps per
 * scheduler instead of two; csrc/dmma_ws16.cuh).  Process-wide, for A/B measurements; results are bitwise the same. */
int ipm_set_syrk_consumers(int warps);
/* Panel kernel of the blocked Cholesky of ONE large matrix (m > 512; solve_linear, main.py:176-182): 1 (default) = the
 * diagonal blocks are factored by the fused kernel of the batched solver (32-wide sub-panels, rank-32 updates on the
 * tensor pipe, look-ahead inside the CTA), 0 = by the round-1 shared-memory kernel.  Process-wide, for A/B. */
int ipm_set_chol_fused_diag(int on);
/* Small sparse LPs (n <= 512, m <= 256, no dependent-row mask, no refinement): ipm_solve runs the whole
 * predictor-corrector loop (main.py:776-815) in ONE launch of one CTA - the kernels of the single-LP path as device
 * functions, bitwise the same iterates - instead of one CUDA-graph replay and one host round trip per iteration.
 * 1 (default) / 0: process-wide switch for A/B measurements. */
int ipm_set_small_lp_fused(int on);
/* In-place safeguarded Cholesky of a dense row-major device matrix (lower). */
int ipm_potrf_d(int device_ordinal, int m, double *M_d, int64_t ldm, double pivot_rel_thresh, int *n_fixed);

/* M_i (lower) = A_i diag(d_i) A_i^T for B contiguous row-major matrices A[B][m][n], d[B][n] (nullable), M[B][m][ldm]:
 * the SYRK launch of the batched solver, asynchronous on stream 0. */
int ipm_syrk_batched_d(int device_ordinal, int B, int m, int n, const double *A_d, const double *d_d,
                       double *M_d, int64_t ldm);

/* In-place safeguarded Cholesky of B row-major device matrices (lower), matrix i at M_d + i*strideM.
 * m <= 256 runs the fused one-CTA-per-matrix kernel of the batched solver. */
int ipm_potrf_batched_d(int device_ordinal, int B, int m, double *M_d, int64_t ldm, int64_t strideM,
                        double pivot_rel_thresh, int *n_fixed_total);

#ifdef __cplusplus
}
#endif
#endif /* IPM_B200_H */
