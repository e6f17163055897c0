"""TEST INFRASTRUCTURE — CPU oracle for the Newton-step hot path.  NOT product code.

A numpy/scipy restatement of the algorithm of payakorn/InteriorPointMethod for the
path named in BASELINE.json (`north_star`).  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the product
package (`interiorpointmethod_b200/`) never does.

Two linear-algebra back ends restate the SAME Newton system:

* ``linear="kkt"``    — the reference AS WRITTEN: unreduced (m+2n) KKT matrix
  assembled and LU-solved twice per iteration (sparse: SuperLU via scipy `spsolve`,
  main.py:198-212, 250-269, sparse_interior.py:57-96; dense: `np.linalg.solve`,
  main.py:185-194, 232-244, 13-21).
* ``linear="normal"`` — the elimination the GPU path uses: M = A diag(x/s) A^T as in the
  reference's own `method="normal"` predictor (main.py:221-229), applied to predictor AND
  corrector (same matrix, r3 -> r4), factorised once per iteration by a Cholesky with the
  LIPSOL-style tiny-pivot safeguard.

Parity pinning: the reference's tests pin no solver output (test.py checks shapes only).
This oracle is pinned against (i) outputs of the unmodified reference run in the build
container and frozen under tests/golden/ by oracle/make_golden.py, (ii) the known answers
in the reference source (ex1 -775 main.py:1253, ex2 -15000 main.py:1261, Netlib optimum
table main.py:1417-1516).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import warnings

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import spsolve

ETA = 0.91          # main.py:607
PIVOT_TAU = 1e-30
PIVOT_BIG = 1e128

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- inputs
def as_column(v):
    """(k,1) float64 column, the reference's vector layout (sparse_interior.py:203-208).

    The loader keeps loadmat's small integer dtypes (uint8/uint16/int16);
    every use in the reference mixes them with float64, so the value semantics are float64.
    """
    v = np.asarray(v)
    return np.ascontiguousarray(v.reshape(-1, 1), dtype=np.float64)


def initial_point(m, n, y0_is_one=True):
    """x = s = 1; y = 1 for the sparse driver (sparse_interior.py:193-200), 0 for the dense one (main.py:287-302)."""
    x = np.ones((n, 1))
    s = np.ones((n, 1))
    y = np.ones((m, 1)) if y0_is_one else np.zeros((m, 1))
    return x, y, s


def mehrotra_start(A, b, c, tau=PIVOT_TAU, forced=None):
    """NOT in the reference: Mehrotra's starting point (SIAM J. Optim. 2 (1992) sec. 7), the oracle of the product's
    opt-in `ipm_start_mehrotra`.  x = A^T (A A^T)^-1 b, y = (A A^T)^-1 A c, s = c - A^T y,
    then dx = max(-1.5 min x, 0), ds likewise, and the second shift 0.5 x^T s / sum(s) resp. / sum(x)."""
    m, n = A.shape
    one = np.ones((n, 1))
    L, _ = cholesky_safeguarded(normal_matrix(A, one, one), tau, forced=forced)
    x = A.T @ solve_with_factor(L, b)
    y = solve_with_factor(L, A @ c)
    s = c - A.T @ y
    x = x + max(-1.5 * float(x.min()), 0.0)
    s = s + max(-1.5 * float(s.min()), 0.0)
    xs = float((x * s).sum())
    sx, ss = float(x.sum()), float(s.sum())
    x = np.maximum(x + 0.5 * xs / max(ss, 1e-300), 1e-10)
    s = np.maximum(s + 0.5 * xs / max(sx, 1e-300), 1e-10)
    return x, y, s


# --------------------------------------------------------------------------- residuals / convergence
def residuals(A, b, c, x, y, s):
    """rb = A x - b, rc = A^T y + s - c  (main.py:67-70)."""
    rb = A @ x - b
    rc = A.T @ y + s - c
    return rb, rc


def residual_norms(A, b, c, x, y, s):
    """The five scalars check_optimality compares (main.py:169-172): |rb|, |rc|, x^T s, |b|, |c|."""
    rb, rc = residuals(A, b, c, x, y, s)
    return (float(np.linalg.norm(rb)), float(np.linalg.norm(rc)), float((x.T @ s)[0, 0]),
            float(np.linalg.norm(b)), float(np.linalg.norm(c)))


def continue_flag(A, b, c, x, y, s, e1, e2, e3):
    """True = NOT optimal yet (main.py:169-173).  Strict '<'; any NaN makes every test False."""
    nrb, nrc, gap, nb, nc = residual_norms(A, b, c, x, y, s)
    return bool(e1 * (1 + nb) < nrb or e2 * (1 + nc) < nrc or e3 < gap)


# --------------------------------------------------------------------------- as-written linear algebra
def kkt_matrix_sparse(A, x, s):
    """[[0, A^T, I], [A, 0, 0], [S, 0, X]] of order m+2n, unknowns [dx; dy; ds]
    (sparse_interior.py:57-96)."""
    m, n = A.shape
    I = sparse.identity(n, format="csc")
    S = sparse.diags(s.ravel(), format="csc")
    X = sparse.diags(x.ravel(), format="csc")
    K = sparse.bmat([[None, A.T, I], [A, None, None], [S, None, X]], format="csc")
    return K


def kkt_matrix_dense(A, x, s):
    """Same matrix, dense (main.py:13-21)."""
    m, n = A.shape
    K = np.zeros((m + 2 * n, m + 2 * n))
    K[0:n, n:n + m] = A.T
    K[0:n, n + m:] = np.eye(n)
    K[n:n + m, 0:n] = A
    K[n + m:, 0:n] = np.diagflat(s)
    K[n + m:, n + m:] = np.diagflat(x)
    return K


def direction_kkt(A, x, s, rb, rc, rcomp, dense=False):
    """Solve K [dx;dy;ds] = [-rc; -rb; -rcomp]  (main.py:101-109, 201-209; main.py:185-194 dense).

    Singular systems: spsolve warns and returns NaN (main.py:180), np.linalg.solve raises.
    """
    m, n = A.shape
    rhs = np.vstack([-rc, -rb, -rcomp])
    if dense:
        sol = np.linalg.solve(kkt_matrix_dense(A, x, s), rhs)
    else:
        sol = spsolve(kkt_matrix_sparse(A, x, s), rhs).reshape(-1, 1)
    return sol[0:n], sol[n:n + m], sol[n + m:]


def direction_augmented(A, x, s, rb, rc, rcomp):
    """The same Newton system with only ds eliminated (ds = -s dx/x - rcomp/x, exact):
        [ -D^-1  A^T ] [dx]   [ -(rc - rcomp/x) ]
        [   A     0  ] [dy] = [       -rb        ]          D = diag(x/s)
    solved by LU with partial pivoting (np.linalg.solve = LAPACK dgesv, the routine the reference's dense path calls
    on its (m+2n) matrix, main.py:178).  Order n + m instead of m + 2n; unlike the normal equations it never pivots
    on the tiny s_j/x_j of the basic variables, so it keeps the primal block row A dx = -rb to working precision on
    (nearly) degenerate LPs - the executable spec of the batched solver's fallback kernel (csrc/kkt_dense.cuh)."""
    A = np.asarray(A.todense()) if sparse.issparse(A) else A
    m, n = A.shape
    K = np.zeros((n + m, n + m))
    K[np.arange(n), np.arange(n)] = -(s / x).ravel()
    K[0:n, n:] = A.T
    K[n:, 0:n] = A
    t = rc - rcomp / x
    sol = np.linalg.solve(K, np.vstack([-t, -rb]))
    dx, dy = sol[0:n], sol[n:]
    ds = (-s * dx / x) - (rcomp / x)
    return dx, dy, ds


# --------------------------------------------------------------------------- normal equations
def normal_matrix(A, x, s):
    """M = A diag(x/s) A^T  (main.py:223-224); returns a dense ndarray."""
    d = (x / s).ravel()
    if sparse.issparse(A):
        AD = A @ sparse.diags(d)
        return np.asarray((AD @ A.T).todense())
    return (A * d) @ A.T


_chol_lib = None


def _load_chol_lib():
    """C helper for the safeguarded Cholesky (oracle/chol_safeguard.c), built on first use."""
    global _chol_lib
    if _chol_lib is not None:
        return _chol_lib
    so = os.path.join(_HERE, "liboracle_chol.so")
    src = os.path.join(_HERE, "chol_safeguard.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        tmp = so + ".tmp%d" % os.getpid()
        subprocess.check_call(["gcc", "-O3", "-mavx2", "-mfma", "-shared", "-fPIC", src, "-o", tmp, "-lm"])
        os.replace(tmp, so)            # atomic: processes that mapped the old file keep their copy
    lib = ctypes.CDLL(so)
    lib.oracle_chol_safeguard.restype = ctypes.c_int
    lib.oracle_chol_safeguard.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_void_p]
    lib.oracle_chol_safeguard_masked.restype = ctypes.c_int
    lib.oracle_chol_safeguard_masked.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    _chol_lib = lib
    return lib


def cholesky_safeguarded_numpy(M, tau=PIVOT_TAU, big=PIVOT_BIG):
    """Unblocked right-looking Cholesky, lower factor, with the tiny-pivot rule:
    pivot p_j <= tau * max_i M_ii  (or NaN)  ->  p_j = big   (SURVEY.md App. A.4).
    Pure-numpy version for small m (the executable spec of the CUDA panel kernel)."""
    L = np.array(M, dtype=np.float64, copy=True)
    m = L.shape[0]
    thresh = tau * np.max(np.diag(M)) if m else 0.0
    nfixed = 0
    for j in range(m):
        p = L[j, j]
        if not (p > thresh):
            p = big
            nfixed += 1
        ljj = np.sqrt(p)
        L[j, j] = ljj
        if j + 1 < m:
            L[j + 1:, j] /= ljj
            col = L[j + 1:, j]
            L[j + 1:, j + 1:] -= np.outer(col, col)
    return np.tril(L), nfixed


def cholesky_safeguarded(M, tau=PIVOT_TAU, big=PIVOT_BIG, forced=None, replaced_out=None):
    """Same rule, C implementation (left-looking, oracle/chol_safeguard.c).  Returns (L lower, n_fixed).
    forced: optional uint8 mask of rows whose pivot is replaced unconditionally (dependent rows of A, see
    detect_dependent_rows); replaced_out: optional uint8 array that receives which pivots were replaced.

    Iteration counts on degenerate LPs are sensitive to the rounding of the factorisation (SC50A: 35 with
    this accumulation order, 41 with the right-looking numpy one at tau=1e-30), so the oracle fixes one
    order; the numpy version stays as the readable spec and cross-check."""
    m = M.shape[0]
    if m == 0:
        return np.zeros((0, 0)), 0
    lib = _load_chol_lib()
    L = np.array(M, dtype=np.float64, order="C", copy=True)
    nfixed = ctypes.c_int(0)
    if forced is None and replaced_out is None:
        rc = lib.oracle_chol_safeguard(L.ctypes.data, m, tau, big, ctypes.byref(nfixed))
    else:
        f = None if forced is None else np.ascontiguousarray(forced, dtype=np.uint8)
        rc = lib.oracle_chol_safeguard_masked(L.ctypes.data, m, tau, big, None if f is None else f.ctypes.data,
                                              None if replaced_out is None else replaced_out.ctypes.data,
                                              ctypes.byref(nfixed))
    if rc != 0:
        raise RuntimeError("oracle_chol_safeguard failed")
    return L, nfixed.value


def detect_dependent_rows(A, rel_tol=1e-10):
    """NOT in the reference (opt-in, the oracle of the product's `ipm_detect_dependent_rows`): rows of A that are
    linear combinations of the rows before them, found by the Cholesky factorisation of A A^T (d = 1, where M is as
    well scaled as it will ever be): pivot <= rel_tol * max diag  =>  dependent.  Their pivots are replaced by
    PIVOT_BIG in every later factorisation, which removes the row from the normal equations (dy_i = 0) - the
    LIPSOL / PCx treatment of rank-deficient constraint matrices.  Without it the pivots of such rows are
    round-off (sometimes above the 1e-30 threshold of the safeguard, sometimes negative) and the iteration
    crawls: QAP8 needs 196 iterations instead of 27."""
    m, n = A.shape
    one = np.ones((n, 1))
    rep = np.zeros(m, dtype=np.uint8)
    cholesky_safeguarded(normal_matrix(A, one, one), rel_tol, replaced_out=rep)
    return rep


def solve_with_factor(L, rhs):
    """dy = (L L^T)^-1 rhs by forward/back substitution."""
    from scipy.linalg import solve_triangular

    z = solve_triangular(L, rhs, lower=True, check_finite=False)
    return solve_triangular(L, z, lower=True, trans="T", check_finite=False)


def direction_normal(A, L, x, s, rb, rc, rcomp):
    """Normal-equations direction (main.py:225-228; corrector = same with r3 -> r4):
    t = rc - rcomp/x; dy = M^-1 (-rb - A (d*t)); dx = d*(A^T dy) + d*t; ds = -s*dx/x - rcomp/x."""
    d = x / s
    t = rc - rcomp / x
    rhs = -rb - A @ (d * t)
    dy = solve_with_factor(L, rhs)
    dx = d * (A.T @ dy) + d * t
    ds = (-s * dx / x) - (rcomp / x)
    return dx, dy, ds


# --------------------------------------------------------------------------- step lengths, sigma, update
def ratio_test(v, dv):
    """min({-v_i/dv_i : dv_i < 0} U {1})  (main.py:308-309, 318-319)."""
    i = dv < 0
    if not np.any(i):
        return 1.0
    with np.errstate(all="ignore"):
        return float(min(np.min(-v[i] / dv[i]), 1.0))


def predicted_stepsize(dx_aff, ds_aff, x, s):
    """(alpha_p^aff, alpha_d^aff)  (main.py:305-322)."""
    return ratio_test(x, dx_aff), ratio_test(s, ds_aff)


def sigma_mu(x, s, dx_aff, ds_aff):
    """mu_aff, mu, sigma = (mu_aff/mu)^3, sigma unclamped  (main.py:582-584, 598-600)."""
    n = x.shape[0]
    ap, ad = predicted_stepsize(dx_aff, ds_aff, x, s)
    x_aff = x + ap * dx_aff
    s_aff = s + ad * ds_aff
    mu_aff = float((x_aff.T @ s_aff)[0, 0]) / n
    mu = float((x.T @ s)[0, 0]) / n
    with np.errstate(all="ignore"):
        sigma = (mu_aff / mu) ** 3
    return mu_aff, mu, sigma


def full_stepsize(x, s, dx, ds, eta=ETA):
    """alpha = min(1, eta * min({-v/dv: dv<0} U {1}))  => alpha <= eta always (main.py:616-623)."""
    return min(1.0, eta * ratio_test(x, dx)), min(1.0, eta * ratio_test(s, ds))


def step_size_bounded(x, s, dx, ds, lb=None, ub=None, corrector=False, eta=ETA):
    """step_size (main.py:325-547) restated case by case: the ratio test with bounds lb <= x <= ub kept implicit.
    `one(...)` = `min(np.append(r, 1)) if any else 1`, `raw(...)` = `min(r) if any else 1` - the reference uses one or the
    other depending on the case, and scales by eta before or after.  Corrector without bounds: alpha_dual = 1 (the
    reference reads the unbound name delta_s_aff at main.py:449 and its bare `except:` returns 1)."""
    x, s, dx, ds = (np.asarray(v, dtype=float).ravel() for v in (x, s, dx, ds))
    lo = None if lb is None else np.asarray(lb, dtype=float).ravel()
    up = None if ub is None else np.asarray(ub, dtype=float).ravel()

    def ratios(num, den, mask):
        with np.errstate(all="ignore"):
            return num[mask] / den[mask]

    neg, pos, dual = dx < 0, dx > 0, ds < 0
    r_neg = ratios((lo - x) if lo is not None else -x, dx, neg)
    r_pos = ratios(up - x, dx, pos) if up is not None else np.empty(0)
    r_dual = ratios(-s, ds, dual)
    one = lambda r: float(min(np.min(r), 1.0)) if r.size else 1.0       # noqa: E731
    raw = lambda r: float(np.min(r)) if r.size else 1.0                  # noqa: E731
    e = eta if corrector else 1.0
    if lo is None and up is None:
        ap = min(1.0, eta * one(r_neg)) if corrector else one(r_neg)
        ad = 1.0 if corrector else one(r_dual)
    elif lo is None:
        ap = min(e * raw(r_neg), e * raw(r_pos), 1.0)
        ad = min(e * raw(r_dual), 1.0)
    elif up is None:
        ap = min(e * raw(r_neg), 1.0)
        ad = min(e * raw(r_dual), 1.0)
    else:
        ap = min(e * one(r_pos), e * one(r_neg), 1.0)
        ad = min(e * one(r_dual), 1.0) if corrector else one(r_dual)
    return ap, ad


def objective(x, c):
    """Python `sum` over rows = sequential left-to-right add (main.py:815)."""
    acc = 0.0
    for v in (x * c).ravel():
        acc += v
    return acc


# --------------------------------------------------------------------------- one iteration / whole solve
def newton_iteration(A, b, c, x, y, s, linear="normal", dense=False, tau=PIVOT_TAU, info=None, refine_thresh=None,
                     forced=None, refine_abs=None, refine_rounds=1, handoff_floor=None):
    """One predictor-corrector iteration (main.py:781-805).  Returns new (x, y, s).

    refine_thresh (normal equations only; what the batched GPU solver does with thresh = 1): the corrector's dy
    comes from the factored M = A D A^T, and in the last iterations (max d / min d > 1e19) that factor no longer
    reproduces the primal block row A dx = -rb of the Newton system the reference solves directly (main.py:13-21).
    When delta = -rb - A dx is LARGER than thresh * |rb| - the step would not reduce the primal residual at all -
    one step of iterative refinement on the same factor is taken: M ddy = delta, dy += ddy, dx and ds re-formed
    from the refined dy."""
    rb, rc = residuals(A, b, c, x, y, s)
    r3 = x * s
    if linear == "normal":
        L, nfixed = cholesky_safeguarded(normal_matrix(A, x, s), tau, forced=forced)
        if info is not None:
            info["pivots_fixed"] = info.get("pivots_fixed", 0) + nfixed
        dxa, dya, dsa = direction_normal(A, L, x, s, rb, rc, r3)
    elif linear == "augmented":
        dxa, dya, dsa = direction_augmented(A, x, s, rb, rc, r3)
    else:
        dxa, dya, dsa = direction_kkt(A, x, s, rb, rc, r3, dense=dense)
    mu_aff, mu, sigma = sigma_mu(x, s, dxa, dsa)
    r4 = r3 + dxa * dsa - sigma * mu * np.ones_like(x)   # main.py:150-152
    if linear == "normal":
        dx, dy, ds = direction_normal(A, L, x, s, rb, rc, r4)
        if refine_thresh is not None or refine_abs is not None:
            if refine_thresh is None:
                refine_thresh = np.inf
            d = x / s
            nrb = np.linalg.norm(rb)
            for _ in range(refine_rounds):
                delta = -rb - A @ dx
                nd = np.linalg.norm(delta)
                if not ((nd > refine_thresh * nrb or (refine_abs is not None and nd > refine_abs))
                        and (handoff_floor is None or nd > handoff_floor)):
                    break
                # INCREMENTAL: the correction d * (A^T ddy) is added to dx.  Re-forming dx = d (A^T dy + t) from the
                # refined dy would bring back the cancellation noise eps * d_max * |t| of the large-d columns, which is
                # exactly what the step is meant to remove (measured: |delta| 24.9 |rb| -> 10.3 |rb| re-formed,
                # -> 1e-3 |rb| incremental, generator LP 31186).
                ddy = solve_with_factor(L, delta)
                dy = dy + ddy
                dx = dx + d * (A.T @ ddy)
                ds = (-s * dx / x) - (r4 / x)
                if info is not None:
                    info["refinements"] = info.get("refinements", 0) + 1
            else:
                # every round taken and the primal block row is still not restored: the normal equations have broken
                # down (safeguarded pivot on a row that is NOT dependent - a numerically degenerate vertex).  Hand
                # the LP to the augmented system from THIS iterate (what the batched solver's fallback kernel does).
                if handoff_floor is not None:
                    nd = np.linalg.norm(-rb - A @ dx)
                    if nd > nrb and nd > handoff_floor:
                        if info is not None:
                            info["handoff"] = True
                        return newton_iteration(A, b, c, x, y, s, linear="augmented", info=info)
    elif linear == "augmented":
        dx, dy, ds = direction_augmented(A, x, s, rb, rc, r4)
    else:
        dx, dy, ds = direction_kkt(A, x, s, rb, rc, r4, dense=dense)
    ap, ad = full_stepsize(x, s, dx, ds)
    if info is not None:
        info["last"] = dict(dx_aff=dxa, dy_aff=dya, ds_aff=dsa, mu_aff=mu_aff, mu=mu, sigma=sigma,
                            dx=dx, dy=dy, ds=ds, alpha=(ap, ad))
    return x + ap * dx, y + ad * dy, s + ad * ds       # main.py:694-696


def solve(A, b, c, cTlb=0.0, tol=1e-8, max_iter=5000, y0_is_one=True, linear="normal", tau=PIVOT_TAU,
          start="reference", refine_thresh=None, dependent_tol=None, refine_abs_kappa=None, refine_rounds=1,
          handoff=False):
    """(see below)  handoff=True (with refine_thresh): the batched GPU solver's rule - when the refined corrector
    still has |delta| > |rb| (and |delta| above 1e-3 of the stopping threshold of the primal residual), this and every
    later iteration of the LP run on the augmented system (`direction_augmented`)."""
    return _solve(A, b, c, cTlb, tol, max_iter, y0_is_one, linear, tau, start, refine_thresh, dependent_tol,
                  refine_abs_kappa, refine_rounds, handoff)


def _solve(A, b, c, cTlb, tol, max_iter, y0_is_one, linear, tau, start, refine_thresh, dependent_tol, refine_abs_kappa,
           refine_rounds, handoff):
    """Whole solve with `interior_sparse` semantics (main.py:760-815) when y0_is_one, `interior`
    semantics (main.py:707-757; cap 50000, y0 = 0) otherwise.

    Returns dict(x, y, s, k, obj, status, pivots_fixed); status 0 converged, 1 max_iter, 2 nan.
    """
    dense = not sparse.issparse(A)
    if dense:
        A = np.ascontiguousarray(A, dtype=np.float64)
    else:
        A = sparse.csr_matrix(A, dtype=np.float64)
    b = as_column(b)
    c = as_column(c)
    m, n = A.shape
    x, y, s = initial_point(m, n, y0_is_one)
    forced = None
    if dependent_tol is not None and linear == "normal":
        forced = detect_dependent_rows(A, dependent_tol)
    if start == "mehrotra":
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                x, y, s = mehrotra_start(A, b, c, tau, forced=forced)
    k = 0
    info = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            floor = 1e-3 * tol * (1.0 + float(np.linalg.norm(b))) if handoff else None
            while continue_flag(A, b, c, x, y, s, tol, tol, tol) and k < max_iter:
                if info.get("handoff"):
                    linear = "augmented"
                x, y, s = newton_iteration(A, b, c, x, y, s, linear=linear, dense=dense, tau=tau, info=info,
                                           handoff_floor=floor,
                                           refine_thresh=refine_thresh, forced=forced, refine_rounds=refine_rounds,
                                           refine_abs=None if refine_abs_kappa is None
                                           else refine_abs_kappa * tol * (1.0 + float(np.linalg.norm(b))))
                k += 1
    obj = objective(x, c) - float(cTlb)
    if not np.isfinite(obj) or not np.all(np.isfinite(x)):
        status = 2
    elif k >= max_iter:
        status = 1
    else:
        status = 0
    return dict(x=x, y=y, s=s, k=k, obj=float(obj), status=status, pivots_fixed=info.get("pivots_fixed", 0),
                refinements=info.get("refinements", 0),
                dependent_rows=int(forced.sum()) if forced is not None else 0,
                handoff=bool(info.get("handoff", False)), handoff_at=info.get("handoff_at"))


# --------------------------------------------------------------------------- synthetic workloads
def synthetic_dense_lp(m, n, seed):
    """Strictly primal-dual feasible dense LP (SURVEY.md §8d generator); returns A (m,n), b (m,), c (n,)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    xh = rng.uniform(0.1, 1.1, n)
    sh = rng.uniform(0.1, 1.1, n)
    yh = rng.standard_normal(m)
    b = A @ xh
    c = A.T @ yh + sh
    return A, b, c
