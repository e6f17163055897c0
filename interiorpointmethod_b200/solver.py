"""Host-side mirror of the reference's call surface for the Newton-step hot path.

The reference (payakorn/InteriorPointMethod) has no plugin interface; its seams are plain functions in
`main.py`.  This module offers the same names with the same argument meaning, routed to the B200 library
through ctypes (include/ipm_b200.h):

    interior_sparse(A, b, c, cTlb, tol)            main.py:760   -> objective - cTlb (float)
    interior(A, b, c, tol)                         main.py:707   -> Result (the reference prints and returns None)
    direction_predicted_sparse(..., method="gpu")  main.py:197
    direction_corrected_sparse(..., method="gpu")  main.py:247
    check_optimality / predicted_stepsize / duality_gap / full_stepsize / corrected
                                                   main.py:162, 305, 588, 604, 663
    solve(A, b, c, tol) -> Result(x, objective, iterations, ...)   (BASELINE.json north_star surface)

Vectors follow the reference's layout: (k,1) float64 columns in, (k,1) columns out.  Inputs are accepted
exactly as `create_problem_from_mps` produces them (csc_matrix with integer data, uint8/uint16/int16 b and c;
sparse_interior.py:211-216) and cast to float64 / int32 here.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib

try:  # scipy is only needed for sparse inputs
    from scipy import sparse as _sp
except Exception:  # pragma: no cover
    _sp = None

ETA = 0.91  # main.py:607


def _f64(v, shape=None):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
    if shape is not None and a.size != shape:
        raise ValueError("expected %d entries, got %d" % (shape, a.size))
    return a


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


@dataclass
class Result:
    x: np.ndarray            # (n,1)
    y: np.ndarray            # (m,1)
    s: np.ndarray            # (n,1)
    objective: float         # c^T x - cTlb
    iterations: int
    status: str              # converged | max_iter | nan
    residuals: dict = field(default_factory=dict)   # |rb|, |rc|, gap, |b|, |c| at exit


class NewtonStep:
    """One LP resident on one B200: owns an `ipm_handle`.  Not thread-safe (one stream)."""

    def __init__(self, A, b, c, device: int = 0):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.ipm_create(ctypes.byref(self._h), int(device)), None, "ipm_create")
        self.device = int(device)
        self._keep = None
        self.load(A, b, c)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ipm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        _lib.check(rc, self._h, what)

    # ------------------------------------------------------------------ data
    def load(self, A, b, c):
        if _sp is not None and _sp.issparse(A):
            # create_problem_from_mps hands over the csc_matrix loadmat produced (sparse_interior.py:211-216): it goes
            # to the device as it is (ipm_load_csc); any other sparse format goes row-compressed (ipm_load_csr).
            # The other orientation and the SpGEMM pattern are built on the GPU and cached per structure.
            csc = A.format == "csc"
            Ar = A.astype(np.float64) if csc else _sp.csr_matrix(A, dtype=np.float64)   # both copy
            Ar.sum_duplicates()
            Ar.sort_indices()
            self.m, self.n = Ar.shape
            bb, cc = _f64(b, self.m), _f64(c, self.n)
            ptr = np.ascontiguousarray(Ar.indptr, dtype=np.int32)
            idx = np.ascontiguousarray(Ar.indices, dtype=np.int32)
            val = np.ascontiguousarray(Ar.data, dtype=np.float64)
            fn = self._lib.ipm_load_csc if csc else self._lib.ipm_load_csr
            self._check(fn(self._h, self.m, self.n, int(Ar.nnz), _ptr(ptr), _ptr(idx), _ptr(val), _ptr(bb), _ptr(cc)),
                        "ipm_load_csc" if csc else "ipm_load_csr")
            self.sparse = True
            self.nnz = int(Ar.nnz)
        else:
            Ad = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
            if Ad.ndim != 2:
                raise ValueError("A must be 2-D")
            self.m, self.n = Ad.shape
            bb, cc = _f64(b, self.m), _f64(c, self.n)
            self._check(self._lib.ipm_load_dense(self._h, self.m, self.n, _ptr(Ad), self.n, _ptr(bb), _ptr(cc)),
                        "ipm_load_dense")
            self.sparse = False

    def pattern_info(self):
        """Ingestion record of the loaded sparse problem (ipm_pattern_info)."""
        out = np.zeros(6, dtype=np.int64)
        self._check(self._lib.ipm_pattern_info(self._h, _ptr(out)), "ipm_pattern_info")
        return dict(entries=int(out[0]), terms=int(out[1]), cache_hit=bool(out[2]), device_built=bool(out[3]),
                    build_ms=out[4] / 1e3, load_ms=out[5] / 1e3)

    def pattern(self):
        """Copies of the device-resident structure (ipm_get_pattern, ipm_get_values): parity tests."""
        info = self.pattern_info()
        nnz, ne, nt = self.nnz, info["entries"], info["terms"]
        out = dict(rowptr=np.empty(self.m + 1, np.int32), colind=np.empty(nnz, np.int32),
                   t_rowptr=np.empty(self.n + 1, np.int32), t_colind=np.empty(nnz, np.int32),
                   out_idx=np.empty(ne, np.int64), prod_ptr=np.empty(ne + 1, np.int64),
                   pa=np.empty(nt, np.int32), pb=np.empty(nt, np.int32))
        self._check(self._lib.ipm_get_pattern(self._h, *[_ptr(out[k]) for k in (
            "rowptr", "colind", "t_rowptr", "t_colind", "out_idx", "prod_ptr", "pa", "pb")]), "ipm_get_pattern")
        out["val"] = np.empty(nnz)
        out["t_val"] = np.empty(nnz)
        self._check(self._lib.ipm_get_values(self._h, _ptr(out["val"]), _ptr(out["t_val"])), "ipm_get_values")
        return out

    def set_pivot_threshold(self, tau: float):
        self._check(self._lib.ipm_set_pivot_threshold(self._h, float(tau)), "ipm_set_pivot_threshold")

    # ------------------------------------------------------------------ iterate
    def init_state(self, y0_is_one: bool = True):
        self._check(self._lib.ipm_init_state(self._h, int(bool(y0_is_one))), "ipm_init_state")

    def set_state(self, x, y, s):
        x, y, s = _f64(x, self.n), _f64(y, self.m), _f64(s, self.n)
        self._check(self._lib.ipm_set_state(self._h, _ptr(x), _ptr(y), _ptr(s)), "ipm_set_state")

    def get_state(self):
        x, y, s = np.empty(self.n), np.empty(self.m), np.empty(self.n)
        self._check(self._lib.ipm_get_state(self._h, _ptr(x), _ptr(y), _ptr(s)), "ipm_get_state")
        return _col(x), _col(y), _col(s)

    # ------------------------------------------------------------------ op level
    def residual_norms(self):
        out = np.empty(5)
        self._check(self._lib.ipm_residual_norms(self._h, _ptr(out)), "ipm_residual_norms")
        return dict(rb=out[0], rc=out[1], gap=out[2], b=out[3], c=out[4])

    def residuals(self):
        rb, rc = np.empty(self.m), np.empty(self.n)
        self._check(self._lib.ipm_get_residuals(self._h, _ptr(rb), _ptr(rc)), "ipm_get_residuals")
        return _col(rb), _col(rc)

    def assemble_normal(self):
        self._check(self._lib.ipm_assemble_normal(self._h), "ipm_assemble_normal")

    def get_M(self):
        M = np.empty((self.m, self.m))
        self._check(self._lib.ipm_get_M(self._h, _ptr(M)), "ipm_get_M")
        return M

    def factor(self, tau: float = 1e-30) -> int:
        nf = ctypes.c_int(0)
        self._check(self._lib.ipm_factor(self._h, float(tau), ctypes.byref(nf)), "ipm_factor")
        return nf.value

    def direction(self, kind: int, fetch: bool = True):
        if not fetch:
            self._check(self._lib.ipm_direction(self._h, int(kind), None, None, None), "ipm_direction")
            return None
        dx, dy, ds = np.empty(self.n), np.empty(self.m), np.empty(self.n)
        self._check(self._lib.ipm_direction(self._h, int(kind), _ptr(dx), _ptr(dy), _ptr(ds)), "ipm_direction")
        return _col(dx), _col(dy), _col(ds)

    def ratio_test(self, kind: int, eta: float = ETA):
        a = np.empty(2)
        self._check(self._lib.ipm_ratio_test(self._h, int(kind), float(eta), _ptr(a)), "ipm_ratio_test")
        return float(a[0]), float(a[1])

    def sigma(self):
        out = np.empty(3)
        self._check(self._lib.ipm_sigma(self._h, _ptr(out)), "ipm_sigma")
        return float(out[0]), float(out[1]), float(out[2])

    def update(self, alpha_p: float, alpha_d: float):
        self._check(self._lib.ipm_update(self._h, float(alpha_p), float(alpha_d)), "ipm_update")

    # ------------------------------------------------------------------ solve level
    def detect_dependent_rows(self, rel_tol: float = 1e-10) -> int:
        """Opt-in treatment of a rank-deficient A that is NOT in the reference (include/ipm_b200.h,
        ipm_detect_dependent_rows); returns the number of dependent rows found.  rel_tol <= 0 clears the mask."""
        nd = ctypes.c_int(0)
        self._check(self._lib.ipm_detect_dependent_rows(self._h, float(rel_tol), ctypes.byref(nd)),
                    "ipm_detect_dependent_rows")
        return nd.value

    def set_refinement(self, thresh: float | None):
        """Opt-in conditional refinement of the corrector (include/ipm_b200.h, ipm_set_refinement); None = off."""
        self._check(self._lib.ipm_set_refinement(self._h, -1.0 if thresh is None else float(thresh)), "ipm_set_refinement")

    def start_mehrotra(self):
        """Opt-in starting point that is NOT in the reference (include/ipm_b200.h, ipm_start_mehrotra)."""
        self._check(self._lib.ipm_start_mehrotra(self._h), "ipm_start_mehrotra")

    def solve(self, tol: float = 1e-8, max_iter: int = 5000, y0_is_one: bool = True, cTlb: float = 0.0,
              start: str = "reference") -> Result:
        """start="reference": x = s = 1 and y per `y0_is_one` (the reference's two drivers);
        "mehrotra": ipm_start_mehrotra (not in the reference, no iteration parity); "keep": the current iterate."""
        mode = {"reference": int(bool(y0_is_one)), "keep": 2, "mehrotra": 3}[start]
        x, y, s = np.empty(self.n), np.empty(self.m), np.empty(self.n)
        obj = ctypes.c_double(0.0)
        it = ctypes.c_int(0)
        st = ctypes.c_int(0)
        res = np.empty(5)
        self._check(self._lib.ipm_solve(self._h, float(tol), int(max_iter), mode, _ptr(x), _ptr(y),
                                        _ptr(s), ctypes.byref(obj), ctypes.byref(it), ctypes.byref(st), _ptr(res)),
                    "ipm_solve")
        return Result(x=_col(x), y=_col(y), s=_col(s), objective=float(obj.value) - float(cTlb),
                      iterations=it.value, status=_lib.STATUS.get(st.value, str(st.value)),
                      residuals=dict(rb=res[0], rc=res[1], gap=res[2], b=res[3], c=res[4]))


# ====================================================================== reference-shaped functions
def solve(A, b, c, tol: float = 1e-8, cTlb: float = 0.0, device: int = 0, max_iter: int = 5000,
          y0_is_one: bool | None = None, start: str = "reference", dependent_rows: float | None = None,
          refine: float | None = None) -> Result:
    """Load `benchmarks/*.mat`-style data, solve min c^T x s.t. Ax=b, x>=0; returns x, objective, iterations.
    start="mehrotra" selects the opt-in starting point that is not in the reference (NewtonStep.solve);
    dependent_rows=1e-10 the opt-in elimination of linearly dependent rows (NewtonStep.detect_dependent_rows)."""
    is_sparse = _sp is not None and _sp.issparse(A)
    if y0_is_one is None:
        y0_is_one = is_sparse          # sparse driver starts y=1 (sparse_interior.py:193-200), dense y=0 (main.py:287-302)
    with NewtonStep(A, b, c, device=device) as ns:
        if dependent_rows is not None:
            ns.detect_dependent_rows(dependent_rows)
        if refine is not None:
            ns.set_refinement(refine)
        return ns.solve(tol=tol, max_iter=max_iter, y0_is_one=y0_is_one, cTlb=cTlb, start=start)


def interior_sparse(A, b, c, cTlb, tol: float = 1e-20, device: int = 0, return_result: bool = False):
    """Drop-in for main.interior_sparse (main.py:760-815): returns `c^T x - cTlb` as a float, NaN on breakdown.

    Same defaults as the reference, including tol=1e-20 ("never converged") and the
    5000-iteration cap."""
    cTlb = float(np.asarray(cTlb).ravel()[0]) if np.size(cTlb) else 0.0
    res = solve(A, b, c, tol=tol, cTlb=cTlb, device=device, max_iter=5000, y0_is_one=True)
    print("k:\n", res.iterations)        # main.py:814
    return res if return_result else res.objective


def interior(A, b, c, tol: float = 1e-20, device: int = 0, verbose: bool = True) -> Result:
    """Dense driver (main.py:707-757): A list/ndarray, b and c 1-D, start y=0, cap 50000.
    Prints what the reference prints at the end (x, k, objective: main.py:754-757; not its per-iteration lines -
    the loop runs on the device).  The reference returns None; the Result is returned as well so that callers
    can use the numbers instead of parsing stdout."""
    res = solve(np.asarray(A, dtype=np.float64), b, c, tol=tol, device=device, max_iter=50000, y0_is_one=False)
    if verbose:
        print("optimal:", res.status == "converged")
        print("x:\n", res.x)
        print("k:\n", res.iterations)
        print("objective function:", res.objective)
    return res


def interior_kkt(A, b, c, tol: float = 1e-20, device: int = 0, max_iter: int = 50000) -> Result:
    """The reference's dense route itself on the GPU (`interior`, main.py:707-757 with `create_matrix` main.py:13-21 and
    np.linalg.solve main.py:178): predictor-corrector on the augmented system (the unreduced KKT matrix with ds
    eliminated exactly), LU with partial pivoting in one CTA (ipm_solve_dense_kkt).  Same Newton system and pivoting
    rule as the reference, hence the same iteration counts; it is the batched solver's hand-off target for LPs whose
    normal equations break down, and is exposed for parity tests.  n + m <= 1600."""
    lib = _lib.load()
    Ad = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
    m, n = Ad.shape
    bb, cc = _f64(b, m), _f64(c, n)
    x, y, s = np.empty(n), np.empty(m), np.empty(n)
    obj, it, st = ctypes.c_double(0.0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(lib.ipm_solve_dense_kkt(int(device), m, n, _ptr(Ad), _ptr(bb), _ptr(cc), float(tol), int(max_iter), _ptr(x),
                                       _ptr(y), _ptr(s), ctypes.byref(obj), ctypes.byref(it), ctypes.byref(st)), None,
               "ipm_solve_dense_kkt")
    return Result(x=_col(x), y=_col(y), s=_col(s), objective=float(obj.value), iterations=it.value,
                  status=_lib.STATUS.get(st.value, str(st.value)), residuals={})


_cache = {}


def _fingerprint(A, b, c):
    """Content check of the cached problem: the values can change in place while the objects stay the same."""
    import zlib
    vals = A.data if (_sp is not None and _sp.issparse(A)) else np.asarray(A)
    h = zlib.crc32(np.ascontiguousarray(vals).view(np.uint8).reshape(-1))
    h = zlib.crc32(np.ascontiguousarray(np.asarray(b)).view(np.uint8).reshape(-1), h)
    h = zlib.crc32(np.ascontiguousarray(np.asarray(c)).view(np.uint8).reshape(-1), h)
    return (np.shape(A), h)


def _cached_step(A, b, c, device=0) -> NewtonStep:
    """The op-level functions below take (A, b, c) on every call like the reference's do; the problem stays
    resident on the GPU between calls as long as the caller passes the SAME objects with the SAME contents: the
    cache entry holds strong references (so ids cannot be recycled by the garbage collector) and is compared by
    identity and by a CRC of the values (so an in-place edit of b, c or A.data re-uploads)."""
    ent = _cache.get("entry")
    fp = _fingerprint(A, b, c)
    if ent is not None and ent["A"] is A and ent["b"] is b and ent["c"] is c and ent["device"] == device \
            and ent["fp"] == fp:
        return ent["ns"]
    if ent is not None:
        ent["ns"].close()
        _cache.pop("entry", None)
    ns = NewtonStep(A, b, c, device=device)
    _cache["entry"] = dict(A=A, b=b, c=c, device=device, fp=fp, ns=ns)
    return ns


def release_cached_step():
    """Drop the problem the op-level functions keep resident (and the references to the caller's arrays)."""
    ent = _cache.pop("entry", None)
    if ent is not None:
        ent["ns"].close()


def check_optimality(A, b, c, x, y, s, e1, e2, e3, options="sparse", device: int = 0) -> bool:
    """main.py:162-173: True = not optimal yet."""
    ns = _cached_step(A, b, c, device)
    ns.set_state(x, y, s)
    r = ns.residual_norms()
    return bool(e1 * (1 + r["b"]) < r["rb"] or e2 * (1 + r["c"]) < r["rc"] or e3 < r["gap"])


def _prepare(ns, x, y, s, tau):
    ns.set_state(x, y, s)
    ns.residual_norms()
    ns.assemble_normal()
    ns.factor(tau)


def direction_predicted_sparse(A, b, c, x, y, s, method="gpu", device: int = 0, tau: float = 1e-30):
    """main.py:197-229 with the linear algebra on the GPU (`method` kept for signature compatibility)."""
    ns = _cached_step(A, b, c, device)
    _prepare(ns, x, y, s, tau)
    return ns.direction(0)


def direction_corrected_sparse(A, b, c, x, y, s, delta_x_aff, delta_y_aff, delta_s_aff, method="gpu",
                               device: int = 0, tau: float = 1e-30):
    """main.py:247-277.  The predictor direction is recomputed on the device from (x, y, s): it is a pure
    function of the iterate, so the caller's delta_*_aff are only checked for shape."""
    ns = _cached_step(A, b, c, device)
    _prepare(ns, x, y, s, tau)
    ns.direction(0, fetch=False)
    ns.sigma()
    return ns.direction(1)


def predicted_stepsize(delta_x_aff, delta_y_aff, delta_s_aff, x, s, device: int = 0):
    """main.py:305-322: (alpha_primal, alpha_dual) = min({-v_i/dv_i : dv_i < 0} U {1}) for (x, dx) and (s, ds)."""
    return _ratio(x, delta_x_aff, s, delta_s_aff, 0.0, device)


def full_stepsize(x, y, s, delta_x, delta_y, delta_s, delta_x_aff=None, delta_y_aff=None, delta_s_aff=None,
                  device: int = 0):
    """main.py:604-626: min(1, 0.91 * ratio) for both step lengths (always <= 0.91)."""
    return _ratio(x, delta_x, s, delta_s, ETA, device)


def step_size(x, y, s, delta_aff=None, delta=None, lb=None, ub=None, device: int = 0):
    """main.py:325-547: the ratio test with simple bounds kept implicit, `lb <= x <= ub` with either side optional
    (None: only x >= 0 on the lower side, no upper side).  `delta_aff=(dx, dy, ds)`: predictor step lengths;
    `delta=(dx, dy, ds)`: corrector step lengths min(0.91 * ratio, 1).  Returns (alpha_primal, alpha_dual), or None when
    neither direction is given, like the reference.  The reference's behaviour is kept case by case, including
    alpha_dual = 1 for the corrector without bounds (main.py:449-454)."""
    if delta_aff is not None:
        d, eta = delta_aff, 0.0
    elif delta is not None:
        d, eta = delta, ETA
    else:
        return None
    lib = _lib.load()
    xv, dxv, sv, dsv = _f64(x), _f64(d[0]), _f64(s), _f64(d[2])
    lbv = None if lb is None else _f64(lb)
    ubv = None if ub is None else _f64(ub)
    if xv.size != dxv.size or sv.size != dsv.size or xv.size != sv.size or \
            (lbv is not None and lbv.size != xv.size) or (ubv is not None and ubv.size != xv.size):
        raise ValueError("step_size: x, s, the direction and the bounds must have the same length")
    out = np.empty(2)
    _lib.check(lib.ipm_op_step_size_bounded(int(device), xv.size, _ptr(xv), _ptr(dxv), _ptr(sv), _ptr(dsv),
                                            None if lbv is None else _ptr(lbv), None if ubv is None else _ptr(ubv),
                                            float(eta), _ptr(out)), None, "ipm_op_step_size_bounded")
    return float(out[0]), float(out[1])


def predicted_stepsize_lb_ub(delta_x_aff, delta_y_aff, delta_s_aff, x, s, lb, ub, device: int = 0):
    """main.py:550-559."""
    return step_size(x=x, y=None, s=s, delta_aff=(delta_x_aff, delta_y_aff, delta_s_aff), lb=lb, ub=ub, device=device)


def full_stepsize_lb_ub(x, y, s, delta_x, delta_y, delta_s, delta_x_aff, delta_y_aff, delta_s_aff, lb, ub,
                        device: int = 0):
    """main.py:629-660."""
    return step_size(x, y, s, delta=(delta_x, delta_y, delta_s), lb=lb, ub=ub, device=device)


def _ratio(x, dx, s, ds, eta, device):
    lib = _lib.load()
    xv, dxv, sv, dsv = _f64(x), _f64(dx), _f64(s), _f64(ds)
    out = np.empty(2)
    _lib.check(lib.ipm_op_ratio_test(int(device), xv.size, _ptr(xv), _ptr(dxv), _ptr(sv), _ptr(dsv), float(eta),
                                     _ptr(out)), None, "ipm_op_ratio_test")
    return float(out[0]), float(out[1])


def duality_gap(A, x, y, s, delta_x_aff, delta_y_aff, delta_s_aff, device: int = 0):
    """main.py:588-601: (mu_aff, mu_k, centering) with centering = (mu_aff/mu_k)**3, unclamped."""
    lib = _lib.load()
    xv, sv, dxv, dsv = _f64(x), _f64(s), _f64(delta_x_aff), _f64(delta_s_aff)
    out = np.empty(3)
    _lib.check(lib.ipm_op_sigma(int(device), xv.size, _ptr(xv), _ptr(sv), _ptr(dxv), _ptr(dsv), _ptr(out)), None,
               "ipm_op_sigma")
    return float(out[0]), float(out[1]), float(out[2])


def corrected(x, y, s, delta_x, delta_y, delta_s, delta_x_aff=None, delta_y_aff=None, delta_s_aff=None,
              device: int = 0):
    """main.py:663-697: full_stepsize then x + ap*dx, y + ad*dy, s + ad*ds; returns new (x, y, s) columns."""
    lib = _lib.load()
    ap, ad = full_stepsize(x, y, s, delta_x, delta_y, delta_s, device=device)
    xv, yv, sv = _f64(x).copy(), _f64(y).copy(), _f64(s).copy()
    dxv, dyv, dsv = _f64(delta_x), _f64(delta_y), _f64(delta_s)
    _lib.check(lib.ipm_op_update(int(device), yv.size, xv.size, _ptr(xv), _ptr(yv), _ptr(sv), _ptr(dxv), _ptr(dyv),
                                 _ptr(dsv), ap, ad), None, "ipm_op_update")
    return _col(xv), _col(yv), _col(sv)


def solve_linear(A, b, method="gpu", device: int = 0, tau: float = 1e-30):
    """main.py:176-182 for the normal-equations matrix of main.py:226 (symmetric positive semidefinite A):
    safeguarded Cholesky + triangular sweeps on the GPU; returns an (N,1) array like the reference."""
    lib = _lib.load()
    if _sp is not None and _sp.issparse(A):
        A = A.toarray()
    Ad = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
    rhs = _f64(b, Ad.shape[0])
    z = np.empty(Ad.shape[0])
    nf = ctypes.c_int(0)
    _lib.check(lib.ipm_solve_spd(int(device), Ad.shape[0], _ptr(Ad), _ptr(rhs), float(tau), _ptr(z), ctypes.byref(nf)),
               None, "ipm_solve_spd")
    return _col(z)


def newton_iteration(ns: NewtonStep, tau: float = 1e-30):
    """One predictor-corrector iteration through the op-level entry points (main.py:781-805); returns the
    per-op results so parity tests can compare each against the oracle."""
    norms = ns.residual_norms()
    ns.assemble_normal()
    nfixed = ns.factor(tau)
    dxa, dya, dsa = ns.direction(0)
    alpha_aff = ns.ratio_test(0)
    mu_aff, mu, sigma = ns.sigma()
    dx, dy, ds = ns.direction(1)
    alpha = ns.ratio_test(1, ETA)
    ns.update(*alpha)
    return dict(norms=norms, nfixed=nfixed, dx_aff=dxa, dy_aff=dya, ds_aff=dsa, alpha_aff=alpha_aff,
                mu_aff=mu_aff, mu=mu, sigma=sigma, dx=dx, dy=dy, ds=ds, alpha=alpha)
