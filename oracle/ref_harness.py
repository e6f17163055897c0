"""TEST INFRASTRUCTURE — not product code.

In-process harness around the UNMODIFIED reference at /root/reference (build
container only; the GPU box has no /root/reference).  It is used by
`oracle/make_golden.py` to freeze golden vectors under `tests/golden/` and by the
CPU tests that validate `oracle/ipm_oracle.py` against the real reference when the
reference tree is present.

Nothing in the product package (`interiorpointmethod_b200/`) may import this module.

The reference's drivers do not return x or k (main.py:815, main.py:754-757), so the
harness replays the loop bodies by calling the reference's OWN functions in the same
order (main.py:780-807 for `interior_sparse`, main.py:725-751 for `interior`).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

# The reference tree in the build container; on the GPU box (no /root/reference) the verbatim copy of its two
# solver modules that `__graft_entry__.build()` places in the git-ignored oracle/_ref/ (it travels like a built .so).
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "main.py")) and os.path.isfile(os.path.join(_HERE, "_ref", "main.py")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main.py"))


def reference_has_benchmarks() -> bool:
    """The .mat files only exist in the build container's reference tree (the loaders are cwd-relative)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "benchmarks"))


_ref = None


def load_reference():
    """Import the reference's `main` and `sparse_interior` without touching its tree.

    main.py:10 imports matplotlib (absent here, never used by solver code) -> stub it.
    main.py:2 imports scipy.optimize.linprog (present).  Loader paths are cwd-relative
    (sparse_interior.py:157) -> callers use `in_reference_cwd()`.
    """
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import main as ref_main  # noqa: E402
    import sparse_interior as ref_sparse  # noqa: E402

    _ref = (ref_main, ref_sparse)
    return _ref


@contextlib.contextmanager
def in_reference_cwd():
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_problem(name):
    """create_problem_from_mps(name) exactly as the reference does it (sparse_interior.py:211-216)."""
    _, ref_sparse = load_reference()
    with in_reference_cwd():
        return ref_sparse.create_problem_from_mps(name)


def replay_interior_sparse(A, b, c, cTlb, tol=1e-8, max_iter=5000, trace_at=()):
    """Replay of main.py:776-815 calling the reference's own functions.

    Returns dict(x, y, s, k, obj, trace) where trace[k] holds the state and the
    reference's per-op results at iteration k for every k in `trace_at`.
    """
    import numpy as np
    import warnings

    ref_main, ref_sparse = load_reference()
    m, n = np.shape(A)
    k = 0
    x, y, s = ref_sparse.initial_vector_sparse(m, n)
    trace = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        while ref_main.check_optimality(A, b, c, x, y, s, tol, tol, tol, options="sparse") and k < max_iter:
            dxa, dya, dsa = ref_main.direction_predicted_sparse(A, b, c, x, y, s)
            ap_aff, ad_aff = ref_main.predicted_stepsize(dxa, dya, dsa, x, s)
            mu_aff, mu_k, sigma = ref_main.duality_gap(A, x, y, s, dxa, dya, dsa)
            dx, dy, ds = ref_main.direction_corrected_sparse(A, b, c, x, y, s, dxa, dya, dsa)
            ap, ad = ref_main.full_stepsize(x, y, s, dx, dy, ds, dxa, dya, dsa)
            if k in trace_at:
                trace[k] = dict(
                    x=x.copy(), y=y.copy(), s=s.copy(),
                    dx_aff=dxa, dy_aff=dya, ds_aff=dsa,
                    alpha_aff=(float(ap_aff), float(ad_aff)),
                    mu_aff=float(mu_aff), mu=float(mu_k), sigma=float(sigma),
                    dx=dx, dy=dy, ds=ds, alpha=(float(ap), float(ad)),
                )
            x, y, s = ref_main.corrected(x, y, s, dx, dy, ds, dxa, dya, dsa)
            k += 1
    obj = (sum(x * c) - cTlb)[0]
    return dict(x=x, y=y, s=s, k=k, obj=float(obj), trace=trace)


def replay_interior_dense(A, b, c, tol=1e-8, max_iter=50000):
    """Replay of main.py:718-751 (`interior`, dense KKT + np.linalg.solve, y0 = 0)."""
    import numpy as np
    import warnings

    ref_main, _ = load_reference()
    A, b, c = ref_main.convert_to_array(A, b, c)
    k = 0
    x, y, s = ref_main.initial_vector(A)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        while ref_main.check_optimality(A, b, c, x, y, s, tol, tol, tol) and k < max_iter:
            dxa, dya, dsa = ref_main.direction_predicted(A, b, c, x, y, s)
            dx, dy, ds = ref_main.direction_corrected(A, b, c, x, y, s, dxa, dya, dsa)
            x, y, s = ref_main.corrected(x, y, s, dx, dy, ds, dxa, dya, dsa)
            k += 1
    obj = float(np.sum(x * c))
    return dict(x=x, y=y, s=s, k=k, obj=obj)


def call_interior_dense(A, b, c, tol=1e-8):
    """The reference's dense driver itself (main.py:707-757), output silenced; it returns None (prints x, k,
    objective), so this is for timing the unmodified code path - values come from replay_interior_dense."""
    import warnings

    ref_main, _ = load_reference()
    with quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_main.interior(A, b, c, tol=tol)


def call_interior_sparse(A, b, c, cTlb, tol=1e-8):
    """The reference driver itself (main.py:760), output silenced."""
    import warnings

    ref_main, _ = load_reference()
    with quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_main.interior_sparse(A=A, b=b, c=c, cTlb=cTlb, tol=tol)
