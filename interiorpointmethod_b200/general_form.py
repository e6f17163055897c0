"""General-form front end: the callers' side of the Newton-step path.

The reference's `new_interior_sparse` (main.py:1081-1246) takes a linprog-style problem

    min f^T x   s.t.  Aineq x <= bineq,  Aeq x = beq,  lb <= x <= ub

as stored in `benchmarks_full/*.mat` (loader sparse_interior.py:294-315), turns it into the standard form
`min c^T x  s.t.  A x = b, x >= 0` with `get_Abc(options="no-bound")` (main.py:818-965) and folds the bounds into
the matrix with `add_bound_into_matrix` (main.py:968-1060).  That last function is unfinished in the reference:
the branch with non-zero lower AND finite upper bounds returns the matrix unchanged ("FIXME complete this!!",
main.py:1047-1060) and a lower bound of -inf raises.  This module mirrors the three entry points with the same
argument meaning and return shapes, agrees with the reference wherever the reference's code is complete (tests
compare against frozen outputs of the unmodified reference), and finishes the other branches:

    lower bounds      x = lb + x'            constant  f^T lb,  b -= A lb            (main.py:1041-1045)
    upper bounds      rows [U I] [x'; w] = ub - lb  for every finite ub             (main.py:1013-1039)
    free variables    x_j = x_j^+ - x_j^-   (the reference raises "there are -inf in lower bound")

The standard-form LP then goes through the same C ABI as every other solve (`solver.solve`).
"""
from __future__ import annotations

import os

import numpy as np
from scipy import sparse


# ------------------------------------------------------------------------------------------ loading
def load_data_mps_matlab(name, root="benchmarks_full"):
    """(c, Aineq, bineq, Aeq, beq, lb, ub) of `<root>/<name>.mat`, fields of the MATLAB struct `data`
    (sparse_interior.py:305-315)."""
    from scipy.io import loadmat

    path = os.path.join(root, name if name.endswith(".mat") else name + ".mat")
    d = loadmat(path)["data"]

    def get(key):
        return d[key][0][0]

    return get("f"), get("Aineq"), get("bineq"), get("Aeq"), get("beq"), get("lb"), get("ub")


def create_problem_from_mps_matlab(name, root="benchmarks_full"):
    """Same as the reference (sparse_interior.py:294-303): empty constraint blocks become None."""
    c, Aineq, bineq, Aeq, beq, lb, ub = load_data_mps_matlab(name, root)
    if np.size(bineq) == 0:
        Aineq = bineq = None
    if np.size(beq) == 0:
        Aeq = beq = None
    return c, Aineq, bineq, Aeq, beq, lb, ub


def load_golden_general(name):
    """Frozen copy of the reference loader's output for `benchmarks_full/<name>.mat` (tests/golden/full/)."""
    from .problems import _REPO

    z = np.load(os.path.join(_REPO, "tests", "golden", "full", name + ".npz"))

    def mat(prefix):
        shape = tuple(int(v) for v in z[prefix + "_shape"])
        if shape[0] == 0:
            return None
        return sparse.csc_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=shape)

    Aineq, Aeq = mat("Aineq"), mat("Aeq")
    bineq = z["bineq"].reshape(-1, 1) if Aineq is not None else None
    beq = z["beq"].reshape(-1, 1) if Aeq is not None else None
    return z["f"].reshape(-1, 1), Aineq, bineq, Aeq, beq, z["lb"].reshape(-1, 1), z["ub"].reshape(-1, 1)


# ------------------------------------------------------------------------------------------ standard form
def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def get_Abc(c, Aeq=None, beq=None, Aineq=None, bineq=None, lb=None, ub=None, options="no-bound"):
    """Equality form with one slack column per inequality row: returns (A, b, c, bound) like main.py:895-965
    (`options="no-bound"`, the only mode `new_interior_sparse` uses).

        A = [[Aineq, I], [Aeq, 0]],  b = [bineq; beq],  c = [c; 0]
        bound = None when every lb is finite... and every ub infinite; else (lb or None, ub or None), with the
        upper bounds of the slack columns set to +inf.

    Differences from the reference, all outside its working range: a lower bound of -inf is kept in `bound`
    instead of raising (add_bound_into_matrix splits the variable), lower bounds are extended with zeros for the
    slack columns (the reference leaves `lb` short), and the Aineq-only dense branch returns four values (the
    reference forgets `c`, main.py:955)."""
    if options != "no-bound":
        raise ValueError("only options='no-bound' is mirrored (the mode new_interior_sparse uses, main.py:1092-1101)")
    c = _col(c)
    n = c.shape[0]
    lb = np.zeros((n, 1)) if lb is None else _col(lb)
    ub = np.full((n, 1), np.inf) if ub is None else _col(ub)
    blocks, rhs = [], []
    n_slack = 0
    if Aineq is not None:
        Aineq = sparse.csc_matrix(Aineq, dtype=np.float64)
        n_slack = Aineq.shape[0]
        blocks.append(sparse.hstack([Aineq, sparse.identity(n_slack, format="csc")], format="csc"))
        rhs.append(_col(bineq))
    if Aeq is not None:
        Aeq = sparse.csc_matrix(Aeq, dtype=np.float64)
        if n_slack:
            Aeq = sparse.hstack([Aeq, sparse.csc_matrix((Aeq.shape[0], n_slack))], format="csc")
        blocks.append(Aeq)
        rhs.append(_col(beq))
    if not blocks:
        raise ValueError("no constraints")
    A = sparse.vstack(blocks, format="csc") if len(blocks) > 1 else blocks[0]
    b = np.vstack(rhs)
    if n_slack:
        c = np.vstack([c, np.zeros((n_slack, 1))])
        lb = np.vstack([lb, np.zeros((n_slack, 1))])
        ub = np.vstack([ub, np.full((n_slack, 1), np.inf)])
    lb_out = None if np.count_nonzero(lb) == 0 else lb       # all-zero lower bounds: already x >= 0
    ub_out = None if np.isinf(ub).all() else ub
    bound = None if (lb_out is None and ub_out is None) else (lb_out, ub_out)
    return A, b, c, bound


def add_bound_into_matrix(A, b, c, bound):
    """Folds `bound = (lb, ub)` into the equality system: returns (A, b, c, (None, None), constant) like
    main.py:968-1060, where `constant` has to be ADDED to c^T x' to obtain the objective in the caller's variables.

    The reference's convention for the constant (main.py:1043: `constant = -c.T @ lb`, then `b + A @ lb`) has
    both signs flipped with respect to the substitution x = lb + x' it documents; here constant = +c^T lb and
    b' = b - A lb, which is what makes the optimum of the shifted LP agree with the Netlib optimum.
    Column order of the result: [x' (or x^+) | x^- of the free variables | slacks of the upper-bound rows]."""
    lb, ub = bound if bound is not None else (None, None)
    A = sparse.csc_matrix(A, dtype=np.float64)
    b = _col(b).copy()
    c = _col(c).copy()
    m, n = A.shape
    if lb is None and ub is None:
        return A, b, c, (None, None), 0.0
    lb = np.zeros((n, 1)) if lb is None else _col(lb).copy()
    ub = np.full((n, 1), np.inf) if ub is None else _col(ub).copy()
    constant = 0.0
    # ---- free variables: x_j = x_j^+ - x_j^-  (with a finite upper bound: x_j = ub_j - x_j', x_j' >= 0)
    free = np.nonzero(np.isinf(lb.ravel()) & (lb.ravel() < 0))[0]
    flip = [j for j in free if np.isfinite(ub[j, 0])]
    split = np.array([j for j in free if not np.isfinite(ub[j, 0])], dtype=np.int64)
    if len(flip):
        # x_j = ub_j - x_j'  ->  column negated, constant and b shifted, bounds 0 <= x_j' < inf
        flip = np.asarray(flip, dtype=np.int64)
        u = np.zeros((n, 1))
        u[flip, 0] = ub[flip, 0]
        constant += float((c.T @ u)[0, 0])
        b -= A @ u
        sign = np.ones(n)
        sign[flip] = -1.0
        A = (A @ sparse.diags(sign)).tocsc()
        c[flip] *= -1.0
        lb[flip] = 0.0
        ub[flip] = np.inf
    if len(split):
        A = sparse.hstack([A, -A[:, split]], format="csc")
        c = np.vstack([c, -c[split]])
        lb[split] = 0.0
        lb = np.vstack([lb, np.zeros((len(split), 1))])
        ub = np.vstack([ub, np.full((len(split), 1), np.inf)])
        n = A.shape[1]
    # ---- lower bounds: x = lb + x'
    if np.count_nonzero(lb):
        constant += float((c.T @ lb)[0, 0])
        b -= A @ lb
        ub = ub - lb                     # inf stays inf
    # ---- finite upper bounds: [U I] [x'; w] = ub
    cols = np.nonzero(np.isfinite(ub.ravel()))[0]
    k = len(cols)
    if k:
        U = sparse.csc_matrix((np.ones(k), (np.arange(k), cols)), shape=(k, n))
        top = sparse.hstack([A, sparse.csc_matrix((m, k))], format="csc")
        bottom = sparse.hstack([U, sparse.identity(k, format="csc")], format="csc")
        A = sparse.vstack([top, bottom], format="csc")
        b = np.vstack([b, ub[cols]])
        c = np.vstack([c, np.zeros((k, 1))])
    return A, b, c, (None, None), constant


def standard_form(c, Aeq=None, beq=None, Aineq=None, bineq=None, lb=None, ub=None):
    """get_Abc + add_bound_into_matrix: (A, b, c_std, constant, n_original, recover) where
    recover(x_std) returns x in the caller's variables."""
    c0 = _col(c)
    n0 = c0.shape[0]
    lb0 = np.zeros((n0, 1)) if lb is None else _col(lb)
    ub0 = np.full((n0, 1), np.inf) if ub is None else _col(ub)
    A, b, cs, bound = get_Abc(c0, Aeq=Aeq, beq=beq, Aineq=Aineq, bineq=bineq, lb=lb0, ub=ub0)
    n1 = A.shape[1]
    if bound is not None:
        A, b, cs, _, constant = add_bound_into_matrix(A, b, cs, bound)
    else:
        constant = 0.0
    free = np.isinf(lb0.ravel()) & (lb0.ravel() < 0)
    flip = free & np.isfinite(ub0.ravel())
    split = np.nonzero(free & ~np.isfinite(ub0.ravel()))[0]

    def recover(x_std):
        xs = np.asarray(x_std, dtype=np.float64).ravel()
        x = xs[:n0].copy()
        if len(split):
            x[split] -= xs[n1:n1 + len(split)]
        x[flip] = ub0.ravel()[flip] - x[flip]
        shift = np.where(free, 0.0, lb0.ravel())
        return (x + shift).reshape(-1, 1)

    return A, b, cs, constant, n0, recover


def new_interior_sparse(c, Aeq=None, beq=None, Aineq=None, bineq=None, lb=None, ub=None, tol=1e-8, device=0,
                        start="mehrotra", max_iter=1000):
    """GPU counterpart of main.new_interior_sparse (main.py:1081-1246): standard form, then the Newton-step path
    through the C ABI.  Returns a `solver.Result` whose objective is f^T x in the caller's variables and whose
    `x` has the caller's length.  The reference starts from x = s = y = 1 with e3 = 1e-6 and caps at 1000
    iterations (main.py:1086-1088, 1124-1126); the default here is the opt-in Mehrotra start, `start="reference"`
    gives the reference's."""
    from . import solver

    A, b, cs, constant, n0, recover = standard_form(c, Aeq, beq, Aineq, bineq, lb, ub)
    res = solver.solve(A, b, cs, tol=tol, cTlb=-constant, device=device, max_iter=max_iter, y0_is_one=True, start=start)
    res.x_standard = res.x
    res.x = recover(res.x)
    return res
