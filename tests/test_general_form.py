"""General-form front end: get_Abc / add_bound_into_matrix / new_interior_sparse mirrors
against frozen outputs of the unmodified reference (tests/golden/full/, written by oracle/make_golden_full.py),
against the Netlib optima listed in main.py:1417-1616, and against scipy's HiGHS on hand-made bounded LPs."""
import json
import os

import numpy as np
import pytest
from scipy import sparse

from interiorpointmethod_b200 import general_form as gf

FULL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full")
META = json.load(open(os.path.join(FULL, "standard_form.json")))
NAMES = sorted(META)
# Target objective: the optimum scipy's HiGHS finds on the frozen data (oracle/make_golden_full.py).  It agrees with
# the table in main.py:1417-1616 to 1e-7 on 45 of the 52 LPs; the table's entries for BNL1, FINNIS, FORPLAN, GANGES,
# SCAGR7 and SCRS8 differ from the optimum of the data in benchmarks_full by 1e-7..6e-6 relative (SCRS8 looks like a
# typo: 904.29998619 for 904.29695380) and STANDGUB is listed as "(see NOTES)".
OPT = {k: v["highs_optimum"] for k, v in META.items()}
TABLE_DISAGREES = {"BNL1", "FINNIS", "FORPLAN", "GANGES", "SCAGR7", "SCRS8", "STANDGUB"}
# degenerate LP on which the normal equations can stall at rb = 3e-6 with the objective correct to 1e-9 (same
# mechanism as QAP15): it did with the per-block substitution solves and converges with the
# inverse-block pipelined ones
STALLS_AT_OPTIMUM = {"DEGEN3"}


def test_netlib_table_of_the_reference_agrees_with_the_data():
    for name, v in META.items():
        if name in TABLE_DISAGREES:
            continue
        assert abs(v["netlib_optimum"] - v["highs_optimum"]) <= 1e-7 * max(1.0, abs(v["highs_optimum"])), name


def _checksum(A, b, c):
    A = sparse.csr_matrix(A, dtype=np.float64)
    m, n = A.shape
    u = np.cos(np.arange(n, dtype=np.float64))[:, None]
    v = np.sin(np.arange(m, dtype=np.float64))[:, None]
    return {"shape": [int(m), int(n)], "nnz": int(A.nnz), "A_u": float(np.abs(A @ u).sum()),
            "At_v": float(np.abs(A.T @ v).sum()), "b_sum": float(np.sum(b)), "b_abs": float(np.abs(b).sum()),
            "c_sum": float(np.sum(c)), "c_abs": float(np.abs(c).sum())}


def _same(got, ref):
    assert got["shape"] == ref["shape"] and got["nnz"] == ref["nnz"], (got, ref)
    for k in ("A_u", "At_v", "b_sum", "b_abs", "c_sum", "c_abs"):
        assert abs(got[k] - ref[k]) <= 1e-10 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


@pytest.mark.parametrize("name", NAMES)
def test_get_Abc_matches_reference(name):
    """[[Aineq, I], [Aeq, 0]], [bineq; beq], [c; 0] exactly as main.get_Abc(options="no-bound") builds them."""
    c, Aineq, bineq, Aeq, beq, lb, ub = gf.load_golden_general(name)
    A, b, cs, bound = gf.get_Abc(c, Aeq=Aeq, beq=beq, Aineq=Aineq, bineq=bineq, lb=lb, ub=ub)
    _same(_checksum(A, b, cs), META[name]["get_Abc"])
    ref_bound = META[name]["get_Abc"]["bound"]
    if META[name]["lb_nonzero"] == 0:
        # the reference drops finite lower bounds (`(lb > -inf).all() -> lb = None`, main.py:901-904): identical
        # classification only where they are all zero
        mine = "none" if bound is None else "lb=%s,ub=%s" % ("None" if bound[0] is None else "set",
                                                             "None" if bound[1] is None else "set")
        assert mine == ref_bound
    else:
        assert bound is not None and bound[0] is not None       # kept here, lost in the reference


@pytest.mark.parametrize("name", [n for n in NAMES if "add_bound" in META[n] and META[n]["lb_nonzero"] == 0])
def test_add_bound_into_matrix_matches_reference_upper_bound_branch(name):
    """[A 0; U I], [b; ub], [c; 0], constant 0 - the `lb is None` branch the reference implements (main.py:1013-1039)."""
    c, Aineq, bineq, Aeq, beq, lb, ub = gf.load_golden_general(name)
    A, b, cs, bound = gf.get_Abc(c, Aeq=Aeq, beq=beq, Aineq=Aineq, bineq=bineq, lb=lb, ub=ub)
    A2, b2, c2, bound2, const = gf.add_bound_into_matrix(A, b, cs, bound)
    ref = META[name]["add_bound"]
    _same(_checksum(A2, b2, c2), ref)
    assert bound2 == (None, None) and const == ref["constant"] == 0.0


def _highs(c, Aineq, bineq, Aeq, beq, lb, ub):
    from scipy.optimize import linprog
    bounds = [(None if np.isneginf(l) else float(l), None if np.isposinf(u) else float(u))
              for l, u in zip(np.ravel(lb), np.ravel(ub))]
    r = linprog(np.ravel(c), A_ub=Aineq, b_ub=None if bineq is None else np.ravel(bineq), A_eq=Aeq,
                b_eq=None if beq is None else np.ravel(beq), bounds=bounds, method="highs")
    assert r.status == 0
    return r


def test_standard_form_with_lower_upper_and_free_variables_against_highs():
    """Branches the reference leaves unfinished or rejects: lb != 0 together with finite ub, free variables with and
    without an upper bound.  The standard form is solved by the CPU oracle and mapped back."""
    from oracle import ipm_oracle as orc
    rng = np.random.default_rng(3)
    n, mi, me = 9, 5, 3
    Aineq = rng.standard_normal((mi, n)); Aeq = rng.standard_normal((me, n))
    x0 = rng.uniform(-1, 2, n)
    bineq = (Aineq @ x0 + rng.uniform(0.1, 1.0, mi)).reshape(-1, 1); beq = (Aeq @ x0).reshape(-1, 1)
    lb = np.array([-1.5, 0, 0.5, -np.inf, -np.inf, 0, -2, 0, -np.inf], float).reshape(-1, 1)
    ub = np.array([3, np.inf, 4, np.inf, 2.5, 1.5 + 2, np.inf, np.inf, np.inf], float).reshape(-1, 1)
    x0 = np.clip(x0, np.where(np.isinf(lb.ravel()), -5, lb.ravel()), np.where(np.isinf(ub.ravel()), 5, ub.ravel()))
    bineq = (Aineq @ x0 + 0.5).reshape(-1, 1); beq = (Aeq @ x0).reshape(-1, 1)
    c = rng.uniform(0.2, 1.0, n).reshape(-1, 1)
    c[[3, 8]] = 0.0 * c[[3, 8]] + np.array([[0.3], [-0.2]])
    # bounded below overall? add a box through an extra inequality to keep the LP bounded
    Aineq = np.vstack([Aineq, np.eye(n), -np.eye(n)]); bineq = np.vstack([bineq, np.full((n, 1), 6.0), np.full((n, 1), 6.0)])
    ref = _highs(c, Aineq, bineq, Aeq, beq, lb, ub)
    A, b, cs, const, n0, rec = gf.standard_form(c, Aeq=sparse.csc_matrix(Aeq), beq=beq, Aineq=sparse.csc_matrix(Aineq),
                                                bineq=bineq, lb=lb, ub=ub)
    r = orc.solve(A, b, cs, cTlb=-const, tol=1e-9, start="mehrotra", max_iter=200)
    assert r["status"] == 0
    x = rec(r["x"])
    assert abs(r["obj"] - ref.fun) <= 1e-7 * max(1.0, abs(ref.fun))
    assert abs(float((c.T @ x)[0, 0]) - ref.fun) <= 1e-7 * max(1.0, abs(ref.fun))
    assert np.all(x >= lb - 1e-7) and np.all(x <= ub + 1e-7)
    assert np.max(Aineq @ x - bineq) <= 1e-7 and np.max(np.abs(Aeq @ x - beq)) <= 1e-7


@pytest.mark.parametrize("name", ["AFIRO", "ADLITTLE", "KB2", "BOEING2", "SC50A", "BORE3D"])
def test_front_end_plus_oracle_reaches_netlib_optimum(name):
    """benchmarks_full carries the true bounded forms (benchmarks/KB2.mat has b = 0): with the
    bounds folded in, the Netlib optimum of main.py:1417-1616 is reproduced."""
    from oracle import ipm_oracle as orc
    c, Aineq, bineq, Aeq, beq, lb, ub = gf.load_golden_general(name)
    A, b, cs, const, n0, rec = gf.standard_form(c, Aeq, beq, Aineq, bineq, lb, ub)
    r = orc.solve(A, b, cs, cTlb=-const, tol=1e-8, start="mehrotra", max_iter=300)
    assert r["status"] == 0
    assert abs(r["obj"] - OPT[name]) <= 1e-7 * max(1.0, abs(OPT[name]))
    assert abs(r["obj"] - META[name]["netlib_optimum"]) <= 1e-7 * max(1.0, abs(OPT[name]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_new_interior_sparse_reaches_netlib_optimum_on_gpu(built_library, name):
    """The whole caller-side path on the GPU: general form -> standard form -> C ABI -> x in the caller's
    variables.  52 LPs of benchmarks_full, objective against main.py:1417-1616, feasibility on the host."""
    c, Aineq, bineq, Aeq, beq, lb, ub = gf.load_golden_general(name)
    res = gf.new_interior_sparse(c, Aeq=Aeq, beq=beq, Aineq=Aineq, bineq=bineq, lb=lb, ub=ub, tol=1e-8)
    if name in STALLS_AT_OPTIMUM:
        assert res.status in ("max_iter", "converged")      # depends on the rounding of the triangular solves
    else:
        assert res.status == "converged", (name, res.status, res.iterations)
    assert abs(res.objective - OPT[name]) <= 1e-6 * max(1.0, abs(OPT[name])), (res.objective, OPT[name])
    x = res.x
    assert abs(float((c.T @ x)[0, 0]) - res.objective) <= 1e-9 * max(1.0, abs(res.objective))
    scale = 1 + res.residuals["b"]                   # |b| of the standard form, the scale of the stopping rule
    if Aineq is not None:
        assert np.max(Aineq @ x - bineq) <= 1e-6 * scale
    if Aeq is not None:
        assert np.max(np.abs(Aeq @ x - beq)) <= 1e-6 * scale
    assert np.all(x >= lb - 1e-6 * (1 + np.abs(lb))) and np.all(x[np.isfinite(ub)] <= (ub + 1e-6 * (1 + np.abs(ub)))[np.isfinite(ub)])
