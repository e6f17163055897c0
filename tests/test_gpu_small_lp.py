"""The one-launch solve of small sparse LPs (csrc/small_lp.cuh, ipm_set_small_lp_fused) against the multi-kernel path
of ipm_solve: the same device functions with the same block size, so iteration counts, objectives and iterates are
BITWISE equal - on every golden Netlib LP that is eligible (n <= 512, m <= 256), from both starting points."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLE = json.load(open(os.path.join(GOLD, "netlib_all.json")))["problems"]
ELIGIBLE = sorted(k for k, e in TABLE.items() if e["finite"] and e["n"] <= 512 and e["m"] <= 256)


@pytest.fixture(scope="module")
def ipm(built_library):
    import interiorpointmethod_b200 as pkg
    return pkg


def test_some_lps_are_eligible():
    assert "AFIRO" in ELIGIBLE and len(ELIGIBLE) >= 8


@pytest.mark.parametrize("name", ELIGIBLE)
@pytest.mark.parametrize("start", ["reference", "mehrotra"])
def test_one_launch_solve_bitwise_equal(ipm, name, start):
    from interiorpointmethod_b200 import _lib
    lib = _lib.load()
    A, b, c, cTlb = ipm.load_golden_problem(name)
    out = []
    try:
        for fused in (0, 1):
            assert lib.ipm_set_small_lp_fused(fused) == 0
            with ipm.NewtonStep(A, b, c) as ns:
                kw = dict(start="mehrotra") if start == "mehrotra" else {}
                r = ns.solve(tol=1e-8, max_iter=400, cTlb=cTlb, **kw)
                out.append((r.status, r.iterations, r.objective, np.asarray(r.x).copy()))
    finally:
        lib.ipm_set_small_lp_fused(1)
    (s0, k0, o0, x0), (s1, k1, o1, x1) = out
    assert s0 == s1 and k0 == k1
    assert (o0 == o1) or (np.isnan(o0) and np.isnan(o1))
    assert np.array_equal(x0, x1, equal_nan=True)


def test_iteration_cap_and_reference_count(ipm, reference_results):
    """max_iter is honoured inside the kernel; AFIRO takes the reference's 93 iterations (main.py:776-815)."""
    A, b, c, cTlb = ipm.load_golden_problem("AFIRO")
    with ipm.NewtonStep(A, b, c) as ns:
        r = ns.solve(tol=1e-8, max_iter=7, cTlb=cTlb)
        assert r.status == "max_iter" and r.iterations == 7
        r = ns.solve(tol=1e-8, max_iter=400, cTlb=cTlb)
        assert r.status == "converged" and r.iterations == reference_results["AFIRO"]["k"]
        assert abs(r.objective - reference_results["AFIRO"]["obj"]) <= 1e-8 * abs(reference_results["AFIRO"]["obj"])
