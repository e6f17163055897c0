"""SURVEY 8(f) row 3 on the GPU: step_size / predicted_stepsize_lb_ub / full_stepsize_lb_ub (main.py:325-547, 550-559,
629-660) through ipm_op_step_size_bounded against outputs frozen from the UNMODIFIED reference and against the CPU
restatement on vectors long enough for the multi-block reduction.  Divisions and minima are exact operations: equality."""
import os

import numpy as np
import pytest

from oracle import ipm_oracle as orc
from oracle.step_size_cases import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_size_cases.npz"))
BOUNDS = ("none", "ub", "lb", "both")


@pytest.fixture(scope="module")
def ipm(built_library):
    import interiorpointmethod_b200 as pkg
    return pkg


def test_against_frozen_reference_outputs(ipm):
    col = lambda v: None if v is None else v.reshape(-1, 1)      # noqa: E731  the reference's (n,1) columns
    for i, (n, flavour, x, s, dx, ds, lb, ub) in enumerate(cases()):
        for j, bounds in enumerate(BOUNDS):
            L = lb if bounds in ("lb", "both") else None
            U = ub if bounds in ("ub", "both") else None
            p = ipm.predicted_stepsize_lb_ub(col(dx), None, col(ds), col(x), col(s), col(L), col(U))
            c = ipm.full_stepsize_lb_ub(col(x), None, col(s), col(dx), None, col(ds), None, None, None, col(L), col(U))
            assert np.array_equal(np.array(p + c), GOLD["out"][i, j]), (n, flavour, bounds, p, c, GOLD["out"][i, j])


@pytest.mark.parametrize("n", [3000, 1 << 20])
def test_long_vectors_against_the_restatement(ipm, n):
    rng = np.random.default_rng(n)
    lb = rng.uniform(-1, 0.5, n); ub = lb + rng.uniform(0.5, 3, n)
    ub[rng.uniform(size=n) < 0.5] = np.inf
    x = np.maximum(lb + rng.uniform(0.1, 0.9, n) * np.minimum(ub - lb, 3.0), 1e-3)
    s = rng.uniform(0.1, 2, n); dx = rng.standard_normal(n); ds = rng.standard_normal(n)
    for bounds in BOUNDS:
        L = lb if bounds in ("lb", "both") else None
        U = ub if bounds in ("ub", "both") else None
        assert ipm.step_size(x, None, s, delta_aff=(dx, None, ds), lb=L, ub=U) == orc.step_size_bounded(x, s, dx, ds, L, U, False)
        assert ipm.step_size(x, None, s, delta=(dx, None, ds), lb=L, ub=U) == orc.step_size_bounded(x, s, dx, ds, L, U, True)
    assert ipm.step_size(x, None, s) is None
    with pytest.raises(ValueError):
        ipm.step_size(x, None, s, delta=(dx[:-1], None, ds))
