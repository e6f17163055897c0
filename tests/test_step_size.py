"""SURVEY 8(f) row 3 - the bounded-variable ratio tests (step_size, main.py:325-547): the CPU restatement against outputs
frozen from the UNMODIFIED reference (tests/golden/step_size_cases.npz, oracle/make_golden_step_size.py), and against
the live reference where it is present."""
import os

import numpy as np
import pytest

from oracle import ipm_oracle as orc
from oracle import ref_harness
from oracle.step_size_cases import cases, checksum

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_size_cases.npz"))
BOUNDS = ("none", "ub", "lb", "both")


def bound_args(bounds, lb, ub):
    return (lb if bounds in ("lb", "both") else None), (ub if bounds in ("ub", "both") else None)


def test_inputs_are_the_frozen_ones():
    cs = cases()
    assert len(cs) == GOLD["out"].shape[0] == 30
    for i, (n, flavour, x, s, dx, ds, lb, ub) in enumerate(cs):
        assert checksum(x, s, dx, ds, lb, ub) == GOLD["input_checksum"][i], (i, n, flavour)


def test_oracle_equals_frozen_reference_outputs():
    for i, (n, flavour, x, s, dx, ds, lb, ub) in enumerate(cases()):
        for j, bounds in enumerate(BOUNDS):
            L, U = bound_args(bounds, lb, ub)
            got = orc.step_size_bounded(x, s, dx, ds, L, U, corrector=False) + orc.step_size_bounded(x, s, dx, ds, L, U, corrector=True)
            assert np.array_equal(np.array(got), GOLD["out"][i, j]), (n, flavour, bounds, got, GOLD["out"][i, j])


def test_reference_quirks_are_in_the_table():
    out = GOLD["out"]
    assert (out[:, 0, 3] == 1.0).all()              # corrector without bounds: alpha_dual = 1 (main.py:449-454)
    flav = [c[1] for c in cases()]
    i = flav.index("all_dx_positive")
    assert out[i, 0, 0] == 1.0 and out[i, 0, 2] == 0.91       # empty index set: 1, then eta
    assert out[i, 2, 0] == 1.0 and out[i, 2, 2] == 0.91       # lb only: a growing x never meets its bound


@pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")
def test_oracle_equals_live_reference():
    ref_main, _ = ref_harness.load_reference()
    rng = np.random.default_rng(5)
    col = lambda v: None if v is None else np.asarray(v).reshape(-1, 1)      # noqa: E731
    for trial in range(40):
        n = int(rng.integers(1, 200))
        lb = rng.uniform(-1, 0.5, n); ub = lb + rng.uniform(0.5, 3, n)
        x = np.maximum(lb + rng.uniform(0.1, 0.9, n) * (ub - lb), 1e-3); ub = np.maximum(ub, x + 1e-3)
        s = rng.uniform(0.1, 2, n); dx = rng.standard_normal(n); ds = rng.standard_normal(n)
        for bounds in BOUNDS:
            L, U = bound_args(bounds, lb, ub)
            with np.errstate(all="ignore"):
                r0 = ref_main.step_size(col(x), None, col(s), delta_aff=(col(dx), None, col(ds)), lb=col(L), ub=col(U))
                r1 = ref_main.step_size(col(x), None, col(s), delta=(col(dx), None, col(ds)), lb=col(L), ub=col(U))
            assert tuple(float(v) for v in r0) == orc.step_size_bounded(x, s, dx, ds, L, U, corrector=False)
            assert tuple(float(v) for v in r1) == orc.step_size_bounded(x, s, dx, ds, L, U, corrector=True)
