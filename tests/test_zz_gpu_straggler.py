"""Straggler restart of the batched solver: stage 1 (restart under the literal six-pass
iteration) on the LP that needs it.  Kept in its own file, last in collection order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ipm(built_library):
    import interiorpointmethod_b200 as pkg
    return pkg


def test_straggler_restart_lp_16893(ipm, dense_results):
    """LP 16893 of the benchmark generator (second GPU's share of the weak-scaling workload): the four-pass
    iteration traps it at the boundary (3527 iterations), the literal six-pass iteration and the CPU port of the
    reference need 17-18.  The restart (ipm_batched_set_straggler_restart, default on) must hand it to the literal
    iteration; every LP of the batch stays within +-1 of the six-pass count and of the oracle."""
    from interiorpointmethod_b200 import _lib
    from interiorpointmethod_b200.batch import solve_batched_host
    from oracle import ipm_oracle as orc
    lib = _lib.load()
    first, B, at = 16864, 64, 16893 - 16864
    A, b, c = ipm.synthetic_dense_batch(first, B, 256, 512)
    try:
        lib.ipm_batched_set_variant(0, 3)
        obj6, it6, st6 = solve_batched_host(A, b, c, tol=1e-8)
        lib.ipm_batched_set_variant(1, 3)
        obj, it, st = solve_batched_host(A, b, c, tol=1e-8)
        assert (st == 0).all() and (st6 == 0).all()
        keep = np.arange(B) != at
        assert np.abs(it.astype(int) - it6.astype(int))[keep].max() <= 1, (it, it6)
        assert np.abs((obj - obj6) / obj6)[keep].max() <= 1e-8
        assert abs(obj[at] - obj6[at]) <= 1e-8 * abs(obj6[at])
        assert int(it[at]) == int(it6[at])            # the restarted LP ran the literal iteration from the start
        g = dense_results["synthetic_256x512_seed16893"]       # the unmodified reference's `interior`: k = 18
        assert abs(int(it[at]) - g["k"]) <= 1 and abs(obj[at] - g["obj"]) <= 1e-8 * abs(g["obj"])
        o = orc.solve(A[at], b[at], c[at], tol=1e-8, max_iter=50000, y0_is_one=False, linear="normal")
        assert abs(int(it[at]) - o["k"]) <= 1 and abs(obj[at] - o["obj"]) <= 1e-7 * abs(o["obj"])
        # without the restart the trap is there (documents why the restart exists; capped to keep the test short)
        lib.ipm_batched_set_straggler_restart(0)
        _, it0, st0 = solve_batched_host(A, b, c, tol=1e-8, max_iter=120)
        assert int(it0[at]) == 120 and int(st0[at]) == 1
        assert np.array_equal(it0[keep], it[keep])
    finally:
        lib.ipm_batched_set_variant(1, 3)
        lib.ipm_batched_set_straggler_restart(8)
