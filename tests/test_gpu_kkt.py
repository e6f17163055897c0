"""The augmented-system kernel (csrc/kkt_dense.cuh, ipm_solve_dense_kkt): the reference's dense route - unreduced KKT
matrix (main.py:13-21) + LAPACK dgesv (main.py:178) - on the GPU with ds eliminated exactly.  Same Newton system and
the same pivoting rule as the reference, so the iteration counts of the frozen reference runs must be reproduced
EXACTLY (tests/golden/dense_results.json, batch_256x512_reference.json), the objectives to 1e-9."""
import json
import os

import numpy as np
import pytest

from test_gpu_parity import EXAMPLES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ipm(built_library):
    import interiorpointmethod_b200 as pkg
    return pkg


@pytest.mark.parametrize("name", sorted(EXAMPLES))
def test_kkt_examples(ipm, name, dense_results):
    A, b, c = EXAMPLES[name]
    res = ipm.interior_kkt(A, b, c, tol=1e-8)
    g = dense_results[name]
    assert res.status == "converged" and res.iterations == g["k"], (res.status, res.iterations, g["k"])
    assert abs(res.objective - g["obj"]) <= 1e-9 * abs(g["obj"])
    if name in ("ex1", "ex3"):
        assert np.allclose(res.x.ravel(), g["x"], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("shape,seed", [((64, 128), 0), ((64, 128), 1), ((256, 512), 0), ((256, 512), 1), ((256, 512), 2),
                                         ((256, 512), 3), ((256, 512), 16893), ((256, 512), 31186), ((256, 512), 7466)])
def test_kkt_synthetic_reproduces_the_reference_iteration_counts(ipm, dense_results, shape, seed):
    m, n = shape
    key = "synthetic_%dx%d_seed%d" % (m, n, seed)
    if key in dense_results:
        k_ref, obj_ref = dense_results[key]["k"], dense_results[key]["obj"]
    else:
        k_ref, obj_ref = json.load(open(os.path.join(GOLD, "batch_256x512_reference.json")))["seeds"][str(seed)]
    A, b, c = ipm.synthetic_dense_lp(m, n, seed)
    res = ipm.interior_kkt(A, b, c, tol=1e-8)
    assert res.status == "converged" and res.iterations == k_ref, (res.status, res.iterations, k_ref)
    assert abs(res.objective - obj_ref) <= 1e-9 * max(1.0, abs(obj_ref))
    x = res.x.ravel()
    assert (x > 0).all()
    assert np.linalg.norm(A @ x - b) <= 1.001e-8 * (1 + np.linalg.norm(b))           # main.py:170


@pytest.mark.parametrize("m,n", [(5, 9), (17, 40), (33, 47), (100, 300), (31, 16 * 7 + 3)])
def test_kkt_odd_shapes_against_the_oracle(ipm, m, n):
    """Orders that are not multiples of the panel width / block size, against oracle.solve(linear="augmented")."""
    from oracle import ipm_oracle as orc
    A, b, c = ipm.synthetic_dense_lp(m, n, 11 * m + n)
    res = ipm.interior_kkt(A, b, c, tol=1e-8)
    o = orc.solve(A, b, c, tol=1e-8, max_iter=200, y0_is_one=False, linear="augmented")
    assert o["status"] == 0 and res.status == "converged"
    assert res.iterations == o["k"]
    assert abs(res.objective - o["obj"]) <= 1e-9 * max(1.0, abs(o["obj"]))
    assert np.allclose(res.x.ravel(), o["x"].ravel(), rtol=1e-6, atol=1e-9)


def test_kkt_iteration_cap_and_nan(ipm):
    A, b, c = ipm.synthetic_dense_lp(20, 50, 3)
    r = ipm.interior_kkt(A, b, c, tol=1e-8, max_iter=3)
    assert r.iterations == 3 and r.status == "max_iter"
    b2 = b.copy(); b2[0] = np.nan
    r = ipm.interior_kkt(A, b2, c, tol=1e-8)
    assert r.status == "nan" and r.iterations <= 1


@pytest.mark.parametrize("shape,seed", [((64, 128), 0), ((256, 512), 3), ((33, 47), 5)])
def test_kkt_cluster_sizes_give_bitwise_equal_results(ipm, shape, seed):
    """The CTAs of a cluster split only the column-parallel work of the factorisation; everything else is redundant
    and deterministic, so 1, 2, 4 and 8 CTAs per LP must agree bit for bit."""
    from interiorpointmethod_b200 import _lib
    lib = _lib.load()
    m, n = shape
    A, b, c = ipm.synthetic_dense_lp(m, n, seed)
    res = {}
    try:
        for cl in (1, 2, 4, 8):
            assert lib.ipm_set_kkt_cluster(cl) == 0
            res[cl] = ipm.interior_kkt(A, b, c, tol=1e-8)
    finally:
        lib.ipm_set_kkt_cluster(4)
    for cl in (2, 4, 8):
        assert res[cl].status == "converged" and res[cl].iterations == res[1].iterations
        assert res[cl].objective == res[1].objective and np.array_equal(res[cl].x, res[1].x)
