"""Device-side ingestion (ipm_load_csc / ipm_load_csr, csrc/ingest.cuh): the structure the
GPU builds from the loader's compressed-sparse arrays - the other orientation and the symbolic pattern of
M = A diag(d) A^T - must equal, integer for integer, (a) what scipy derives from the same matrix and (b) the
host routine it replaces, for both input orientations; and the per-structure cache must hand the same arrays back.
Index work: the bar is bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "problems")
ALL = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))
KEYS = ("rowptr", "colind", "t_rowptr", "t_colind", "out_idx", "prod_ptr", "pa", "pb", "val", "t_val")


@pytest.fixture(scope="module")
def ipm(built_library):
    import interiorpointmethod_b200 as pkg
    return pkg


@pytest.fixture()
def lib(ipm):
    from interiorpointmethod_b200 import _lib
    L = _lib.load()
    L.ipm_set_ingest_mode(0, 1)
    L.ipm_pattern_cache_clear()
    yield L
    L.ipm_set_ingest_mode(0, 1)
    L.ipm_pattern_cache_clear()


def _expected_structure(A):
    """Independent restatement with scipy: both orientations, entry list and term counts of tril(A A^T)."""
    from scipy import sparse
    csr = sparse.csr_matrix(A, dtype=np.float64)
    csr.sort_indices()
    csc = sparse.csc_matrix(A, dtype=np.float64)
    csc.sort_indices()
    m = A.shape[0]
    ldm = (m + 15) // 16 * 16
    S = sparse.csr_matrix((np.ones(csr.nnz, dtype=np.int64), csr.indices, csr.indptr), shape=A.shape)
    C = sparse.tril(S @ S.T, format="csr")          # C_ij = number of shared columns = terms of M_ij
    C.sort_indices()
    rows = np.repeat(np.arange(m), np.diff(C.indptr))
    out_idx = rows.astype(np.int64) * ldm + C.indices
    prod_ptr = np.concatenate([[0], np.cumsum(C.data)]).astype(np.int64)
    return csr, csc, out_idx, prod_ptr, rows, C.indices


def _check_against_scipy(A, pat):
    csr, csc, out_idx, prod_ptr, ei, ej = _expected_structure(A)
    assert np.array_equal(pat["rowptr"], csr.indptr) and np.array_equal(pat["colind"], csr.indices)
    assert np.array_equal(pat["t_rowptr"], csc.indptr) and np.array_equal(pat["t_colind"], csc.indices)
    assert np.array_equal(pat["val"], csr.data) and np.array_equal(pat["t_val"], csc.data)
    assert np.array_equal(pat["out_idx"], out_idx)
    assert np.array_equal(pat["prod_ptr"], prod_ptr)
    # every term (pa, pb) of entry (i, j): a_ik from row i, a_jk from row j, same k, k strictly ascending
    nt = int(prod_ptr[-1])
    assert pat["pa"].size == nt and pat["pb"].size == nt
    if nt:
        row_of = np.repeat(np.arange(A.shape[0]), np.diff(csr.indptr))
        ent_of = np.repeat(np.arange(out_idx.size), np.diff(prod_ptr))
        assert np.array_equal(row_of[pat["pa"]], ei[ent_of])
        assert np.array_equal(row_of[pat["pb"]], ej[ent_of])
        ka, kb = csr.indices[pat["pa"]], csr.indices[pat["pb"]]
        assert np.array_equal(ka, kb)
        same = ent_of[1:] == ent_of[:-1]
        assert np.all(ka[1:][same] > ka[:-1][same])


def _patterns(ipm, lib, A, b, c):
    """(device build from CSC, device build from CSR, host symbolic) for the same matrix, cache bypassed."""
    from scipy import sparse
    out = []
    for host_symbolic, M in ((0, sparse.csc_matrix(A)), (0, sparse.csr_matrix(A)), (1, sparse.csc_matrix(A))):
        lib.ipm_set_ingest_mode(host_symbolic, 0)
        with ipm.NewtonStep(M, b, c) as ns:
            info = ns.pattern_info()
            assert info["device_built"] == (not host_symbolic) and not info["cache_hit"]
            out.append(ns.pattern())
    lib.ipm_set_ingest_mode(0, 1)
    return out


@pytest.mark.parametrize("name", ALL)
def test_device_structure_equals_scipy_and_host_routine(ipm, lib, name):
    A, b, c, _ = ipm.load_golden_problem(name)
    dev_csc, dev_csr, host = _patterns(ipm, lib, A, b, c)
    for k in KEYS:
        assert np.array_equal(dev_csc[k], host[k]), k
        assert np.array_equal(dev_csr[k], host[k]), k
    _check_against_scipy(A, dev_csc)


def _ragged_cases():
    from scipy import sparse
    rng = np.random.default_rng(7)
    cases = {}
    cases["one_by_one"] = sparse.csc_matrix(np.array([[2.5]]))
    cases["single_row"] = sparse.csc_matrix(rng.standard_normal((1, 37)))
    cases["single_col"] = sparse.csc_matrix(rng.standard_normal((41, 1)))
    E = sparse.random(60, 90, density=0.05, random_state=3, format="lil")
    E[7, :] = 0          # empty row (25FV47 has one)
    E[:, 11] = 0         # empty column
    E[59, :] = 0         # last row empty
    cases["empty_row_and_col"] = sparse.csc_matrix(E)
    cases["all_zero"] = sparse.csc_matrix((5, 9))
    # a dense row and a dense column longer than the shared-memory staging of the transposition (2048) and wider
    # than one scan chunk (1024) / one CTA (256)
    D = sparse.random(2500, 2700, density=0.002, random_state=5, format="lil")
    D[3, :] = rng.standard_normal(2700)
    D[:, 5] = rng.standard_normal((2500, 1))
    cases["dense_row_and_col"] = sparse.csc_matrix(D)
    cases["dense_block"] = sparse.csc_matrix(rng.standard_normal((300, 77)))
    X = sparse.random(700, 1500, density=0.01, random_state=9, format="csc")
    X.data[::7] = 0.0    # explicit zeros stay in the structure
    cases["explicit_zeros"] = X
    return cases


@pytest.mark.parametrize("case", ["one_by_one", "single_row", "single_col", "empty_row_and_col", "all_zero",
                                  "dense_row_and_col", "dense_block", "explicit_zeros"])
def test_ragged_structures(ipm, lib, case):
    A = _ragged_cases()[case]
    m, n = A.shape
    b, c = np.ones(m), np.ones(n)
    dev_csc, dev_csr, host = _patterns(ipm, lib, A, b, c)
    for k in KEYS:
        assert np.array_equal(dev_csc[k], host[k]), k
        assert np.array_equal(dev_csr[k], host[k]), k
    _check_against_scipy(A, dev_csc)


def test_numeric_product_on_device_pattern(ipm, lib):
    """M = A diag(x/s) A^T assembled on the device-built pattern equals scipy's product (main.py:223-224)."""
    from scipy import sparse
    A, b, c, _ = ipm.load_golden_problem("SCSD8")
    rng = np.random.default_rng(0)
    with ipm.NewtonStep(A, b, c) as ns:
        x, s, y = rng.uniform(0.5, 2, ns.n), rng.uniform(0.5, 2, ns.n), rng.standard_normal(ns.m)
        ns.set_state(x, y, s)
        ns.residual_norms()
        ns.assemble_normal()
        M = np.tril(ns.get_M())
    Ar = sparse.csr_matrix(A, dtype=np.float64)
    ref = np.tril(((Ar @ sparse.diags(x / s)) @ Ar.T).toarray())
    assert np.max(np.abs(M - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_cache_hit_same_structure_new_values(ipm, lib):
    from scipy import sparse
    A, b, c, cT = ipm.load_golden_problem("SCTAP1")
    stats, stats0 = np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int64)
    lib.ipm_pattern_cache_stats(stats0.ctypes.data)          # hit/miss counters are cumulative
    with ipm.NewtonStep(A, b, c) as first:
        i1 = first.pattern_info()
        assert i1["device_built"] and not i1["cache_hit"]
        p1 = first.pattern()
        r1 = first.solve(tol=1e-8, cTlb=cT)
        # same LP again: hit, identical arrays, identical solve (bitwise)
        with ipm.NewtonStep(A, b, c) as again:
            assert again.pattern_info()["cache_hit"]
            p2 = again.pattern()
            for k in KEYS:
                assert np.array_equal(p1[k], p2[k]), k
            r2 = again.solve(tol=1e-8, cTlb=cT)
            assert r2.iterations == r1.iterations and r2.objective == r1.objective
        # new values on the same structure: hit; the value arrays follow, the pattern stays
        A2 = sparse.csc_matrix(A, dtype=np.float64, copy=True)
        A2.data *= np.linspace(0.5, 1.5, A2.nnz)
        with ipm.NewtonStep(A2, 2.0 * b, c) as scaled:
            assert scaled.pattern_info()["cache_hit"]
            p3 = scaled.pattern()
            for k in ("rowptr", "colind", "t_rowptr", "t_colind", "out_idx", "prod_ptr", "pa", "pb"):
                assert np.array_equal(p1[k], p3[k]), k
            assert np.array_equal(p3["t_val"], A2.data)
            assert np.array_equal(p3["val"], sparse.csr_matrix(A2).data)
        # the other orientation is a different key
        with ipm.NewtonStep(sparse.csr_matrix(A), b, c) as as_csr:
            assert not as_csr.pattern_info()["cache_hit"]
        lib.ipm_pattern_cache_stats(stats.ctypes.data)
        assert stats[0] == 2 and stats[1] - stats0[1] == 2 and stats[2] - stats0[2] == 2 and stats[3] > 0
        # clearing the cache does not pull the structure from under a live handle
        lib.ipm_release_cached()
        lib.ipm_pattern_cache_stats(stats.ctypes.data)
        assert stats[0] == 0 and stats[3] == 0
        r3 = first.solve(tol=1e-8, cTlb=cT)
        assert r3.iterations == r1.iterations and r3.objective == r1.objective
    with ipm.NewtonStep(A, b, c) as rebuilt:
        assert not rebuilt.pattern_info()["cache_hit"]


def test_solve_identical_with_host_and_device_pattern(ipm, lib, reference_results):
    """Same pattern => same bits: the Netlib parity results do not depend on where the pattern was built."""
    for name in ("AFIRO", "SC50A", "SHARE2B"):
        A, b, c, cT = ipm.load_golden_problem(name)
        res = []
        for host_symbolic in (0, 1):
            lib.ipm_set_ingest_mode(host_symbolic, 0)
            res.append(ipm.solve(A, b, c, tol=1e-8, cTlb=cT))
        assert res[0].iterations == res[1].iterations == reference_results[name]["k"]
        assert res[0].objective == res[1].objective
        assert np.array_equal(res[0].x, res[1].x)


def test_bad_structure_is_rejected(ipm, lib):
    import ctypes
    from interiorpointmethod_b200 import _lib
    h = ctypes.c_void_p()
    assert lib.ipm_create(ctypes.byref(h), 0) == 0
    try:
        b, c, val = np.ones(2), np.ones(3), np.ones(3)
        def call(fn, ptr, idx):
            ptr, idx = np.asarray(ptr, np.int32), np.asarray(idx, np.int32)
            return fn(h, 2, 3, 3, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data, b.ctypes.data, c.ctypes.data)
        assert call(lib.ipm_load_csc, [0, 1, 2, 3], [0, 1, 0]) == 0
        assert call(lib.ipm_load_csc, [0, 2, 2, 3], [1, 0, 0]) == -3        # rows not ascending inside a column
        assert call(lib.ipm_load_csc, [0, 1, 2, 3], [0, 2, 0]) == -3        # row index out of range
        assert call(lib.ipm_load_csc, [0, 1, 2, 2], [0, 1, 0]) == -3        # colptr[n] != nnz
        assert call(lib.ipm_load_csr, [0, 2, 3], [0, 2, 1]) == 0
        assert call(lib.ipm_load_csr, [0, 2, 3], [2, 0, 1]) == -3
        assert b"ascending" in lib.ipm_last_error(h)
        assert lib.ipm_load_csc(h, 2, 3, 3, None, None, None, None, None) == -2
        assert _lib.ERRORS[-3] == "IPM_ERR_SHAPE"
    finally:
        lib.ipm_destroy(h)
