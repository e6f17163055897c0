"""CPU tests: the oracle (oracle/ipm_oracle.py) against the golden vectors frozen from the unmodified
reference (tests/golden/, made by oracle/make_golden.py) and against the known answers in the reference's
source (ex1 = -775 main.py:1253, ex2 = -15000 main.py:1261, Netlib optima main.py:1417-1516)."""
import os

import numpy as np
import pytest

from oracle import ipm_oracle as orc
from interiorpointmethod_b200 import load_golden_problem

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# small enough that the as-written full-KKT path finishes in about a second each
KKT_SET = ["AFIRO", "SC50A", "SC50B", "KB2", "SCSD1", "SHARE2B", "SC105", "STOCFOR1", "SC205", "E226", "SCTAP1"]
NORMAL_SET = ["AFIRO", "SC50A", "SC50B", "KB2", "SCSD1", "SHARE2B", "SC105", "STOCFOR1", "SCSD6", "SC205", "E226",
              "SCTAP1", "BANDM", "SCSD8", "GROW7"]
NETLIB_OPT = {"AFIRO": -4.6475314286e02, "SC50A": -6.4575077059e01, "SC50B": -7.0000000000e01,
              "SCSD1": 8.6666666743e00, "SHARE2B": -4.1573224074e02, "SC105": -5.2202061212e01,
              "STOCFOR1": -4.1131976219e04, "SCSD8": 9.0499999993e02, "E226": -1.8751929066e01}


@pytest.mark.parametrize("name", KKT_SET)
def test_as_written_path_reproduces_reference(name, reference_results):
    """linear='kkt' restates main.py:780-807 with the same scipy calls: same k, same objective up to the
    last digits (SuperLU's BLAS calls round differently with a different OpenBLAS thread count)."""
    A, b, c, cTlb = load_golden_problem(name)
    res = orc.solve(A, b, c, cTlb, tol=1e-8, linear="kkt")
    g = reference_results[name]
    assert res["k"] == g["k"]
    assert abs(res["obj"] - g["obj"]) <= 1e-12 * max(1.0, abs(g["obj"]))
    assert res["status"] == 0


@pytest.mark.parametrize("name", NORMAL_SET)
def test_normal_equations_path_matches_reference(name, reference_results):
    """The elimination the GPU uses: iteration count within 1, objective within 1e-8 relative (north_star)."""
    A, b, c, cTlb = load_golden_problem(name)
    res = orc.solve(A, b, c, cTlb, tol=1e-8, linear="normal")
    g = reference_results[name]
    assert abs(res["k"] - g["k"]) <= 1
    assert abs(res["obj"] - g["obj"]) <= 1e-8 * max(1.0, abs(g["obj"]))
    nrb, nrc, gap, nb, nc = orc.residual_norms(*_std(A, b, c), res["x"], res["y"], res["s"])
    assert nrb <= 1e-8 * (1 + nb) and nrc <= 1e-8 * (1 + nc) and gap <= 1e-8


def _std(A, b, c):
    from scipy import sparse
    return sparse.csr_matrix(A, dtype=np.float64), orc.as_column(b), orc.as_column(c)


@pytest.mark.parametrize("name", sorted(NETLIB_OPT))
def test_golden_objectives_agree_with_netlib_table(name, reference_results):
    """main.py:1417-1516 lists the Netlib optima to 11 digits; the frozen reference runs reproduce them."""
    g = reference_results[name]
    assert abs(g["obj"] - NETLIB_OPT[name]) <= 2e-8 * max(1.0, abs(NETLIB_OPT[name]))


def test_reference_outcome_on_25fv47_is_nan_at_k1(reference_results):
    assert reference_results["25FV47"]["reference_outcome"] == {"k": 1, "obj": None}


EX3_A = [[10, 7.5, 4, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 10, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0],
         [0.5, 0.4, 0.5, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0.4, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0],
         [0.5, 0.1, 0.5, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0], [0.4, 0.2, 0.4, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0],
         [1, 1.5, 0.5, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0],
         [0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0], [0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]]
EXAMPLES = {
    "ex1": ([[3, 6, 8], [8, 4, 1]], [30, 44], [-100, -125, -20], -775.0),
    "ex2": ([[1, 1.5, 1, 0, 0], [2, 3, 0, 1, 0], [2, 1, 0, 0, 1]], [750, 1500, 1000], [-20, -30, 0, 0, 0], -15000.0),
    "ex3": (EX3_A, [4350, 2500, 280, 140, 280, 140, 700, 300, 180, 400], [-300, -500, -200] + [0] * 10, -168000.0),
}


@pytest.mark.parametrize("name", sorted(EXAMPLES))
@pytest.mark.parametrize("linear", ["kkt", "normal"])
def test_dense_examples(name, linear, dense_results):
    A, b, c, known = EXAMPLES[name]
    res = orc.solve(np.array(A, float), b, c, tol=1e-8, max_iter=50000, y0_is_one=False, linear=linear)
    g = dense_results[name]
    assert abs(res["k"] - g["k"]) <= (0 if linear == "kkt" else 1)
    assert abs(res["obj"] - g["obj"]) <= 1e-8 * abs(g["obj"])
    assert abs(res["obj"] - known) <= 1e-6 * abs(known)
    if linear == "kkt":
        assert np.allclose(res["x"].ravel(), g["x"], rtol=0, atol=1e-9 * max(1.0, np.max(np.abs(g["x"]))))


@pytest.mark.parametrize("seed", [0, 1])
def test_synthetic_dense_generator_matches_golden(seed, dense_results):
    A, b, c = orc.synthetic_dense_lp(64, 128, seed)
    g = dense_results["synthetic_64x128_seed%d" % seed]
    for linear in ("kkt", "normal"):
        res = orc.solve(A, b, c, tol=1e-8, max_iter=50000, y0_is_one=False, linear=linear)
        assert res["k"] == g["k"]
        assert abs(res["obj"] - g["obj"]) <= 1e-8 * abs(g["obj"])


def test_op_level_vectors_match_reference_trace():
    """Per-op parity on AFIRO states frozen from the reference (directions via main.py:198-212/250-269,
    step lengths main.py:305-322/604-626, sigma main.py:588-601)."""
    A, b, c, _ = load_golden_problem("AFIRO")
    As, bc, cc = _std(A, b, c)
    tr = np.load(os.path.join(GOLDEN, "trace_AFIRO.npz"))
    for k in (0, 1, 10, 40):
        x, y, s = tr["k%d_x" % k], tr["k%d_y" % k], tr["k%d_s" % k]
        for linear in ("kkt", "normal"):
            info = {}
            orc.newton_iteration(As, bc, cc, x, y, s, linear=linear, info=info)
            last = info["last"]
            tol = 1e-12 if linear == "kkt" else 1e-9
            for key in ("dx_aff", "dy_aff", "ds_aff", "dx", "dy", "ds"):
                ref = tr["k%d_%s" % (k, key)]
                err = np.linalg.norm(last[key] - ref) / np.linalg.norm(ref)
                assert err <= tol, (k, linear, key, err)
            assert abs(last["sigma"] - float(tr["k%d_sigma" % k])) <= (1e-12 if linear == "kkt" else 1e-8) * abs(float(tr["k%d_sigma" % k]))
            assert np.allclose(last["alpha"], tr["k%d_alpha" % k], rtol=(1e-12 if linear == "kkt" else 1e-8), atol=0)


def test_safeguarded_cholesky_c_and_numpy_agree():
    rng = np.random.default_rng(0)
    B = rng.standard_normal((90, 200))
    M = B @ B.T
    M[:, 7] = 0.0
    M[7, :] = 0.0          # an empty row of A gives a zero row/column of M (25FV47)
    L1, n1 = orc.cholesky_safeguarded_numpy(M)
    L2, n2 = orc.cholesky_safeguarded(M)
    assert n1 == n2 == 1
    assert L1[7, 7] == L2[7, 7] == 1e64
    mask = np.ones(90, bool); mask[7] = False
    assert np.allclose(L1[np.ix_(mask, mask)], L2[np.ix_(mask, mask)], rtol=1e-10, atol=1e-12)
    assert np.allclose((L2 @ L2.T)[np.ix_(mask, mask)], M[np.ix_(mask, mask)], rtol=1e-10, atol=1e-10)


def test_ratio_test_edge_cases():
    x = np.array([[1.0], [2.0], [3.0]])
    assert orc.ratio_test(x, np.array([[1.0], [0.0], [2.0]])) == 1.0           # nothing blocks
    assert orc.ratio_test(x, np.array([[-4.0], [1.0], [-3.0]])) == 0.25
    assert orc.full_stepsize(x, x, np.ones_like(x), np.ones_like(x)) == (0.91, 0.91)   # alpha <= eta always (Q4)


def test_nan_input_stops_like_the_reference():
    """Any NaN makes every comparison in check_optimality False -> loop exits (main.py:170-173, Q5)."""
    A, b, c = orc.synthetic_dense_lp(8, 16, 0)
    b = b.copy(); b[0] = np.nan
    res = orc.solve(A, b, c, tol=1e-8, y0_is_one=False, linear="normal")
    assert res["k"] <= 1 and res["status"] == 2


def test_mehrotra_start_oracle_reaches_netlib_optima():
    """The opt-in starting point (not in the reference): the oracle with it converges on 25FV47, where the
    reference's start x = s = 1 ends in NaN, to the Netlib optimum of main.py:1417-1516, and needs 15 instead of 93
    iterations on AFIRO."""
    from oracle import ipm_oracle as orc
    from interiorpointmethod_b200.problems import load_golden_problem
    A, b, c, cTlb = load_golden_problem("AFIRO")
    r = orc.solve(A, b, c, cTlb=cTlb, tol=1e-8, start="mehrotra")
    assert r["status"] == 0 and r["k"] <= 20
    assert abs(r["obj"] - NETLIB_OPT["AFIRO"]) <= 1e-8 * abs(NETLIB_OPT["AFIRO"])
    A, b, c, cTlb = load_golden_problem("25FV47")
    r = orc.solve(A, b, c, cTlb=cTlb, tol=1e-8, start="mehrotra", max_iter=200)
    assert r["status"] == 0 and r["k"] <= 40
    assert abs(r["obj"] - 5.5018458883e03) <= 1e-8 * 5.5018458883e03


@pytest.mark.parametrize("seed", [0, 7466, 16893])
def test_fourpass_identities_against_literal_iteration(seed):
    """The two identities behind the batched solver's four-pass iteration (tests/fourpass_emulation.py), checked
    on the CPU against the literal iteration: LP 0, LP 7466 (the one that needed the periodic refresh on the GPU)
    and LP 16893 (the straggler of the weak-scaling workload)."""
    from oracle import ipm_oracle as orc
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fourpass_emulation", os.path.join(os.path.dirname(__file__), "fourpass_emulation.py"))
    emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu)
    solve_fourpass = emu.solve_fourpass
    A, b, c = orc.synthetic_dense_lp(256, 512, seed)
    lit = orc.solve(A, b, c, tol=1e-8, max_iter=50000, y0_is_one=False, linear="normal")
    k, x, y, s, hist = solve_fourpass(A, b, c, tol=1e-8, refresh_every=3)
    assert lit["status"] == 0 and abs(k - lit["k"]) <= 1
    obj = float((c.reshape(1, -1) @ x)[0, 0])
    assert abs(obj - lit["obj"]) <= 1e-8 * abs(lit["obj"])
    floor = 1e-8 * (1 + np.linalg.norm(b))
    for (_, rb_rec, rb_true, rc_rec, rc_true, lin_err) in hist:
        assert abs(rb_rec - rb_true) <= 1e-3 * max(rb_true, floor)      # the recurrences track the true residuals
        assert abs(rc_rec - rc_true) <= 1e-3 * max(rc_true, floor)
        assert lin_err <= 1e-9                                          # linearity of main.py:150-152 in r4


@pytest.mark.parametrize("seed", [16893, 31186])
def test_refined_corrector_converges_where_the_normal_equations_stall(seed):
    """LP 31186 of the generator: the reference as written (dense KKT + dgesv) needs 18 iterations, the literal
    iteration on the normal equations stalls in this container (|rb| stuck near 1e-5 once d_max/d_min passes 1e19;
    150+ iterations, also from Mehrotra's start) - one refinement step of the corrector, which is what the batched
    solver applies to restarted LPs, brings it back to 19.  LP 16893 is the GPU straggler of DESIGN.md section 4."""
    import importlib.util
    import os
    from oracle import ipm_oracle as orc
    spec = importlib.util.spec_from_file_location("fourpass_emulation", os.path.join(os.path.dirname(__file__), "fourpass_emulation.py"))
    emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu)
    A, b, c = orc.synthetic_dense_lp(256, 512, seed)
    kkt = orc.solve(A, b, c, tol=1e-8, max_iter=200, y0_is_one=False, linear="kkt")
    assert kkt["status"] == 0 and 15 <= kkt["k"] <= 20
    k, obj = emu.solve_refined(A, b, c, tol=1e-8, max_iter=60)
    assert abs(k - kkt["k"]) <= 2
    assert abs(obj - kkt["obj"]) <= 1e-8 * abs(kkt["obj"])


@pytest.mark.parametrize("seed", [16893, 31186])
def test_oracle_kkt_matches_reference_on_the_straggler_lps(seed, dense_results):
    """The as-written port against the unmodified reference's `interior` on the two generator LPs behind the batched
    solver's straggler handling (frozen by oracle/make_golden_stragglers.py)."""
    from oracle import ipm_oracle as orc
    g = dense_results["synthetic_256x512_seed%d" % seed]
    A, b, c = orc.synthetic_dense_lp(256, 512, seed)
    r = orc.solve(A, b, c, tol=1e-8, max_iter=50000, y0_is_one=False, linear="kkt")
    assert r["status"] == 0 and r["k"] == g["k"]
    assert abs(r["obj"] - g["obj"]) <= 1e-10 * abs(g["obj"])


# ------------------------------------------------------------------------------ batch tables (round 2)
def _batch_tables():
    import json
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ref = {int(k): v for k, v in json.load(open(os.path.join(gold, "batch_256x512_reference.json")))["seeds"].items()}
    tab = np.load(os.path.join(gold, "batch_256x512_oracle.npz"))
    code = tab["refinements"].astype(int)          # refinement steps, + 64 when the LP was handed off
    return ref, tab["k"].astype(int), tab["obj"], code % 64, code >= 64


def test_batch_tables_are_complete_and_agree_with_the_unmodified_reference():
    """tests/golden/batch_256x512_oracle.npz (oracle: normal equations + refinement rule + hand-off to the augmented
    system, all 65536 generator seeds) against tests/golden/batch_256x512_reference.json (the UNMODIFIED reference's
    `interior`, seeds 0..511 and the three seeds with a history): every LP converges in 15..20 iterations; on every
    seed both tables hold the iteration count is within +-1 (in fact equal) and the objective within 1e-8 relative."""
    ref, k, obj, nref, handed = _batch_tables()
    assert k.shape == (65536,) and (k >= 15).all() and (k <= 20).all() and np.isfinite(obj).all()
    seeds = np.array(sorted(ref))
    assert seeds.size >= 515 and {7466, 16893, 31186} <= set(ref)
    kr = np.array([ref[s][0] for s in seeds])
    orf = np.array([ref[s][1] for s in seeds])
    assert np.abs(k[seeds] - kr).max() <= 1
    assert (np.abs(obj[seeds] - orf) <= 1e-9 * np.maximum(1.0, np.abs(orf))).all()      # measured: 1.6e-10
    assert np.array_equal(k[seeds], kr)
    assert 0.01 < (nref > 0).mean() < 0.08           # the refinement step: about one LP in thirty
    assert 0.0002 < handed.mean() < 0.005            # the hand-off: about one LP in a thousand
    assert handed[[16893, 31186]].all()              # the two LPs with a history are among them


@pytest.mark.parametrize("seed", [0, 5, 7466, 7954, 16893, 31186, 40000, 54456, 65535])
def test_oracle_rule_reproduces_its_table_and_the_trap_is_real(seed):
    """The table is what oracle.solve(linear="normal", refine_thresh=0.1, handoff=True) returns (regenerated here for a
    few seeds); without refinement and hand-off LP 31186 is trapped (> 60 iterations) - the reason both exist."""
    from oracle import ipm_oracle as orc
    _, k, obj, nref, handed = _batch_tables()
    A, b, c = orc.synthetic_dense_lp(256, 512, seed)
    r = orc.solve(A, b, c, tol=1e-8, max_iter=150, y0_is_one=False, linear="normal", refine_thresh=0.1, handoff=True)
    assert r["status"] == 0 and r["k"] == k[seed] and r["refinements"] == nref[seed] and r["handoff"] == handed[seed]
    assert abs(r["obj"] - obj[seed]) <= 1e-12 * max(1.0, abs(obj[seed]))
    if seed == 31186:
        lit = orc.solve(A, b, c, tol=1e-8, max_iter=60, y0_is_one=False, linear="normal")
        assert lit["k"] == 60 and lit["status"] == 1


@pytest.mark.parametrize("seed", [0, 3, 16893, 31186])
def test_augmented_system_path_reproduces_the_reference(seed, dense_results):
    """oracle.solve(linear="augmented") - the executable spec of the GPU's augmented-system kernel - against the
    unmodified reference's `interior` (dense (m+2n) KKT + dgesv): same iteration count, objective to 1e-10."""
    from oracle import ipm_oracle as orc
    g = dense_results["synthetic_256x512_seed%d" % seed]
    A, b, c = orc.synthetic_dense_lp(256, 512, seed)
    r = orc.solve(A, b, c, tol=1e-8, max_iter=100, y0_is_one=False, linear="augmented")
    assert r["status"] == 0 and r["k"] == g["k"]
    assert abs(r["obj"] - g["obj"]) <= 1e-10 * abs(g["obj"])


def test_dependent_row_elimination_on_qap8():
    """Opt-in, not in the reference: rows of A that depend on the rows before them are found once from the factorisation
    of A A^T and removed from every later factorisation.  QAP8 (742 of 912 rows independent): 17-19
    iterations to the Netlib optimum 203.5 (main.py:1417-1516) instead of 196; a full-rank LP is untouched."""
    from oracle import ipm_oracle as orc
    import interiorpointmethod_b200.problems as P
    A, b, c, cTlb = P.load_golden_problem("QAP8")
    for start in ("reference", "mehrotra"):
        r = orc.solve(A, b, c, cTlb=cTlb, tol=1e-8, max_iter=100, start=start, dependent_tol=1e-10)
        assert r["status"] == 0 and r["k"] <= 25 and r["dependent_rows"] == 170
        assert abs(r["obj"] - 203.5) <= 1e-8 * 203.5
    A, b, c, cTlb = P.load_golden_problem("AFIRO")
    r0 = orc.solve(A, b, c, cTlb=cTlb, tol=1e-8)
    r1 = orc.solve(A, b, c, cTlb=cTlb, tol=1e-8, dependent_tol=1e-10)
    assert r1["dependent_rows"] == 0 and r1["k"] == r0["k"] and r1["obj"] == r0["obj"]
