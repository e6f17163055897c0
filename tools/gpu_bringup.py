"""Bring-up script for a GPU box: exercises every kernel family against the CPU oracle and prints what it sees.
Not a test (tests/ holds those) and not a benchmark; it exists so one gpurun call answers "what works".

    python tools/gpu_bringup.py [--quick]
"""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import interiorpointmethod_b200 as ipm  # noqa: E402
from oracle import ipm_oracle as orc  # noqa: E402  (checker only)


def rel(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(np.linalg.norm(a - b) / max(1e-300, np.linalg.norm(b)))


def section(name):
    print("\n=== %s" % name, flush=True)


def check_ops(name, A, b, c, y0_one, iters=3):
    from scipy import sparse
    As = sparse.csr_matrix(A) if sparse.issparse(A) else np.asarray(A, float)
    bc, cc = orc.as_column(b), orc.as_column(c)
    m, n = As.shape
    x, y, s = orc.initial_point(m, n, y0_one)
    ns = ipm.NewtonStep(A, b, c)
    ns.set_state(x, y, s)
    for k in range(iters):
        nrm = ns.residual_norms()
        onrm = orc.residual_norms(As, bc, cc, x, y, s)
        ns.assemble_normal()
        M = np.tril(ns.get_M())
        Mo = np.tril(orc.normal_matrix(As, x, s))
        nf = ns.factor(1e-30)
        L = np.tril(ns.get_M())
        Lo, nfo = orc.cholesky_safeguarded(Mo)
        info = {}
        x2, y2, s2 = orc.newton_iteration(As, bc, cc, x, y, s, linear="normal", info=info)
        last = info["last"]
        dxa, dya, dsa = ns.direction(0)
        aaff = ns.ratio_test(0)
        sig = ns.sigma()
        dx, dy, ds = ns.direction(1)
        al = ns.ratio_test(1)
        ns.update(*al)
        gx, gy, gs = ns.get_state()
        print("%s it%d norms %.2e M %.2e L %.2e nfix %d/%d dxa %.2e dya %.2e dsa %.2e sigma %.3e/%.3e dx %.2e alpha (%.4g,%.4g)/(%.4g,%.4g) x %.2e"
              % (name, k, max(abs(nrm[k2] - v) / max(1e-300, abs(v)) for k2, v in zip(("rb", "rc", "gap", "b", "c"), onrm)),
                 rel(M, Mo), rel(L, Lo), nf, nfo, rel(dxa, last["dx_aff"]), rel(dya, last["dy_aff"]),
                 rel(dsa, last["ds_aff"]), sig[2], last["sigma"], rel(dx, last["dx"]), al[0], al[1],
                 last["alpha"][0], last["alpha"][1], rel(gx, x2)), flush=True)
        x, y, s = x2, y2, s2
        ns.set_state(x, y, s)
    ns.close()


def main():
    quick = "--quick" in sys.argv
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_results.json")))
    dense_gold = json.load(open(os.path.join(ROOT, "tests", "golden", "dense_results.json")))

    section("op-level parity (sparse AFIRO, SCSD8; dense 64x128)")
    for name in ("AFIRO", "SCSD8"):
        try:
            A, b, c, cTlb = ipm.load_golden_problem(name)
            check_ops(name, A, b, c, True)
        except Exception:
            traceback.print_exc()
    try:
        A, b, c = ipm.synthetic_dense_lp(64, 128, 0)
        check_ops("dense64x128", A, b, c, False)
        A, b, c = ipm.synthetic_dense_lp(300, 700, 1)
        check_ops("dense300x700", A, b, c, False)
    except Exception:
        traceback.print_exc()

    section("whole solves, sparse Netlib")
    names = ["AFIRO", "SC50A", "SC50B", "KB2", "SCSD1", "SHARE2B", "SC105", "STOCFOR1", "SCSD6", "SC205", "E226",
             "SCTAP1", "BANDM", "SCSD8", "GROW7", "DEGEN2", "GROW15", "TRUSS", "FIT1P", "SCTAP2", "WOODW", "GROW22",
             "SCTAP3", "STOCFOR2", "MAROS-R7", "25FV47", "QAP8", "QAP15", "STOCFOR3"]
    if quick:
        names = names[:14]
    for name in names:
        try:
            A, b, c, cTlb = ipm.load_golden_problem(name)
            t0 = time.time()
            ns = ipm.NewtonStep(A, b, c)
            t1 = time.time()
            r = ns.solve(tol=1e-8, max_iter=5000, y0_is_one=True, cTlb=cTlb)
            t2 = time.time()
            ns.close()
            g = gold.get(name, {})
            ref = ("ref k=%s obj=%s" % (g.get("k"), g.get("obj"))) if "k" in g else "ref: n/a"
            relerr = abs(r.objective - g["obj"]) / max(1.0, abs(g["obj"])) if "obj" in g else float("nan")
            print("%-9s m=%5d n=%5d k=%4d obj=%.12e status=%s rel=%.2e | %s | load %.3fs solve %.3fs (%.1f it/s) rb=%.1e rc=%.1e gap=%.1e"
                  % (name, ns.m, ns.n, r.iterations, r.objective, r.status, relerr, ref, t1 - t0, t2 - t1,
                     r.iterations / max(t2 - t1, 1e-9), r.residuals["rb"], r.residuals["rc"], r.residuals["gap"]),
                  flush=True)
        except Exception:
            traceback.print_exc()

    section("whole solves, dense")
    ex = {
        "ex1": ([[3, 6, 8], [8, 4, 1]], [30, 44], [-100, -125, -20]),
        "ex2": ([[1, 1.5, 1, 0, 0], [2, 3, 0, 1, 0], [2, 1, 0, 0, 1]], [750, 1500, 1000], [-20, -30, 0, 0, 0]),
    }
    for nm, (A, b, c) in ex.items():
        try:
            r = ipm.interior(A, b, c, tol=1e-8)
            print(nm, r.iterations, r.objective, r.status, "ref", dense_gold[nm]["k"], dense_gold[nm]["obj"], flush=True)
        except Exception:
            traceback.print_exc()
    for (m, n, seed) in ((64, 128, 0), (64, 128, 1), (256, 512, 0), (256, 512, 1), (256, 512, 2), (256, 512, 3)):
        try:
            A, b, c = ipm.synthetic_dense_lp(m, n, seed)
            t0 = time.time()
            r = ipm.interior(A, b, c, tol=1e-8)
            g = dense_gold["synthetic_%dx%d_seed%d" % (m, n, seed)]
            print("dense %dx%d seed%d k=%d obj=%.12e %s | ref k=%d obj=%.12e rel=%.2e | %.3fs"
                  % (m, n, seed, r.iterations, r.objective, r.status, g["k"], g["obj"],
                     abs(r.objective - g["obj"]) / abs(g["obj"]), time.time() - t0), flush=True)
        except Exception:
            traceback.print_exc()

    section("batched dense (host-pointer entry point)")
    try:
        from interiorpointmethod_b200.batch import solve_batched_host
        for (B, m, n) in ((8, 64, 128), (64, 256, 512)):
            A, b, c = ipm.synthetic_dense_batch(0, B, m, n)
            t0 = time.time()
            obj, its, st = solve_batched_host(A, b, c, tol=1e-8, max_iter=50000)
            dt = time.time() - t0
            print("batched B=%d %dx%d: iters %s status %s  %.3fs" % (B, m, n, its[:8], st[:8], dt))
            for i in range(min(B, 4)):
                key = "synthetic_%dx%d_seed%d" % (m, n, i)
                if key in dense_gold:
                    g = dense_gold[key]
                    print("   LP%d k=%d obj=%.12e | ref k=%d obj=%.12e rel=%.2e"
                          % (i, its[i], obj[i], g["k"], g["obj"], abs(obj[i] - g["obj"]) / abs(g["obj"])), flush=True)
    except Exception:
        traceback.print_exc()

    from interiorpointmethod_b200 import _lib
    print("\nkernel launches:", _lib.load().ipm_launch_count())


if __name__ == "__main__":
    main()
