"""TEST INFRASTRUCTURE — freezes golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py [--only NAME ...] [--skip-slow]

Writes, under tests/golden/:
  problems/<NAME>.npz         the standard-form LP exactly as create_problem_from_mps returns it
                              (sparse_interior.py:211-216): CSC arrays of A (cast to float64), b, c, cTlb
  reference_results.json      per problem: iteration count k and objective of the reference's
                              interior_sparse semantics (replayed main.py:780-807, tol=1e-8), wall
                              seconds, library versions
  trace_<NAME>.npz            per-iteration op-level vectors (state, directions, step lengths, sigma)
  dense_results.json          reference `interior` (main.py:707-757) on ex1..ex3 and on the synthetic
                              dense generator
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# the 26 LPs on which the reference converges
CONVERGED = ["AFIRO", "BANDM", "DEGEN2", "E226", "FIT1P", "GROW15", "GROW22", "GROW7", "KB2", "MAROS-R7",
             "SC105", "SC205", "SC50A", "SC50B", "SCSD1", "SCSD6", "SCSD8", "SCTAP1", "SCTAP2", "SCTAP3",
             "SHARE2B", "STOCFOR1", "STOCFOR2", "STOCFOR3", "TRUSS", "WOODW"]
# BASELINE.json configs on which the reference fails (data only + reference outcome where cheap)
DATA_ONLY = ["25FV47", "QAP15", "QAP8"]
SLOW = {"WOODW", "MAROS-R7", "STOCFOR3", "TRUSS"}
TRACES = {"AFIRO": (0, 1, 10, 40, 60, 92), "SCSD8": (0, 5, 19), "E226": (0, 31)}


def save_problem(name):
    from oracle import ref_harness as rh
    from scipy import sparse

    A, b, c, cTlb = rh.load_problem(name)
    A = sparse.csc_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    os.makedirs(os.path.join(GOLD, "problems"), exist_ok=True)
    np.savez_compressed(
        os.path.join(GOLD, "problems", name + ".npz"),
        m=A.shape[0], n=A.shape[1],
        indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32),
        data=A.data.astype(np.float64),
        b=np.asarray(b, dtype=np.float64).ravel(), c=np.asarray(c, dtype=np.float64).ravel(),
        cTlb=np.float64(cTlb),
    )
    return A, b, c, cTlb


def run_one(name):
    from oracle import ref_harness as rh

    A, b, c, cTlb = save_problem(name)
    t0 = time.time()
    res = rh.replay_interior_sparse(A, b, c, cTlb, tol=1e-8, trace_at=TRACES.get(name, ()))
    wall = time.time() - t0
    if name in TRACES:
        flat = {}
        for k, tr in res["trace"].items():
            for key, val in tr.items():
                flat["k%d_%s" % (k, key)] = np.asarray(val, dtype=np.float64)
        np.savez_compressed(os.path.join(GOLD, "trace_%s.npz" % name), **flat)
    x = res["x"]
    return name, dict(k=int(res["k"]), obj=float(res["obj"]), wall_s=round(wall, 3),
                      m=int(A.shape[0]), n=int(A.shape[1]), nnz=int(A.nnz),
                      x_min=float(np.min(x)), x_sum=float(np.sum(x)))


def run_25fv47():
    """Reference outcome on 25FV47: NaN at k=1."""
    from oracle import ref_harness as rh

    A, b, c, cTlb = save_problem("25FV47")
    res = rh.replay_interior_sparse(A, b, c, cTlb, tol=1e-8)
    return dict(k=int(res["k"]), obj=(None if not np.isfinite(res["obj"]) else float(res["obj"])))


def dense_goldens():
    from oracle import ref_harness as rh
    from oracle import ipm_oracle as orc

    ref_main, _ = rh.load_reference()
    out = {}
    for nm in ("ex1", "ex2", "ex3"):
        A, b, c = getattr(ref_main, nm)()
        r = rh.replay_interior_dense(A, b, c, tol=1e-8)
        out[nm] = dict(k=r["k"], obj=r["obj"], x=[float(v) for v in r["x"].ravel()])
    for (m, n, seeds) in ((64, 128, (0, 1)), (256, 512, (0, 1, 2, 3))):
        for seed in seeds:
            A, b, c = orc.synthetic_dense_lp(m, n, seed)
            t0 = time.time()
            r = rh.replay_interior_dense(A, b, c, tol=1e-8)
            out["synthetic_%dx%d_seed%d" % (m, n, seed)] = dict(k=r["k"], obj=r["obj"],
                                                                 wall_s=round(time.time() - t0, 3))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--skip-slow", action="store_true")
    ap.add_argument("--procs", type=int, default=6)
    ap.add_argument("--no-dense", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    names = args.only if args.only else CONVERGED
    if args.skip_slow:
        names = [n for n in names if n not in SLOW]

    res_path = os.path.join(GOLD, "reference_results.json")
    results = {}
    if os.path.exists(res_path):
        results = json.load(open(res_path))
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, tol=1e-8,
                semantics="interior_sparse replay, main.py:780-807")
    # longest first so the pool finishes sooner
    names = sorted(names, key=lambda n: (n not in SLOW, n))
    with mp.Pool(args.procs) as pool:
        for name, r in pool.imap_unordered(run_one, names):
            results[name] = r
            print(name, r, flush=True)
            json.dump(dict(meta=meta, **{k: v for k, v in results.items() if k != "meta"}),
                      open(res_path, "w"), indent=1, sort_keys=True)
    if not args.only:
        for nm in DATA_ONLY:
            save_problem(nm)
        results["25FV47"] = dict(reference_outcome=run_25fv47())
        json.dump(dict(meta=meta, **{k: v for k, v in results.items() if k != "meta"}),
                  open(res_path, "w"), indent=1, sort_keys=True)
    if not args.no_dense and not args.only:
        d = dense_goldens()
        d["meta"] = dict(numpy=np.__version__, scipy=scipy.__version__, tol=1e-8,
                         semantics="interior replay, main.py:718-751; generator SURVEY.md 8d")
        json.dump(d, open(os.path.join(GOLD, "dense_results.json"), "w"), indent=1, sort_keys=True)
        print("dense done", flush=True)


if __name__ == "__main__":
    main()
