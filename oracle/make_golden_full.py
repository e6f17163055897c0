"""TEST INFRASTRUCTURE — freezes golden data for the general-form front end from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_full.py

Writes under tests/golden/full/:
  <NAME>.npz            the general-form LP exactly as the reference's loader create_problem_from_mps_matlab
                        returns it (sparse_interior.py:294-315): f, Aineq (CSC), bineq, Aeq (CSC), beq, lb, ub
  standard_form.json    per problem: the optimum scipy's HiGHS finds on the frozen data, and what the reference's own get_Abc(options="no-bound") (main.py:895-965) and
                        add_bound_into_matrix (main.py:968-1060) do with it - the exception text if they raise,
                        else shape, nnz and checksums of (A, b, c) and the constant - plus the Netlib optimum
                        listed in main.py:1417-1616
Only LPs whose .mat file is at most MAX_BYTES are frozen (the set has to travel with the repo).
"""
from __future__ import annotations

import ast
import json
import os
import sys
import warnings

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "full")
MAX_BYTES = 60_000


def netlib_table():
    """name -> optimum from the two literal lists inside main.benchmark (main.py:1417-1616)."""
    src = open(os.path.join(ref_harness.REFERENCE_ROOT, "main.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "benchmark")
    lists = {}
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            if node.targets[0].id in ("obj_values", "name") and isinstance(node.value, ast.List):
                lists[node.targets[0].id] = ast.literal_eval(node.value)
    assert len(lists["obj_values"]) == len(lists["name"]), (len(lists["obj_values"]), len(lists["name"]))
    return dict(zip(lists["name"], lists["obj_values"]))


def checksum(A, b, c):
    A = sparse.csr_matrix(A, dtype=np.float64)
    m, n = A.shape
    u = np.cos(np.arange(n, dtype=np.float64))[:, None]
    v = np.sin(np.arange(m, dtype=np.float64))[:, None]
    return {"shape": [int(m), int(n)], "nnz": int(A.nnz), "A_u": float(np.abs(A @ u).sum()),
            "At_v": float(np.abs(A.T @ v).sum()), "b_sum": float(np.sum(b)), "b_abs": float(np.abs(b).sum()),
            "c_sum": float(np.sum(c)), "c_abs": float(np.abs(c).sum())}


def main():
    ref_main, ref_sparse = ref_harness.load_reference()
    os.makedirs(OUT, exist_ok=True)
    table = netlib_table()
    names = sorted(f[:-4] for f in os.listdir(os.path.join(ref_harness.REFERENCE_ROOT, "benchmarks_full"))
                   if f.endswith(".mat"))
    meta = {}
    for name in names:
        size = os.path.getsize(os.path.join(ref_harness.REFERENCE_ROOT, "benchmarks_full", name + ".mat"))
        if size > MAX_BYTES:
            continue
        with ref_harness.in_reference_cwd(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c, Aineq, bineq, Aeq, beq, lb, ub = ref_sparse.create_problem_from_mps_matlab(name)
        n = len(c)

        def pack(M, prefix, out):
            if M is None:
                out[prefix + "_shape"] = np.array([0, n]); out[prefix + "_data"] = np.zeros(0)
                out[prefix + "_indices"] = np.zeros(0, np.int32); out[prefix + "_indptr"] = np.zeros(1, np.int32)
                return
            M = sparse.csc_matrix(M, dtype=np.float64)
            out[prefix + "_shape"] = np.array(M.shape); out[prefix + "_data"] = M.data
            out[prefix + "_indices"] = M.indices.astype(np.int32); out[prefix + "_indptr"] = M.indptr.astype(np.int32)

        arrays = {"f": np.asarray(c, np.float64).ravel(), "lb": np.asarray(lb, np.float64).ravel(),
                  "ub": np.asarray(ub, np.float64).ravel(),
                  "bineq": np.zeros(0) if bineq is None else np.asarray(bineq, np.float64).ravel(),
                  "beq": np.zeros(0) if beq is None else np.asarray(beq, np.float64).ravel()}
        pack(Aineq, "Aineq", arrays)
        pack(Aeq, "Aeq", arrays)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        # independent optimum of the frozen data (scipy HiGHS): main.py's table disagrees with the data for two LPs
        from scipy.optimize import linprog
        bnds = [(None if np.isneginf(l) else float(l), None if np.isposinf(u) else float(u))
                for l, u in zip(np.asarray(lb, float).ravel(), np.asarray(ub, float).ravel())]
        hr = linprog(np.asarray(c, float).ravel(), A_ub=Aineq, b_ub=None if bineq is None else np.asarray(bineq, float).ravel(),
                     A_eq=Aeq, b_eq=None if beq is None else np.asarray(beq, float).ravel(), bounds=bnds, method="highs")
        # what the reference's own front end does with it
        entry = {"netlib_optimum": table.get(name), "highs_optimum": float(hr.fun) if hr.status == 0 else None,
                 "n": int(n),
                 "m_ineq": 0 if Aineq is None else int(Aineq.shape[0]), "m_eq": 0 if Aeq is None else int(Aeq.shape[0]),
                 "lb_nonzero": int(np.count_nonzero(lb)), "lb_neg_inf": int(np.sum(np.isneginf(np.asarray(lb, float)))),
                 "ub_finite": int(np.sum(np.isfinite(np.asarray(ub, float))))}
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                A, b, cc, bound = ref_main.get_Abc(c=c, Aeq=Aeq, beq=beq, Aineq=Aineq, bineq=bineq, lb=lb, ub=ub,
                                                   options="no-bound")
            entry["get_Abc"] = checksum(A, b, cc)
            entry["get_Abc"]["bound"] = ("none" if bound is None else
                                         "lb=%s,ub=%s" % ("None" if bound[0] is None else "set",
                                                          "None" if bound[1] is None else "set"))
            if bound is not None:
                try:
                    with warnings.catch_warnings(), ref_harness.quiet():
                        warnings.simplefilter("ignore")
                        A2, b2, c2, bound2, const = ref_main.add_bound_into_matrix(A, b, cc, bound)
                    entry["add_bound"] = checksum(A2, b2, c2)
                    entry["add_bound"]["constant"] = float(np.asarray(const).ravel()[0])
                except BaseException as e:      # the reference raises str literals -> TypeError
                    entry["add_bound"] = {"raises": "%s: %s" % (type(e).__name__, str(e)[:120])}
        except BaseException as e:
            entry["get_Abc"] = {"raises": "%s: %s" % (type(e).__name__, str(e)[:120])}
        meta[name] = entry
        print(name, size, entry.get("get_Abc", {}).get("bound"), entry.get("add_bound", {}).get("raises", "ok"), flush=True)
    json.dump(meta, open(os.path.join(OUT, "standard_form.json"), "w"), indent=1, sort_keys=True)
    print(len(meta), "problems frozen")


if __name__ == "__main__":
    main()
