"""TEST INFRASTRUCTURE - freezes EVERY standard-form LP of the reference's benchmarks/ directory for the GPU box
and pins each to an independent optimum.  Build container only (needs /root/reference and scipy's HiGHS).

    python oracle/make_golden_netlib_all.py [workers]

For each benchmarks/<NAME>.mat (81 files; the reference's list: main.py:1317-1616):
  tests/golden/problems/<NAME>.npz   exactly what sparse_interior.create_problem_from_mps returns (files already
                                     frozen by make_golden.py are left alone)
  tests/golden/netlib_all.json       per LP: m, n, nnz, finite (b, c, cTlb free of NaN/Inf: 8 files are not, SURVEY
                                     App. C.2), netlib_optimum (the reference's own table, main.benchmark()),
                                     highs = {status, optimum (c^T x - cTlb), seconds} from
                                     scipy.optimize.linprog(method="highs") on the SAME standard-form data (the
                                     table is not valid where the file is not a faithful standard form: bounded
                                     LPs such as GROW*, KB2, FIT1P), reference = outcome of the unmodified reference
                                     where tests/golden/reference_results.json has it
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
HIGHS_LIMIT_S = 900.0


def work(name):
    import numpy as np
    from scipy import sparse
    from scipy.optimize import linprog

    from oracle import ref_harness as rh
    A, b, c, cTlb = rh.load_problem(name)
    A = sparse.csc_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    path = os.path.join(GOLD, "problems", name + ".npz")
    if not os.path.exists(path):
        np.savez_compressed(path, m=A.shape[0], n=A.shape[1], indptr=A.indptr.astype(np.int32),
                            indices=A.indices.astype(np.int32), data=A.data.astype(np.float64),
                            b=np.asarray(b, dtype=np.float64).ravel(), c=np.asarray(c, dtype=np.float64).ravel(),
                            cTlb=np.float64(cTlb))
    bb = np.asarray(b, dtype=np.float64).ravel()
    cc = np.asarray(c, dtype=np.float64).ravel()
    finite = bool(np.isfinite(bb).all() and np.isfinite(cc).all() and np.isfinite(float(cTlb)) and np.isfinite(A.data).all())
    ent = dict(m=int(A.shape[0]), n=int(A.shape[1]), nnz=int(A.nnz), finite=finite, cTlb=float(cTlb) if finite else None)
    if finite:
        t0 = time.time()
        try:
            r = linprog(cc, A_eq=sparse.csr_matrix(A, dtype=np.float64), b_eq=bb, bounds=(0, None), method="highs",
                        options=dict(time_limit=HIGHS_LIMIT_S, presolve=True))
            ent["highs"] = dict(status=int(r.status), message=str(r.message)[:80],
                                optimum=(float(r.fun) - float(cTlb)) if r.status == 0 else None,
                                seconds=round(time.time() - t0, 2))
        except Exception as e:  # pragma: no cover
            ent["highs"] = dict(status=-1, message=str(e)[:80], optimum=None, seconds=round(time.time() - t0, 2))
    return name, ent


def main():
    import multiprocessing as mp

    from oracle import ref_harness as rh
    ref_main, _ = rh.load_reference()
    names, vals = ref_main.benchmark()
    table = dict(zip(names, vals))
    files = sorted(f[:-4] for f in os.listdir(os.path.join(rh.REFERENCE_ROOT, "benchmarks")) if f.endswith(".mat"))
    refres = json.load(open(os.path.join(GOLD, "reference_results.json")))
    out = {}
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    with mp.Pool(workers) as pool:
        for name, ent in pool.imap_unordered(work, files):
            ent["netlib_optimum"] = table.get(name)
            if name in refres and isinstance(refres[name], dict) and "k" in refres[name]:
                ent["reference"] = dict(k=refres[name]["k"], obj=refres[name]["obj"])
            out[name] = ent
            print(name, ent, flush=True)
    import scipy
    out = dict(sorted(out.items()))
    json.dump(dict(meta=dict(scipy=scipy.__version__, highs_time_limit_s=HIGHS_LIMIT_S,
                             note="optimum = c^T x - cTlb of the standard-form data in the file"), problems=out),
              open(os.path.join(GOLD, "netlib_all.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
