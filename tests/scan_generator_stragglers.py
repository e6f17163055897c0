"""CPU scan of the benchmark generator's LPs (256 x 512, seed = LP index) for stragglers - TEST INFRASTRUCTURE
(uses the oracle; not collected by pytest).  Results of the round-1 run: profiles/r1_cpu_straggler_scan.txt.

    python tests/scan_generator_stragglers.py FIRST LAST literal|refined[:THRESH]|always [workers]

literal: the oracle's normal-equations iteration (what the GPU's six-pass iteration restates)
refined: the same with one refinement step of the corrector when |(-rb - A dx)| > THRESH |rb| (default 0.1)
always : the same with the refinement step in every iteration (what restarted LPs run on the GPU)
"""
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MODE = sys.argv[3] if len(sys.argv) > 3 else "literal"
CAP = 150


def _refined(O, A, b, c, thresh):
    import numpy as np
    m, n = A.shape
    b = b.reshape(-1, 1)
    c = c.reshape(-1, 1)
    x, y, s = O.initial_point(m, n, y0_is_one=False)
    k = 0
    while O.continue_flag(A, b, c, x, y, s, 1e-8, 1e-8, 1e-8) and k < CAP:
        rb, rc = O.residuals(A, b, c, x, y, s)
        r3 = x * s
        L, _ = O.cholesky_safeguarded(O.normal_matrix(A, x, s))
        dxa, dya, dsa = O.direction_normal(A, L, x, s, rb, rc, r3)
        _, mu, sigma = O.sigma_mu(x, s, dxa, dsa)
        r4 = r3 + dxa * dsa - sigma * mu
        dx, dy, ds = O.direction_normal(A, L, x, s, rb, rc, r4)
        delta = -rb - A @ dx
        if np.linalg.norm(delta) > thresh * np.linalg.norm(rb):
            ddy = O.solve_with_factor(L, delta)
            dy = dy + ddy
            dx = dx + (x / s) * (A.T @ ddy)
            ds = (-s * dx / x) - (r4 / x)
        ap, ad = O.full_stepsize(x, s, dx, ds)
        x, y, s = x + ap * dx, y + ad * dy, s + ad * ds
        k += 1
    return k


def work(seed):
    from oracle import ipm_oracle as O
    A, b, c = O.synthetic_dense_lp(256, 512, seed)
    if MODE == "literal":
        return seed, O.solve(A, b, c, tol=1e-8, max_iter=CAP, y0_is_one=False, linear="normal")["k"]
    if MODE.startswith("refined"):
        return seed, _refined(O, A, b, c, float(MODE.split(":")[1]) if ":" in MODE else 0.1)
    return seed, _refined(O, A, b, c, 0.0)


if __name__ == "__main__":
    import multiprocessing as mp
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    workers = int(sys.argv[4]) if len(sys.argv) > 4 else max(1, (os.cpu_count() or 2) - 1)
    hist, t0 = {}, time.time()
    with mp.Pool(workers) as pool:
        for seed, k in pool.imap_unordered(work, range(lo, hi), chunksize=16):
            hist[k] = hist.get(k, 0) + 1
            if k > 21:
                print("OUTLIER seed %d k %d%s" % (seed, k, " (cap)" if k >= CAP else ""), flush=True)
    print("%s, seeds %d..%d: iterations -> LPs %s (%.0f s, %d workers)" % (MODE, lo, hi - 1, sorted(hist.items()),
                                                                         time.time() - t0, workers))
