"""Newton iterations/s of small sparse LPs: one-launch solve (ipm_set_small_lp_fused 1) against the CUDA-graph replay per
iteration (0).  python tools/small_lp_rate.py [names...]"""
import sys
import time

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib

lib = _lib.load()
names = sys.argv[1:] or ["AFIRO", "SC50A", "ADLITTLE", "SHARE2B", "SC205", "E226", "BOEING2"]
for name in names:
    A, b, c, cT = ipm.load_golden_problem(name)
    for fused in (0, 1):
        lib.ipm_set_small_lp_fused(fused)
        with ipm.NewtonStep(A, b, c) as ns:
            ns.solve(tol=1e-8, max_iter=400, cTlb=cT)
            best = 1e9
            for _ in range(3):
                t = time.perf_counter(); r = ns.solve(tol=1e-8, max_iter=400, cTlb=cT); best = min(best, time.perf_counter() - t)
            print("%-9s m=%4d n=%4d fused=%d  %3d iterations  %8.1f us/iteration  %9.0f it/s  %s obj %.12g" %
                  (name, ns.m, ns.n, fused, r.iterations, best / max(1, r.iterations) * 1e6, r.iterations / best, r.status, r.objective),
                  flush=True)
lib.ipm_set_small_lp_fused(1)
