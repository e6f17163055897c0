"""A/B of the blocked Cholesky's diagonal-block kernel (ipm_set_chol_fused_diag) on mid-size and large single LPs:
Newton iterations/s and results.  python tools/chol_panel_ab.py [names...]"""
import sys
import time

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib

lib = _lib.load()
names = sys.argv[1:] or ["QAP15", "MAROS-R7", "STOCFOR2", "SCTAP3", "WOODW", "STOCFOR3"]
for name in names:
    A, b, c, cT = ipm.load_golden_problem(name)
    for fused in (0, 1, 0, 1):
        lib.ipm_set_chol_fused_diag(fused)
        with ipm.NewtonStep(A, b, c) as ns:
            ns.solve(tol=1e-8, max_iter=5, cTlb=cT)
            t = time.perf_counter(); r = ns.solve(tol=1e-8, max_iter=100, cTlb=cT); dt = time.perf_counter() - t
            print("%-9s m=%5d n=%5d fused_diag=%d  %3d iterations  %8.3f ms/iteration  %9.1f it/s  status %s obj %.12g" %
                  (name, ns.m, ns.n, fused, r.iterations, dt / r.iterations * 1e3, r.iterations / dt, r.status, r.objective),
                  flush=True)
lib.ipm_set_chol_fused_diag(1)
