"""Where the time of one small solve goes outside the iterations: create / load / first solve (graph capture) /
second solve / destroy."""
import ctypes, sys, time
import numpy as np
import interiorpointmethod_b200 as ipm

def T(f, *a, **k):
    t = time.perf_counter(); r = f(*a, **k); return (time.perf_counter() - t) * 1e3, r

for name in sys.argv[1:] or ["AFIRO", "SC205", "SCSD8", "25FV47"]:
    A, b, c, cT = ipm.load_golden_problem(name)
    ipm.NewtonStep(A, b, c).close()
    for rep in range(3):
        t_new, ns = T(ipm.NewtonStep, A, b, c)
        t_s1, r1 = T(ns.solve, tol=1e-8, cTlb=cT)
        t_s2, r2 = T(ns.solve, tol=1e-8, cTlb=cT)
        t_close, _ = T(ns.close)
        t_all, r3 = T(ipm.interior_sparse, A, b, c, cT, 1e-8)
        print("%-8s rep %d: create+load %6.2f ms (load %5.2f)  solve#1 %7.2f ms  solve#2 %7.2f ms (%d it)  close %5.2f ms | interior_sparse() %7.2f ms"
              % (name, rep, t_new, 0.0, t_s1, t_s2, r2.iterations, t_close, t_all), flush=True)
