"""Newton iterations/s of the single-LP path on a few Netlib LPs (reference start, capped iterations)."""
import sys, time
import interiorpointmethod_b200 as ipm
names = sys.argv[1:] or ["AFIRO", "SCSD8", "25FV47", "TRUSS", "MAROS-R7", "QAP15", "STOCFOR3"]
for name in names:
    A, b, c, cT = ipm.load_golden_problem(name)
    with ipm.NewtonStep(A, b, c) as ns:
        cap = 40
        ns.solve(tol=1e-8, max_iter=10, cTlb=cT)
        t = time.perf_counter(); r = ns.solve(tol=1e-8, max_iter=cap, cTlb=cT); dt = time.perf_counter() - t
        print("%-9s m=%5d n=%5d  %3d iterations  %8.3f ms/iteration  %9.1f it/s  status %s obj %.10g" %
              (name, ns.m, ns.n, r.iterations, dt / r.iterations * 1e3, r.iterations / dt, r.status, r.objective), flush=True)
