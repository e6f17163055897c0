"""Solves one golden Netlib LP with a cap on the iterations (for launch-list profiling)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interiorpointmethod_b200 as ipm
name = sys.argv[1]; cap = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
A, b, c, cTlb = ipm.load_golden_problem(name)
with ipm.NewtonStep(A, b, c) as ns:
    t0 = time.perf_counter(); r = ns.solve(tol=1e-8, max_iter=cap, cTlb=cTlb); dt = time.perf_counter() - t0
print(name, ns.m, ns.n, "k", r.iterations, r.status, "obj", r.objective, "%.4fs %.1f it/s" % (dt, r.iterations / dt))
