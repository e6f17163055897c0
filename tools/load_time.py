"""Time of problem upload (ipm_create + ipm_load_csr: CSR/CSR^T upload, symbolic SpGEMM on the host) vs solve."""
import sys, time
import interiorpointmethod_b200 as ipm
for name in sys.argv[1:] or ["AFIRO", "SCSD8", "25FV47", "TRUSS", "MAROS-R7", "QAP15", "STOCFOR3"]:
    A, b, c, cT = ipm.load_golden_problem(name)
    ipm.NewtonStep(A, b, c).close()
    t = time.perf_counter(); ns = ipm.NewtonStep(A, b, c); t_load = time.perf_counter() - t
    t = time.perf_counter(); r = ns.solve(tol=1e-8, max_iter=40, cTlb=cT, start="mehrotra"); t_solve = time.perf_counter() - t
    ns.close()
    print("%-9s m=%5d n=%5d nnz=%7d  load %8.2f ms   solve(%3d it) %8.2f ms" % (name, ns.m, ns.n, A.nnz, t_load * 1e3, r.iterations, t_solve * 1e3), flush=True)
