"""Problem ingestion time (ipm_create + ipm_load_csc) against solve time: pattern built on the device (cold),
found in the per-structure cache (warm), and built by the host routine the device pass replaced."""
import sys, time
import numpy as np
import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib

L = _lib.load()
A, b, c, cT = ipm.load_golden_problem("AFIRO")
ipm.NewtonStep(A, b, c).close()          # context creation, module load


def load_ms(A, b, c, host_symbolic, use_cache, reps=3):
    best, info = 1e30, None
    for _ in range(reps):
        L.ipm_set_ingest_mode(host_symbolic, use_cache)
        t = time.perf_counter()
        ns = ipm.NewtonStep(A, b, c)
        dt = (time.perf_counter() - t) * 1e3
        info = ns.pattern_info()
        ns.close()
        best = min(best, dt)
    return best, info


for name in sys.argv[1:] or ["AFIRO", "SCSD8", "25FV47", "TRUSS", "MAROS-R7", "QAP15", "STOCFOR3"]:
    A, b, c, cT = ipm.load_golden_problem(name)
    L.ipm_pattern_cache_clear()
    host, ih = load_ms(A, b, c, 1, 0)
    cold, ic = load_ms(A, b, c, 0, 0)
    L.ipm_pattern_cache_clear()
    load_ms(A, b, c, 0, 1, reps=1)
    warm, iw = load_ms(A, b, c, 0, 1)
    assert iw["cache_hit"] and ic["device_built"] and not ih["device_built"]
    L.ipm_set_ingest_mode(0, 1)
    ns = ipm.NewtonStep(A, b, c)
    t = time.perf_counter(); r = ns.solve(tol=1e-8, max_iter=40, cTlb=cT, start="mehrotra"); t_solve = (time.perf_counter() - t) * 1e3
    ns.close()
    print("%-9s m=%5d n=%5d nnz=%7d terms=%8d | load: host-symbolic %7.2f ms (pattern %6.2f) device %7.2f ms (pattern %6.2f) "
          "cached %6.2f ms | solve(%3d it) %8.2f ms" % (name, A.shape[0], A.shape[1], A.nnz, ic["terms"], host, ih["build_ms"],
                                                      cold, ic["build_ms"], warm, r.iterations, t_solve), flush=True)
