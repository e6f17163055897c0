"""Phase profile of the augmented-system kernel (ipm_solve_dense_kkt) on one generator LP: python tools/kkt_profile.py [m n seed]"""
import ctypes
import sys
import time

import numpy as np

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib

lib = _lib.load()
m, n, seed = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 512, 0)
A, b, c = ipm.synthetic_dense_lp(m, n, seed)
names = ["residuals", "build K", "panel load", "panel column steps", "panel store", "row swaps", "U12+update+sync", "solves",
         "elementwise"]
for cl in (1, 4, 8):
    lib.ipm_set_kkt_cluster(cl)
    ipm.interior_kkt(A, b, c, tol=1e-8)
    t0 = time.perf_counter()
    r = ipm.interior_kkt(A, b, c, tol=1e-8)
    dt = time.perf_counter() - t0
    prof = (ctypes.c_int64 * 16)()
    lib.ipm_kkt_last_profile(prof)
    its = max(1, prof[9])
    tot = sum(prof[i] for i in range(9))
    print("cluster %d: %d iterations, %.1f ms wall (incl. copies), %.2f ms per iteration in-kernel (1965 MHz)" % (cl, r.iterations, dt * 1e3, tot / its / 1.965e6))
    print("   " + "  ".join("%s %.0f us" % (names[i], prof[i] / its / 1965.0) for i in range(9)))
lib.ipm_set_kkt_cluster(4)
