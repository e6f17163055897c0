"""Finds LPs of the benchmark batch whose iteration count differs between the two batched variants."""
import sys
import numpy as np
import interiorpointmethod_b200 as pkg
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import solve_batched_host

lib = _lib.load()
m, n = 256, 512
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
for first in range(lo, hi, 1024):
    A, b, c = pkg.synthetic_dense_batch(first, 1024, m, n)
    lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
    o1, k1, s1 = solve_batched_host(A, b, c, tol=1e-8)
    lib.ipm_batched_set_variant(0, 3)
    o0, k0, s0 = solve_batched_host(A, b, c, tol=1e-8)
    d = np.abs(k1.astype(int) - k0.astype(int))
    idx = np.nonzero(d > 1)[0]
    print(first, "max k0", k0.max(), "max k1", k1.max(), "n diff>1:", len(idx), flush=True)
    for i in idx:
        print("  LP", first + i, "k0", k0[i], "k1", k1[i], "obj0 %.12g obj1 %.12g" % (o0[i], o1[i]), "status", s0[i], s1[i])
        # iteration history of the 3-pass variant through max_iter caps
        for var in (0, 1):
            lib.ipm_batched_set_variant(var, _lib.REFRESH_DEFAULT)
            hist = []
            for cap in list(range(10, 26)):
                oo, kk, ss, xx = solve_batched_host(A[i:i+1], b[i:i+1], c[i:i+1], tol=1e-8, max_iter=cap, want_x=True)
                rb = np.linalg.norm(A[i] @ xx[0] - b[i])
                hist.append("%d:%d:%.2e:%.3e" % (cap, kk[0], rb, xx[0].min()))
            print("   var", var, " ".join(hist))
        bad.append(first + i)
lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
print("bad", bad)
