"""A/B timing of the batched solver variants (3-pass vs 6-pass iteration) with the library's phase profiler."""
import ctypes
import sys
import time

import numpy as np
import torch

import interiorpointmethod_b200 as pkg
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
variants = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 1, 0, 1]
m, n = 256, 512
lib = _lib.load()
A, b, c = pkg.synthetic_dense_batch(0, min(B, 256), m, n)
reps = (B + A.shape[0] - 1) // A.shape[0]
A = np.tile(A, (reps, 1, 1))[:B]; b = np.tile(b, (reps, 1))[:B]; c = np.tile(c, (reps, 1))[:B]
dev = torch.device("cuda:0")
db = DeviceBatch(torch.from_numpy(A).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev))
res = {}
for variant in variants:
    lib.ipm_batched_set_variant(variant, _lib.REFRESH_DEFAULT)
    db.solve(tol=1e-8, max_iter=max_iter)
    lib.ipm_profile_enable(1)
    t0 = time.perf_counter()
    nit = db.solve(tol=1e-8, max_iter=max_iter)
    dt = time.perf_counter() - t0
    ms = (ctypes.c_double * 4)(); calls = (ctypes.c_int64 * 4)(); lpi = ctypes.c_int64(0)
    lib.ipm_profile_read(ms, calls, ctypes.byref(lpi))
    lib.ipm_profile_enable(0)
    obj = db.obj.cpu().numpy(); it = db.iters.cpu().numpy(); st = db.status.cpu().numpy()
    res[variant] = (obj, it)
    print("variant %d: %.2f ms, %d lockstep its, %.0f LPs/s, phases resid %.2f syrk %.2f chol %.2f solve %.2f, "
          "lp-iterations %d, status ok %s, iters %d..%d" % (variant, dt * 1e3, nit, B / dt, ms[0], ms[1], ms[2], ms[3],
                                                           lpi.value, (st == 0).all(), it.min(), it.max()), flush=True)
if 0 in res and 1 in res:
    o0, k0 = res[0]; o1, k1 = res[1]
    print("max |dk| =", np.abs(k0.astype(int) - k1.astype(int)).max(), " max rel dobj =", np.abs((o0 - o1) / o0).max())
