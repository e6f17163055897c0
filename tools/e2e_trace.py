"""Per-iteration phase times of ONE end-to-end solve (host buffers, chunks joining as they land) beside the
device-resident solve: where the e2e arm loses its time.  python tools/e2e_trace.py"""
import ctypes
import time

import torch

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch, solve_batched_pinned

lib = _lib.load()
B, m, n = 8192, 256, 512
A_h = torch.empty((B, m, n), dtype=torch.float64, pin_memory=True)
b_h = torch.empty((B, m), dtype=torch.float64, pin_memory=True)
c_h = torch.empty((B, n), dtype=torch.float64, pin_memory=True)
ipm.synthetic_dense_batch(0, B, m, n, out_A=A_h.numpy(), out_b=b_h.numpy(), out_c=c_h.numpy(), threads=16)
obj = torch.empty(B, dtype=torch.float64, pin_memory=True)
it = torch.empty(B, dtype=torch.int32, pin_memory=True)
st = torch.empty(B, dtype=torch.int32, pin_memory=True)
time.sleep(1.0)


def trace(label, fn):
    fn()
    lib.ipm_profile_enable(1)
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = (ctypes.c_double * 1024)(); ph = (ctypes.c_int * 1024)()
    k = lib.ipm_profile_last(ms, ph, 1024)
    lib.ipm_profile_enable(0)
    rows, cur = [], [0.0, 0.0, 0.0, 0.0]
    for i in range(k):
        cur[ph[i]] += ms[i]
        if ph[i] == 3:
            rows.append(cur); cur = [0.0, 0.0, 0.0, 0.0]
    if any(cur):
        rows.append(cur)
    tot = [sum(r[j] for r in rows) for j in range(4)]
    print("%s: wall %.1f ms, %d iterations, phase sums resid %.1f syrk %.1f chol %.1f solves %.1f = %.1f ms" %
          (label, wall, len(rows), *tot, sum(tot)))
    print("  per iteration (resid, syrk, chol, solves): " + " | ".join("%.1f %.1f %.1f %.1f" % tuple(r) for r in rows))


dev = torch.device("cuda:0")
db = DeviceBatch(A_h.to(dev), b_h.to(dev), c_h.to(dev))
trace("device-resident", lambda: db.solve(tol=1e-8))
del db
trace("end to end     ", lambda: solve_batched_pinned(A_h, b_h, c_h, obj, it, st, tol=1e-8, device=0))
# raw copy rate of the same buffers
d = torch.empty_like(A_h, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(A_h, non_blocking=True); torch.cuda.synchronize()
print("H2D of A alone: %.1f ms = %.1f GB/s" % ((time.perf_counter() - t0) * 1e3, A_h.numel() * 8 / (time.perf_counter() - t0) * 1e-9))
