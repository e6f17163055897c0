"""Device-resident solve time of each 8192-LP block of the benchmark generator on ONE GPU (what each rank of the
8-GPU weak run solves): wall per solve, sum of the lockstep phases, LPs handed to the augmented-system kernel.
The difference wall - phases is what runs after the lockstep loop (late hand-offs).  python tools/block_times.py [block ...]"""
import ctypes
import sys
import time

import torch

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch

lib = _lib.load()
m, n, BLK = 256, 512, 8192
dev = torch.device("cuda:0")
A_h = torch.empty((BLK, m, n), dtype=torch.float64, pin_memory=True)
b_h = torch.empty((BLK, m), dtype=torch.float64, pin_memory=True)
c_h = torch.empty((BLK, n), dtype=torch.float64, pin_memory=True)
for blk in [int(v) for v in sys.argv[1:]] or list(range(8)):
    ipm.synthetic_dense_batch(blk * BLK, BLK, m, n, out_A=A_h.numpy(), out_b=b_h.numpy(), out_c=c_h.numpy(), threads=16)
    db = DeviceBatch(A_h.to(dev), b_h.to(dev), c_h.to(dev))
    db.solve(tol=1e-8)
    torch.cuda.synchronize()
    walls = []
    for rep in range(2):
        t0 = time.perf_counter()
        nit = db.solve(tol=1e-8)
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1e3)
    lib.ipm_profile_enable(1)
    t0 = time.perf_counter()
    db.solve(tol=1e-8)
    torch.cuda.synchronize()
    wp = (time.perf_counter() - t0) * 1e3
    ms = (ctypes.c_double * 1024)(); ph = (ctypes.c_int * 1024)()
    k = lib.ipm_profile_last(ms, ph, 1024)
    lib.ipm_profile_enable(0)
    print("block %d: wall %.1f %.1f ms, profiled solve wall %.1f ms = phases %.1f + %.1f, lockstep %d, handed off %d" %
          (blk, walls[0], walls[1], wp, sum(ms[i] for i in range(k)), wp - sum(ms[i] for i in range(k)), nit,
           lib.ipm_batched_last_handoffs()), flush=True)
    del db
