"""The eight shards of the strong-scaling run (8192 generator LPs, 1024 per GPU) solved one after the other on ONE GPU:
which shard is the slowest rank of `config.strong`, and why (lockstep iterations, hand-offs).
python tools/strong_shards.py [shards=8] [reps=3]"""
import ctypes
import sys

import torch

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
B, m, n = 8192 // shards, 256, 512
dev = torch.device("cuda:0")
for r in range(shards):
    A, b, c = ipm.synthetic_dense_batch(r * B, B, m, n, threads=16)
    db = DeviceBatch(torch.from_numpy(A).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev))
    db.solve(tol=1e-8, max_iter=400)
    ts = []
    for _ in range(reps):
        lib.ipm_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        nit = db.solve(tol=1e-8, max_iter=400)
        e1.record()
        torch.cuda.synchronize()
        ms = (ctypes.c_double * 4)(); calls = (ctypes.c_int64 * 4)(); lpi = ctypes.c_int64(0)
        lib.ipm_profile_read(ms, calls, ctypes.byref(lpi))
        lib.ipm_profile_enable(0)
        ts.append(e0.elapsed_time(e1))
    it = db.iters.cpu().numpy()
    print("shard %d (seeds %5d..): %s ms  lockstep %d  k %d..%d  hand-offs %d  phases %s" % (
        r, r * B, " ".join("%.1f" % t for t in ts), nit, it.min(), it.max(), lib.ipm_batched_last_handoffs(),
        [round(v, 1) for v in ms]), flush=True)
    del db
