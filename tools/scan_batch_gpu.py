"""GPU scan of the benchmark generator (256 x 512, LP i = default_rng(i)) against the frozen oracle table
tests/golden/batch_256x512_oracle.npz, in blocks of 8192 LPs:  python tools/scan_batch_gpu.py [block ...] [--ab]

Per block: lockstep iterations, time per solve, LPs not converged, LPs whose iteration count differs from the table
by more than 1, worst relative objective difference, refinements taken.  --ab also times the alternatives on the
first block (strip source tensor map vs strip-major copy, refinement on/off, six-pass)."""
import os
import sys
import time

import numpy as np
import torch

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = _lib.load()
m, n, BLK = 256, 512, 8192
tab = np.load(os.path.join(ROOT, "tests", "golden", "batch_256x512_oracle.npz"))
tk, tobj = tab["k"].astype(int), tab["obj"]
dev = torch.device("cuda:0")
A_h = torch.empty((BLK, m, n), dtype=torch.float64, pin_memory=True)
b_h = torch.empty((BLK, m), dtype=torch.float64, pin_memory=True)
c_h = torch.empty((BLK, n), dtype=torch.float64, pin_memory=True)
blocks = [int(v) for v in sys.argv[1:] if not v.startswith("--")] or list(range(8))
for v in sys.argv[1:]:
    if v.startswith("--refresh="):          # period of the from-scratch residual check of the four-pass iteration (default 3)
        lib.ipm_batched_set_variant(1, int(v.split("=")[1]))
t_start = time.time()
tot_bad = 0


def timed_solve(db, reps=1):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        nit = db.solve(tol=1e-8, max_iter=400)
    torch.cuda.synchronize()
    return nit, (time.perf_counter() - t) / reps


for blk in blocks:
    ipm.synthetic_dense_batch(blk * BLK, BLK, m, n, out_A=A_h.numpy(), out_b=b_h.numpy(), out_c=c_h.numpy(), threads=16)
    db = DeviceBatch(A_h.to(dev), b_h.to(dev), c_h.to(dev))
    nit, dt = timed_solve(db)
    nho = lib.ipm_batched_last_handoffs()
    it, st, ob = db.iters.cpu().numpy().astype(int), db.status.cpu().numpy(), db.obj.cpu().numpy()
    kk, oo = tk[blk * BLK:(blk + 1) * BLK], tobj[blk * BLK:(blk + 1) * BLK]
    dk = np.abs(it - kk)
    rel = np.abs(ob - oo) / np.maximum(1.0, np.abs(oo))
    bad = np.nonzero((st != 0) | (dk > 1) | ~(rel <= 1e-8))[0]
    tot_bad += bad.size
    print("block %d (seeds %d..%d): lockstep %d, %.1f ms, not converged %d, |dk|>1: %d, max rel dobj %.2e, dk histogram %s, "
          "iterations %s, handed off %d  (t=%.0fs)" % (blk, blk * BLK, (blk + 1) * BLK - 1, nit, dt * 1e3, int((st != 0).sum()),
                                        int((dk > 1).sum()), float(np.nanmax(rel)),
                                        dict(zip(*np.unique(it - kk, return_counts=True))),
                                        dict(zip(*np.unique(it, return_counts=True))), nho, time.time() - t_start), flush=True)
    for i in bad[:12]:
        print("   seed %d: status %d, k %d (table %d), obj %.12g (table %.12g)" % (blk * BLK + i, st[i], it[i], kk[i], ob[i], oo[i]))
    if "--ab" in sys.argv and blk == blocks[0]:
        base = (it.copy(), ob.copy())
        for label, setup in (("default (tensor-map strips, refinement on)", []),
                             ("strip-major copy instead of the tensor map", [(_lib.BOPT_STRIP_TMA, 0)]),
                             ("refinement off", [(_lib.BOPT_REFINE, 0)]),
                             ("hand-off off", [(_lib.BOPT_HANDOFF, 0)]),
                             ("six-pass literal iteration", "six")):
            if setup == "six":
                lib.ipm_batched_set_variant(0, 3)
            else:
                for k_, v_ in setup:
                    lib.ipm_batched_set_option(k_, v_)
            db2 = DeviceBatch(db.A, db.b, db.c)         # workspace sized for the option in force
            timed_solve(db2)
            nit2, dt2 = timed_solve(db2, 3)
            it2, ob2, st2 = db2.iters.cpu().numpy().astype(int), db2.obj.cpu().numpy(), db2.status.cpu().numpy()
            print("   A/B %-46s %.1f ms per solve, lockstep %d, converged %d, max|dk vs default| %d, bitwise equal %s"
                  % (label, dt2 * 1e3, nit2, int((st2 == 0).sum()), int(np.abs(it2 - base[0]).max()),
                     bool(np.array_equal(ob2, base[1]))), flush=True)
            del db2
            lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
            lib.ipm_batched_set_option(_lib.BOPT_STRIP_TMA, 1)
            lib.ipm_batched_set_option(_lib.BOPT_REFINE, 1)
            lib.ipm_batched_set_option(_lib.BOPT_HANDOFF, 1)
    del db
print("SCAN %s: %d LPs outside the parity bar over %d blocks" % ("OK" if tot_bad == 0 else "FAILED", tot_bad, len(blocks)))
sys.exit(0 if tot_bad == 0 else 1)
