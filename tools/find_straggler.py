"""Finds the LP(s) of the weak-scaling workload (seeds 8192..65535) that keep the lockstep loop alive (3527
iterations in profiles/r1_bench_n8_weak_nvidia_smi_sampler.json) and re-solves their block with the alternatives."""
import sys, time
import numpy as np, torch
import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch
lib = _lib.load()
m, n, BLK = 256, 512, 8192
dev = torch.device("cuda:0")
A_h = torch.empty((BLK, m, n), dtype=torch.float64, pin_memory=True); b_h = torch.empty((BLK, m), dtype=torch.float64, pin_memory=True)
c_h = torch.empty((BLK, n), dtype=torch.float64, pin_memory=True)
blocks = [int(v) for v in sys.argv[1:] if not v.startswith('--')] or list(range(1, 8))
t_start = time.time()
for blk in blocks:
    ipm.synthetic_dense_batch(blk * BLK, BLK, m, n, out_A=A_h.numpy(), out_b=b_h.numpy(), out_c=c_h.numpy(), threads=16)
    db = DeviceBatch(A_h.to(dev), b_h.to(dev), c_h.to(dev))
    lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT); lib.ipm_batched_set_straggler_restart(0 if '--no-restart' in sys.argv else 8)
    t = time.perf_counter(); nit = db.solve(tol=1e-8); dt = time.perf_counter() - t
    it = db.iters.cpu().numpy()
    bad = np.nonzero(it > 30)[0]
    print("block %d (seeds %d..): lockstep %d, %.0f ms, outliers %s (t=%.0fs)" % (blk, blk * BLK, nit, dt * 1e3,
          [(int(blk * BLK + i), int(it[i])) for i in bad[:8]], time.time() - t_start), flush=True)
    if len(bad) and '--alts' in sys.argv:
        for label, v3, rf, slack in (("four-pass, restart slack 8", 1, 3, 8), ("four-pass, refresh every 1", 1, 1, 0), ("six-pass", 0, 3, 0)):
            lib.ipm_batched_set_variant(v3, rf); lib.ipm_batched_set_straggler_restart(slack)
            t = time.perf_counter(); nit = db.solve(tol=1e-8); dt = time.perf_counter() - t
            it2 = db.iters.cpu().numpy(); st2 = db.status.cpu().numpy(); ob = db.obj.cpu().numpy()
            print("   %-28s lockstep %4d  %.0f ms  outlier its %s status %s obj %s  max|dk| others %d" % (
                label, nit, dt * 1e3, [int(it2[i]) for i in bad[:8]], [int(st2[i]) for i in bad[:8]],
                ["%.10g" % ob[i] for i in bad[:4]], int(np.abs(np.delete(it2, bad).astype(int) - np.delete(it, bad).astype(int)).max())), flush=True)
        lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT); lib.ipm_batched_set_straggler_restart(8)
    del db
