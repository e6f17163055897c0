"""Interleaved A/B timing of the batched solver's options on the benchmark batch (8192 generator LPs 256 x 512, seeds
FIRST..): python tools/batched_variants.py [first_seed] [reps] [B]   -  CUDA-event time per solve, phase breakdown."""
import ctypes
import sys

import numpy as np
import torch

import interiorpointmethod_b200 as ipm
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
m, n = 256, 512
dev = torch.device("cuda:0")
A_h = torch.empty((B, m, n), dtype=torch.float64, pin_memory=True)
b_h = torch.empty((B, m), dtype=torch.float64)
c_h = torch.empty((B, n), dtype=torch.float64)
ipm.synthetic_dense_batch(first, B, m, n, out_A=A_h.numpy(), out_b=b_h.numpy(), out_c=c_h.numpy(), threads=16)
A_d, b_d, c_d = A_h.to(dev), b_h.to(dev), c_h.to(dev)
VARIANTS = [("default", []), ("rhs pass of its own", [(_lib.BOPT_SYRK_RHS, 0)]), ("refinement off", [(_lib.BOPT_REFINE, 0)]), ("hand-off off", [(_lib.BOPT_HANDOFF, 0)]),
            ("strip-major copy", [(_lib.BOPT_STRIP_TMA, 0)]), ("six-pass", "six"), ("syrk stages 32x3", "bk32"),
            ("syrk 16 consumer warps", "c16")]


def setup(v):
    lib.ipm_batched_set_variant(1, _lib.REFRESH_DEFAULT)
    for k in (_lib.BOPT_REFINE, _lib.BOPT_HANDOFF, _lib.BOPT_STRIP_TMA, _lib.BOPT_SYRK_RHS):
        lib.ipm_batched_set_option(k, 1)
    lib.ipm_set_syrk_stage_width(16)
    lib.ipm_set_syrk_consumers(8)
    if v == "c16":
        lib.ipm_set_syrk_consumers(16)
    elif v == "six":
        lib.ipm_batched_set_variant(0, 3)
    elif v == "bk32":
        lib.ipm_set_syrk_stage_width(32)
    else:
        for k, val in v:
            lib.ipm_batched_set_option(k, val)


times = {name: [] for name, _ in VARIANTS}
info = {}
for rep in range(reps + 1):
    for name, v in VARIANTS:
        setup(v)
        db = DeviceBatch(A_d, b_d, c_d)
        if rep == 0:
            db.solve(tol=1e-8, max_iter=400)            # warm-up (attributes, allocations)
            continue
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
        times[name].append(e0.elapsed_time(e1))
        it = db.iters.cpu().numpy()
        info[name] = dict(lockstep=nit, converged=int((db.status == 0).sum()), handoffs=lib.ipm_batched_last_handoffs(),
                          phases=[round(v, 1) for v in ms], newton=int(it.sum()))
        del db
setup([])
for name, _ in VARIANTS:
    t = times[name]
    print("%-18s %s ms (min %.1f)  %s" % (name, " ".join("%.1f" % v for v in t), min(t), info[name]), flush=True)
