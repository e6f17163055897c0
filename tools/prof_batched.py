"""One batched solve of B LPs 256x512 (for ncu captures): python tools/prof_batched.py B variant [max_iter]"""
import sys

import numpy as np
import torch

import interiorpointmethod_b200 as pkg
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
m, n = 256, 512
lib = _lib.load()
A, b, c = pkg.synthetic_dense_batch(0, min(B, 128), m, n)
reps = (B + A.shape[0] - 1) // A.shape[0]
A = np.tile(A, (reps, 1, 1))[:B]; b = np.tile(b, (reps, 1))[:B]; c = np.tile(c, (reps, 1))[:B]
dev = torch.device("cuda:0")
db = DeviceBatch(torch.from_numpy(A).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev))
lib.ipm_batched_set_variant(variant, _lib.REFRESH_DEFAULT)
nit = db.solve(tol=1e-8, max_iter=max_iter)
print("lockstep iterations", nit, "status ok", bool((db.status == 0).all()))
