"""SYRK + blocked Cholesky once at the dense-big shape (for ncu captures): python tools/dense_big_kernels.py [m] [n]"""
import ctypes
import sys

import torch

from interiorpointmethod_b200 import _lib

lib = _lib.load()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * m
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
d = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
M = torch.empty(m, m, dtype=torch.float64, device=dev)
nf = ctypes.c_int(0)
print(lib.ipm_syrk_d(0, m, n, ctypes.c_void_p(A.data_ptr()), n, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(M.data_ptr()), m),
      lib.ipm_potrf_d(0, m, ctypes.c_void_p(M.data_ptr()), m, 1e-30, ctypes.byref(nf)), nf.value)
torch.cuda.synchronize()
