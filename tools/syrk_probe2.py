import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interiorpointmethod_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B, m, n = 1024, 256, 512
A = torch.randn(B, m, n, dtype=torch.float64, device=dev)
d = torch.rand(B, n, dtype=torch.float64, device=dev) + 0.1
M = torch.empty(B, m, m, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for i in range(6):
    lib.ipm_syrk_batched_d(0, B, m, n, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(M.data_ptr()), m)
torch.cuda.synchronize()
