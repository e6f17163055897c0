"""Times ipm_potrf_d / ipm_syrk_d on one large dense matrix (dense-big shape by default)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interiorpointmethod_b200 import _lib
lib = _lib.load()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * m
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
d = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
M = torch.empty(m, m, dtype=torch.float64, device=dev)
nf = ctypes.c_int(0)
for rep in range(reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    lib.ipm_syrk_d(0, m, n, ctypes.c_void_p(A.data_ptr()), n, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(M.data_ptr()), m)
    e[1].record()
    lib.ipm_potrf_d(0, m, ctypes.c_void_p(M.data_ptr()), m, 1e-30, ctypes.byref(nf))
    e[2].record()
    torch.cuda.synchronize()
    ts, tc = e[0].elapsed_time(e[1]) * 1e-3, e[1].elapsed_time(e[2]) * 1e-3
    print("m=%d n=%d syrk %.4f s (%.1f TF)  potrf %.4f s (%.1f TF) nfixed %d" % (m, n, ts, m*m*n/ts*1e-12, tc, m**3/3/tc*1e-12, nf.value), flush=True)
