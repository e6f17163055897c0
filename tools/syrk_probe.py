"""Times the batched SYRK and the fused batched Cholesky in isolation (back to back on one stream)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interiorpointmethod_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
for B in (1024, 4096, 8192):
    m, n = 256, 512
    A = torch.randn(B, m, n, dtype=torch.float64, device=dev)
    d = torch.rand(B, n, dtype=torch.float64, device=dev) + 0.1
    M = torch.empty(B, m, m, dtype=torch.float64, device=dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    torch.cuda.synchronize()
    for i in range(12):
        evs[i].record()
        lib.ipm_syrk_batched_d(0, B, m, n, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(M.data_ptr()), m)
    evs[12].record(); torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i+1]) for i in range(12)]
    print("syrk B=%d ms:" % B, " ".join("%.3f" % t for t in ts), " TF=%.1f" % (B*m*m*n/min(ts)*1e-9))
    nf = ctypes.c_int(0)
    ts = []
    for i in range(4):
        lib.ipm_syrk_batched_d(0, B, m, n, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(M.data_ptr()), m)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.ipm_potrf_batched_d(0, B, m, ctypes.c_void_p(M.data_ptr()), m, m*m, 1e-30, ctypes.byref(nf))
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("chol B=%d ms:" % B, " ".join("%.3f" % t for t in ts), "nfixed", nf.value)
    del A, d, M
print("dmma peak", lib.ipm_measure_dmma_peak(0))
