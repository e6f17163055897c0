"""Batched SYRK rate by tile mix: m = 128 (diagonal tiles only), 256 (2 diagonal + 1 full), 512, 1024 (mostly full tiles).
Executed DMMA flops are counted (diagonal tile = 17/32 of a full one), so the figures compare pipe utilisation."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interiorpointmethod_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
n = 512
for m, B in ((128, 16384), (256, 8192), (512, 2048), (1024, 512)):
    A = torch.randn(B, m, n, dtype=torch.float64, device=dev)
    d = torch.rand(B, n, dtype=torch.float64, device=dev) + 0.1
    M = torch.empty(B, m, m, dtype=torch.float64, device=dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    torch.cuda.synchronize()
    for i in range(6):
        evs[i].record()
        lib.ipm_syrk_batched_d(0, B, m, n, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(M.data_ptr()), m)
    evs[6].record(); torch.cuda.synchronize()
    t = min(evs[i].elapsed_time(evs[i + 1]) for i in range(1, 6))
    T = m // 128
    tiles_full, tiles_diag = T * (T - 1) // 2, T
    executed = B * (tiles_full + tiles_diag * 17.0 / 32.0) * 128 * 128 * n * 2
    print("m=%4d B=%5d  %.3f ms  symmetric-count %.1f TF  executed-DMMA %.1f TF (diag tiles %d, full %d per matrix)" %
          (m, B, t, B * m * m * n / t * 1e-9, executed / t * 1e-9, tiles_diag, tiles_full), flush=True)
    del A, d, M
print("dmma peak", lib.ipm_measure_dmma_peak(0))
