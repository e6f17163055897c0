"""Bitwise fingerprint and timing of the fused batched Cholesky (kb_chol) through the C ABI: run before and after a
change of the kernel; equal hashes = bitwise the same factors.  Usage: python tools/kbc_hash.py [B_timing]"""
import ctypes
import hashlib
import sys

import torch

from interiorpointmethod_b200 import _lib

lib = _lib.load()


def spd_batch(B, m, ldm, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.rand(B, m, 2 * m, dtype=torch.float64, device="cuda", generator=g)
    d = torch.rand(B, 1, 2 * m, dtype=torch.float64, device="cuda", generator=g) ** 8 + 1e-12
    M = (A * d) @ A.transpose(1, 2)
    buf = torch.zeros(B, m, ldm, dtype=torch.float64, device="cuda")
    buf[:, :, :m] = M
    return buf


def factor(buf, m, ldm):
    nf = ctypes.c_int(0)
    rc = lib.ipm_potrf_batched_d(0, buf.shape[0], m, ctypes.c_void_p(buf.data_ptr()), ldm, m * ldm, 1e-30,
                                 ctypes.byref(nf))
    assert rc == 0, rc
    torch.cuda.synchronize()
    return nf.value


for (B, m, ldm) in [(64, 256, 256), (64, 256, 260), (32, 200, 200), (32, 130, 132), (32, 96, 96), (16, 33, 34),
                    (16, 32, 32), (8, 7, 8), (4, 512, 512), (4, 400, 400)]:
    buf = spd_batch(B, m, ldm, 1234 + m)
    if m >= 64:
        buf[1, 5, :] = 0.0
        buf[1, :, 5] = 0.0          # a zero pivot: the safeguard path
    nf = factor(buf, m, ldm)
    L = torch.tril(buf[:, :, :m]).contiguous().cpu().numpy()
    print("m=%d ldm=%d B=%d nfixed=%d sha1=%s" % (m, ldm, B, nf, hashlib.sha1(L.tobytes()).hexdigest()))

Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
src = spd_batch(Bt, 256, 256, 7)
work = src.clone()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
for it in range(6):
    work.copy_(src)
    torch.cuda.synchronize()
    ev[0].record()
    nf = ctypes.c_int(0)
    lib.ipm_potrf_batched_d(0, Bt, 256, ctypes.c_void_p(work.data_ptr()), 256, 256 * 256, 1e-30, ctypes.byref(nf))
    ev[1].record()
    torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]))
print("kb_chol m=256 B=%d: ms per launch %s (min %.3f)" % (Bt, ["%.3f" % t for t in times], min(times[1:])))
