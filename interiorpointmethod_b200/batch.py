"""Batched workload: B independent dense LPs of one shape (BASELINE.json: 8192 x (256 x 512)).

One GPU solves its LPs in lockstep through `ipm_solve_batched_dense[_d]`; across GPUs the batch is partitioned
statically (rank g of G owns LPs [g*B/G, (g+1)*B/G)) with no per-iteration communication, and the objectives,
iteration counts and statuses are gathered once at the end over torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  The dense driver convention applies: start x = s = 1, y = 0, cap 50000 (main.py:287-302, 725).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def shard_range(B: int, rank: int, world: int):
    """Contiguous block partition; the first B % world ranks get one extra LP."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(B, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first, count


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def solve_batched_host(A, b, c, tol=1e-8, max_iter=50000, device=0, want_x=False):
    """Host buffers in, host results out (copies staged inside the call, overlapped with the solve).
    A: (B,m,n) float64 C-contiguous, b: (B,m), c: (B,n).  Returns (obj, iters, status[, x])."""
    lib = _lib.load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    if A.ndim != 3 or b.shape != A.shape[:2] or c.shape != (A.shape[0], A.shape[2]):
        raise ValueError("expected A (B,m,n), b (B,m), c (B,n)")
    B, m, n = A.shape
    obj = np.empty(B)
    iters = np.empty(B, dtype=np.int32)
    status = np.empty(B, dtype=np.int32)
    x = np.empty((B, n)) if want_x else None
    rc = lib.ipm_solve_batched_dense(int(device), B, m, n, _p(A), _p(b), _p(c), float(tol), int(max_iter),
                                     _p(obj), _p(iters), _p(status), _p(x))
    _lib.check(rc, None, "ipm_solve_batched_dense")
    return (obj, iters, status, x) if want_x else (obj, iters, status)


def solve_batched_pinned(A_t, b_t, c_t, obj_t, iters_t, status_t, tol=1e-8, max_iter=50000, device=0):
    """Same entry point on torch CPU tensors in pinned memory (no conversion, no allocation here):
    what bench.py times as the end-to-end path."""
    lib = _lib.load()
    B, m, n = A_t.shape
    rc = lib.ipm_solve_batched_dense(int(device), int(B), int(m), int(n), ctypes.c_void_p(A_t.data_ptr()),
                                     ctypes.c_void_p(b_t.data_ptr()), ctypes.c_void_p(c_t.data_ptr()), float(tol),
                                     int(max_iter), ctypes.c_void_p(obj_t.data_ptr()),
                                     ctypes.c_void_p(iters_t.data_ptr()), ctypes.c_void_p(status_t.data_ptr()), None)
    _lib.check(rc, None, "ipm_solve_batched_dense")


class DeviceBatch:
    """LPs resident on one GPU as torch tensors (torch is only the allocator here); `solve()` runs the whole
    batch through `ipm_solve_batched_dense_d` with a preallocated workspace."""

    def __init__(self, A_d, b_d, c_d):
        import torch

        assert A_d.is_cuda and A_d.dtype == torch.float64 and A_d.is_contiguous()
        self.A, self.b, self.c = A_d, b_d.contiguous(), c_d.contiguous()
        self.B, self.m, self.n = A_d.shape
        self.device = A_d.device.index or 0
        lib = _lib.load()
        nbytes = lib.ipm_batched_workspace_bytes(self.B, self.m, self.n)
        self.work = torch.empty(nbytes, dtype=torch.uint8, device=A_d.device)
        self.obj = torch.empty(self.B, dtype=torch.float64, device=A_d.device)
        self.iters = torch.empty(self.B, dtype=torch.int32, device=A_d.device)
        self.status = torch.empty(self.B, dtype=torch.int32, device=A_d.device)
        self._lib = lib

    def solve(self, tol=1e-8, max_iter=50000) -> int:
        """Returns the number of lockstep iterations run (= max over LPs)."""
        import torch

        torch.cuda.current_stream(self.A.device).synchronize()
        nit = ctypes.c_int(0)
        rc = self._lib.ipm_solve_batched_dense_d(
            self.device, self.B, self.m, self.n, ctypes.c_void_p(self.A.data_ptr()),
            ctypes.c_void_p(self.b.data_ptr()), ctypes.c_void_p(self.c.data_ptr()), float(tol), int(max_iter),
            ctypes.c_void_p(self.obj.data_ptr()), ctypes.c_void_p(self.iters.data_ptr()),
            ctypes.c_void_p(self.status.data_ptr()), None, ctypes.c_void_p(self.work.data_ptr()),
            ctypes.byref(nit))
        _lib.check(rc, None, "ipm_solve_batched_dense_d")
        return nit.value


def gather_results(obj, iters, status, group=None):
    """All-gather of the per-rank result vectors (torch tensors on the backend's device).  Ranks may own
    different counts (B % world != 0): shards are padded to the largest and trimmed after the collective.
    Returns (obj, iters, status) of the whole batch on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    count = torch.tensor([obj.numel()], dtype=torch.int64, device=obj.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(t.item()) for t in counts]
    cap = max(counts)
    packed = torch.zeros(cap, 3, dtype=torch.float64, device=obj.device)
    packed[: obj.numel(), 0] = obj
    packed[: obj.numel(), 1] = iters.to(torch.float64)
    packed[: obj.numel(), 2] = status.to(torch.float64)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    full = torch.cat([o[:k] for o, k in zip(out, counts)], dim=0)
    return full[:, 0].contiguous(), full[:, 1].to(torch.int32), full[:, 2].to(torch.int32)
