"""Per-rank timing of the batched solve under torchrun (diagnosis of multi-GPU scaling): every rank solves the same
number of LPs; prints per-rank solve times, the library's phase sums, the gather time and an nvidia-smi snapshot
taken by rank 0 while all GPUs are busy."""
import ctypes, os, subprocess, sys, threading, time
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch, gather_results
lib = _lib.load()
B, m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 256, 512
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
A = torch.randn(B, m, n, dtype=torch.float64, device=dev, generator=g)
xh = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) + 0.1
sh = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) + 0.1
yh = torch.randn(B, m, dtype=torch.float64, device=dev, generator=g)
b = torch.bmm(A, xh.unsqueeze(2)).squeeze(2); c = torch.bmm(A.transpose(1, 2), yh.unsqueeze(2)).squeeze(2) + sh
db = DeviceBatch(A, b, c)
snap = {}
def smi():
    time.sleep(0.6)
    snap["smi"] = subprocess.run(["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active,utilization.gpu",
                                  "--format=csv,noheader"], capture_output=True, text=True).stdout
def sync_all():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
for _ in range(2): db.solve(tol=1e-8)
sync_all()
if rank == 0: threading.Thread(target=smi).start()
lib.ipm_profile_enable(1)
ts = []
for _ in range(3):
    t = time.perf_counter(); nit = db.solve(tol=1e-8); ts.append((time.perf_counter() - t) * 1e3)
ms = (ctypes.c_double * 4)(); calls = (ctypes.c_int64 * 4)(); lp = ctypes.c_int64(0)
lib.ipm_profile_read(ms, calls, ctypes.byref(lp)); lib.ipm_profile_enable(0)
sync_all()
tg = []
if world > 1:
    for _ in range(3):
        t = time.perf_counter(); out = gather_results(db.obj, db.iters, db.status); torch.cuda.synchronize(); tg.append((time.perf_counter() - t) * 1e3)
# solve + gather back to back, as bench.py's step does
sync_all()
t = time.perf_counter()
for _ in range(3):
    db.solve(tol=1e-8)
    if world > 1: out = gather_results(db.obj, db.iters, db.status)
torch.cuda.synchronize(); loop_ms = (time.perf_counter() - t) * 1e3 / 3
sync_all()
print("rank %d: solve ms %s (its %d) phases/solve %.1f gather ms %s  solve+gather loop %.1f ms  cpus %d load %s" % (
    rank, ["%.1f" % v for v in ts], nit, sum(ms) / 3, ["%.1f" % v for v in tg], loop_ms, len(os.sched_getaffinity(0)), os.getloadavg()), flush=True)
time.sleep(0.5 + 0.05 * rank)
if rank == 0: print(snap.get("smi", ""), flush=True)
if world > 1: dist.destroy_process_group()
