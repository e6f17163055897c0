"""One-shot diagnosis of the 8-GPU weak-scaling anomaly of bench.py (1.9 s per step instead of 0.37 s): the same
solve+gather loop under (a) nothing, (b) bench.py's NVML sampler thread, (c) the old `nvidia-smi -lms 200` child,
(d) 8.6 GB of pinned host memory per rank, (e) the host-buffer entry point (H2D inside the timed region)."""
import importlib.util, os, subprocess, sys, time
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from interiorpointmethod_b200 import _lib
from interiorpointmethod_b200.batch import DeviceBatch, gather_results, solve_batched_pinned
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
lib = _lib.load()
B, m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 256, 512
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
A = torch.randn(B, m, n, dtype=torch.float64, device=dev, generator=g)
xh = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) + 0.1
sh = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) + 0.1
yh = torch.randn(B, m, dtype=torch.float64, device=dev, generator=g)
b = torch.bmm(A, xh.unsqueeze(2)).squeeze(2); c = torch.bmm(A.transpose(1, 2), yh.unsqueeze(2)).squeeze(2) + sh
db = DeviceBatch(A, b, c)

def sync_all():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()

def loop(fn, reps=2):
    sync_all()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) * 1e3 / reps
    sync_all()
    return dt

def step_device():
    db.solve(tol=1e-8)
    if world > 1: gather_results(db.obj, db.iters, db.status)

for _ in range(3): step_device()
res = {}
res["a_plain"] = loop(step_device)
s = bench.ClockSampler(lr) if rank == 0 else None
res["b_nvml_thread"] = loop(step_device)
if s: res["b_samples"] = s.stop()["samples"]
p = None
if rank == 0:
    p = subprocess.Popen(["nvidia-smi", "-i", str(lr), "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    time.sleep(0.3)
res["c_nvidia_smi_lms"] = loop(step_device)
if p: p.terminate(); p.wait()
res["a2_plain_again"] = loop(step_device)
A_h = torch.empty((B, m, n), dtype=torch.float64, pin_memory=True); b_h = torch.empty((B, m), dtype=torch.float64, pin_memory=True)
c_h = torch.empty((B, n), dtype=torch.float64, pin_memory=True)
A_h.copy_(A); b_h.copy_(b); c_h.copy_(c)
res["d_with_pinned_alloc"] = loop(step_device)
obj_h = torch.empty(B, dtype=torch.float64, pin_memory=True); it_h = torch.empty(B, dtype=torch.int32, pin_memory=True); st_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
def step_e2e():
    solve_batched_pinned(A_h, b_h, c_h, obj_h, it_h, st_h, tol=1e-8, device=lr)
    if world > 1: gather_results(obj_h.to(dev), it_h.to(dev), st_h.to(dev))
step_e2e()
res["e_e2e_host_buffers"] = loop(step_e2e)
def h2d_only():
    A.copy_(A_h, non_blocking=True)
res["f_h2d_8.6GB_only"] = loop(h2d_only)
print("rank %d: %s" % (rank, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
if world > 1: dist.destroy_process_group()
