#!/usr/bin/env python
"""Benchmark of the Newton-step hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 8192] [--no-extras]

Workload (BASELINE.json configs[4]): a batch of 8192 synthetic dense LPs, m=256, n=512, LP i drawn from
numpy default_rng(i) (SURVEY.md §8d generator), solved to tol=1e-8 with the dense-driver semantics of the
reference (`interior`, main.py:707-757).  One step = one solve of the whole batch.  LPs are partitioned
statically over the N ranks with no data-path collective; the objectives, iteration counts and statuses are
all-gathered once per step.  --scaling weak (default, the rule for partitioned paths): every GPU solves 8192 LPs
(rank r owns LP seeds 8192 r .. 8192 r + 8191); --scaling strong: 8192 LPs in total, 8192/N per GPU
(BASELINE.json's literal "sharded across 1/2/4/8").

  value  LPs/s with the inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e    LPs/s through the C-ABI call that takes HOST buffers (pinned): H2D of A, b, c and D2H of the
         results inside the timed region
  roofline      the batched FP64 DMMA SYRK kernel (M = A diag(x/s) A^T), timed live with CUDA events
  cpu_baseline  the oracle's port of the reference AS WRITTEN (dense (m+2n)^2 KKT + LAPACK dgesv twice per
                iteration, main.py:13-21, 185-194, 232-244) on a bounded sample, one process per host core
  --impl reference   that same CPU port as its own arm (rank 0 only)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_LP, N_LP = 256, 512
TOL = 1e-8
METRIC = "batched LPs/sec (8192 dense LPs m=256 n=512, tol 1e-8)"


# ----------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One LP through the oracle's restatement of the reference as written (runs in a worker process)."""
    seed, linear = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import ipm_oracle as orc      # the one place bench.py executes oracle/ (CPU baseline leg)
    A, b, c = orc.synthetic_dense_lp(M_LP, N_LP, seed)
    t0 = time.perf_counter()
    r = orc.solve(A, b, c, tol=TOL, max_iter=50000, y0_is_one=False, linear=linear)
    dt = time.perf_counter() - t0
    return r["k"], r["obj"], dt


def cpu_sample(pool, cores, lps_per_core, linear, first_seed=0):
    seeds = [(first_seed + i, linear) for i in range(cores * lps_per_core)]
    t0 = time.perf_counter()
    out = pool.map(_cpu_worker, seeds, chunksize=1)
    wall = time.perf_counter() - t0
    return len(seeds) / wall, wall, out


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port, kind 'port':
    /root/reference is Python and cannot travel to the GPU box), all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = host_cores()
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_sample(pool, cores, 1, "kkt")
        t0 = time.perf_counter()
        n_lp = 0
        for k in range(args.steps):
            _, _, out = cpu_sample(pool, cores, 1, "kkt", first_seed=k * cores)
            n_lp += len(out)
        wall = time.perf_counter() - t0
    value = n_lp / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "LPs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batch of 8192 synthetic dense LPs m=256 n=512 (BASELINE.json configs[4])",
                   "batch": args.batch, "m": M_LP, "n": N_LP, "tol": TOL,
                   "step": "bounded sample: %d LPs per step (one per host core), extrapolated as LPs/s" % cores},
        "cpu_baseline": {"value": value, "unit": "LPs/s", "cores": cores, "kind": "port",
                         "sample": "%d LPs per step x %d steps, reference as written (dense KKT + dgesv), "
                                   "1 BLAS thread per process" % (cores, args.steps)},
        "e2e": {"value": value, "unit": "LPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons of this rank's GPU DURING the timed region, sampled in-process through NVML
    (pynvml) every 100 ms by a thread.  An `nvidia-smi -lms` child was used before: on the 8-GPU box its polling
    serialised against the kernel launches of all eight ranks (weak-scaling step 1.9 s instead of 0.37 s,
    profiles/r1_bench_n8_weak_nvidia_smi_sampler.json); NVML calls on one open handle do not."""
    PERIOD_S = 0.1

    def __init__(self, gpu_index):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self._nv, self._h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def _run(self):
        nv, h = self._nv, self._h
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
        if self._thread is None:
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        sm = list(self.samples)
        if sm:
            hi = sorted(sm)[len(sm) // 2:]          # the upper half = samples under load
            out.update(sm_mhz=float(np.median(sm)), sm_mhz_under_load=float(np.median(hi)),
                       reasons=sorted(self.reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- extras
NETLIB_OPT = {"AFIRO": -4.6475314286e02, "SCSD8": 9.0499999993e02, "25FV47": 5.5018458883e03,
              "QAP15": 1.0409940410e03}      # main.py:1417-1516


def extras_netlib(ipm):
    """Newton iterations/s on the Netlib configs (BASELINE.json configs[0..2]); solve time to 1e-8."""
    out = {}
    for name in ("AFIRO", "SCSD8", "25FV47", "QAP15"):
        try:
            A, b, c, cTlb = ipm.load_golden_problem(name)
            with ipm.NewtonStep(A, b, c) as ns:
                cap = 5000 if name != "QAP15" else 60
                ns.solve(tol=TOL, max_iter=min(cap, 20), cTlb=cTlb)                   # warm-up
                t0 = time.perf_counter()
                r = ns.solve(tol=TOL, max_iter=cap, cTlb=cTlb)
                dt = time.perf_counter() - t0
                out[name] = {"m": ns.m, "n": ns.n, "iterations": r.iterations, "status": r.status,
                             "objective": r.objective, "solve_s": dt, "newton_it_per_s": r.iterations / dt}
                # opt-in starting point that is NOT in the reference (ipm_start_mehrotra): same kernels, no
                # iteration parity; it is what lets 25FV47 and QAP15 reach the Netlib optimum
                ns.solve(tol=TOL, max_iter=min(cap, 20), cTlb=cTlb, start="mehrotra")
                t0 = time.perf_counter()
                r = ns.solve(tol=TOL, max_iter=500, cTlb=cTlb, start="mehrotra")
                dt = time.perf_counter() - t0
                out[name]["mehrotra_start"] = {"iterations": r.iterations, "status": r.status, "objective": r.objective,
                                               "solve_s": dt, "newton_it_per_s": r.iterations / max(dt, 1e-12),
                                               "netlib_optimum": NETLIB_OPT.get(name)}
        except Exception as e:  # pragma: no cover
            out.setdefault(name, {})["error"] = str(e)[:200]
    return out


def extras_dense_big(ipm, lib, peak_tf, m=16384, n=32768):
    """One Newton iteration's SYRK + Cholesky at the dense-big shape (BASELINE.json configs[3])."""
    import ctypes
    import torch
    try:
        dev = torch.device("cuda:0")
        g = torch.Generator(device=dev).manual_seed(0)
        A = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
        d = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
        M = torch.empty(m, m, dtype=torch.float64, device=dev)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        res = {}
        for rep in range(2):
            torch.cuda.synchronize()
            e0.record()
            rc = lib.ipm_syrk_d(0, m, n, ctypes.c_void_p(A.data_ptr()), n, ctypes.c_void_p(d.data_ptr()),
                                ctypes.c_void_p(M.data_ptr()), m)
            e1.record()
            nf = ctypes.c_int(0)
            rc2 = lib.ipm_potrf_d(0, m, ctypes.c_void_p(M.data_ptr()), m, 1e-30, ctypes.byref(nf))
            e2.record()
            torch.cuda.synchronize()
            if rc or rc2:
                return {"error": "rc %d %d" % (rc, rc2)}
            t_syrk, t_chol = e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3
            res = {"m": m, "n": n, "syrk_s": t_syrk, "potrf_s": t_chol,
                   "syrk_tflops": m * m * n / t_syrk * 1e-12, "potrf_tflops": m ** 3 / 3 / t_chol * 1e-12,
                   "syrk_plus_potrf_tflops": (m * m * n + m ** 3 / 3) / (t_syrk + t_chol) * 1e-12,
                   "frac_of_dmma_peak": (m * m * n + m ** 3 / 3) / (t_syrk + t_chol) * 1e-12 / peak_tf,
                   "pivots_fixed": nf.value}
        del A, M
        torch.cuda.empty_cache()
        return res
    except Exception as e:  # pragma: no cover
        return {"error": str(e)[:200]}


# ----------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU solves --batch LPs (LP seeds rank*batch + i), no data-path "
                         "collective; strong: --batch LPs in total, block-partitioned over the GPUs")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3          # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import ctypes
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    import interiorpointmethod_b200 as ipm
    from interiorpointmethod_b200 import _lib, build
    from interiorpointmethod_b200.batch import DeviceBatch, gather_results, shard_range, solve_batched_pinned
    build.build()
    lib = _lib.load()

    # weak scaling (tier rule for partitioned paths): per-GPU work fixed, the batch grows with the GPU count;
    # strong: BASELINE.json's literal "8192 LPs sharded across 1/2/4/8"
    B = args.batch * world if args.scaling == "weak" else args.batch
    first, count = shard_range(B, rank, world)
    # ---- inputs: generated on the host (exactly the reference-side generator), pinned for the e2e arm
    t0 = time.perf_counter()
    A_h = torch.empty((count, M_LP, N_LP), dtype=torch.float64, pin_memory=True)
    b_h = torch.empty((count, M_LP), dtype=torch.float64, pin_memory=True)
    c_h = torch.empty((count, N_LP), dtype=torch.float64, pin_memory=True)
    ipm.synthetic_dense_batch(first, count, M_LP, N_LP, out_A=A_h.numpy(), out_b=b_h.numpy(), out_c=c_h.numpy(),
                              threads=max(1, min(32, host_cores() // max(1, world))))
    gen_s = time.perf_counter() - t0
    obj_h = torch.empty(count, dtype=torch.float64, pin_memory=True)
    it_h = torch.empty(count, dtype=torch.int32, pin_memory=True)
    st_h = torch.empty(count, dtype=torch.int32, pin_memory=True)

    db = DeviceBatch(A_h.to(dev), b_h.to(dev), c_h.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        db.solve(tol=TOL)
        if world > 1:
            return gather_results(db.obj, db.iters, db.status)
        return db.obj, db.iters, db.status

    def step_e2e():
        solve_batched_pinned(A_h, b_h, c_h, obj_h, it_h, st_h, tol=TOL, device=local_rank)
        if world > 1:
            return gather_results(obj_h.to(dev), it_h.to(dev), st_h.to(dev))
        return obj_h, it_h, st_h

    def timed(fn, steps):
        """K steps between barrier+synchronize pairs, CUDA events on the current (launching) stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], dtype=torch.float64, device=dev)
        per_rank = [float(t[0])]
        if world > 1:
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            per_rank = [float(v[0]) for v in every]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timed.per_rank_ms = [v * 1e3 / steps for v in per_rank]
        return float(t[0]), float(t[1]), out

    # ---- device-resident arm
    for _ in range(args.warmup):
        step_device()
    peak_tf = lib.ipm_measure_dmma_peak(local_rank) if rank == 0 else 0.0
    lib.ipm_profile_enable(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.ipm_launch_count()
    t_dev, wall_dev, out = timed(step_device, args.steps)
    per_rank_dev = list(timed.per_rank_ms)
    launches = lib.ipm_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms = (ctypes.c_double * 4)()
    calls = (ctypes.c_int64 * 4)()
    lp_it = ctypes.c_int64(0)
    lib.ipm_profile_read(ms, calls, ctypes.byref(lp_it))
    trace_ms = (ctypes.c_double * 512)()
    trace_ph = (ctypes.c_int * 512)()
    ntr = lib.ipm_profile_last(trace_ms, trace_ph, 512)
    lib.ipm_profile_enable(0)
    obj_all, it_all, st_all = (t.cpu().numpy() for t in out)

    # ---- end-to-end arm (host buffers through the C ABI)
    step_e2e()
    t_e2e, wall_e2e, out2 = timed(step_e2e, args.steps)
    obj2 = out2[0].cpu().numpy()

    lt = torch.tensor([launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- rank 0: checks, roofline, CPU baseline, JSON line
    n_conv = int((st_all == 0).sum())
    syrk_s = ms[1] * 1e-3
    syrk_flops = float(lp_it.value) * M_LP * M_LP * N_LP          # symmetric count m^2 n per LP-iteration
    achieved_tf = syrk_flops / syrk_s * 1e-12 if syrk_s > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            per_lp = json.load(open(tp)).get("syrk_batched_dram_bytes_per_lp")
            # ncu capture (one launch, all LPs active) scaled to the average number of active LPs per launch
            traffic = per_lp * float(lp_it.value) / max(1, calls[1]) if per_lp else None
        except Exception:
            traffic = None
    phase_total = sum(ms) * 1e-3
    roofline = {
        "kernel": "dmma_ws_kernel<0,true> (batched SYRK M = A diag(x/s) A^T, warp-specialised persistent, DMMA.8x8x4)",
        "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic,
        "peak_source": "FP64 DMMA issue-rate ceiling measured live on this GPU (ipm_measure_dmma_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 tensor = 40 TFLOP/s",
        "flops_per_launch": syrk_flops / max(1, calls[1]), "avg_launch_ms": ms[1] / max(1, calls[1]),
        "launches": int(calls[1]),
        "share_of_step": syrk_s / t_dev if t_dev > 0 else None,
        "last_step_syrk_ms_per_iteration": [round(trace_ms[i], 3) for i in range(min(ntr, 512)) if trace_ph[i] == 1],
        "last_step_cholesky_ms_per_iteration": [round(trace_ms[i], 3) for i in range(min(ntr, 512)) if trace_ph[i] == 2],
        "phase_ms_per_step": {"residual_pass": ms[0] / args.steps, "syrk": ms[1] / args.steps,
                              "cholesky": ms[2] / args.steps, "solves_and_update": ms[3] / args.steps,
                              "sum": phase_total / args.steps * 1e3},
    }
    cpu = None
    if not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = host_cores()
        ctx = mp.get_context("spawn")
        with ctx.Pool(cores) as pool:
            v_kkt, wall_kkt, outk = cpu_sample(pool, cores, 1, "kkt")
            v_ne, wall_ne, outn = cpu_sample(pool, cores, 4, "normal")
        # parity spot check of the GPU results against the CPU port on the sample it just solved
        nchk = min(len(outk), B)
        dk = max(abs(int(it_all[i]) - outk[i][0]) for i in range(nchk))
        dobj = max(abs(obj_all[i] - outk[i][1]) / abs(outk[i][1]) for i in range(nchk))
        cpu = {"value": v_kkt, "unit": "LPs/s", "cores": cores, "kind": "port",
               "sample": "%d LPs (seeds 0..%d), reference as written: dense (m+2n)^2 KKT + dgesv twice per "
                         "iteration (main.py:13-21,185-194,232-244), one process per core, 1 BLAS thread each, "
                         "%.1f s wall" % (len(outk), len(outk) - 1, wall_kkt),
               "normal_equations_port": {"value": v_ne, "unit": "LPs/s",
                                         "sample": "%d LPs, same Newton step via main.py:221-229 + safeguarded "
                                                   "Cholesky, %.1f s wall" % (len(outn), wall_ne)},
               "gpu_vs_port_on_sample": {"max_iteration_diff": int(dk), "max_rel_objective_diff": float(dobj)}}

    line = {
        "metric": METRIC, "value": B * args.steps / t_dev, "unit": "LPs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batch of 8192 synthetic dense LPs m=256 n=512 (BASELINE.json configs[4])"
                               + (" on every GPU, LP seeds 0..%d" % (B - 1) if world > 1 and args.scaling == "weak" else ""),
                   "batch": B, "m": M_LP, "n": N_LP, "tol": TOL, "partition": "static block, %d LPs per GPU" % count,
                   "l2": "inputs per GPU (%.2f GB) exceed L2, no flush" % (count * M_LP * N_LP * 8 / 1e9),
                   "newton_iterations_per_step": int(it_all.sum()), "lockstep_iterations": int(it_all.max())},
        "newton_it_per_s": float(it_all.sum()) * args.steps / t_dev,
        "wall_ms_per_step": wall_dev / args.steps * 1e3,
        "per_rank_ms_per_step": [round(v, 2) for v in per_rank_dev],
        "converged": n_conv, "iterations_min_max": [int(it_all.min()), int(it_all.max())],
        "e2e": {"value": B * args.steps / t_e2e, "unit": "LPs/s",
                "h2d_bytes_per_step": int(B) * (M_LP * N_LP + M_LP + N_LP) * 8,
                "d2h_bytes_per_step": int(B) * 16, "ms_per_step": t_e2e / args.steps * 1e3,
                "max_abs_diff_vs_device_arm": float(np.max(np.abs(obj2 - obj_all)))},
        "gpu_launches": int(lt.item()),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "setup": {"host_generation_s": gen_s},
    }
    if not args.no_extras and world == 1:
        del db
        torch.cuda.empty_cache()
        line["extras"] = {"netlib": extras_netlib(ipm), "dense_big": extras_dense_big(ipm, lib, peak_tf)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
