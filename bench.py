#!/usr/bin/env python
"""Benchmark of the Newton-step hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 8192] [--no-extras]

Workload (BASELINE.json configs[4]): a batch of 8192 synthetic dense LPs, m=256, n=512, LP i drawn from
numpy default_rng(i) (SURVEY.md §8d generator), solved to tol=1e-8 with the dense-driver semantics of the
reference (`interior`, main.py:707-757).  One step = one solve of the whole batch.  LPs are partitioned
statically over the N ranks with no data-path collective; the objectives, iteration counts and statuses are
all-gathered once per step.  --scaling weak (default, the rule for partitioned paths): every GPU solves 8192 LPs
(rank r owns LP seeds 8192 r .. 8192 r + 8191); --scaling strong: 8192 LPs in total, 8192/N per GPU
(BASELINE.json's literal "sharded across 1/2/4/8").  With N > 1 the weak line also carries the strong-scaling
measurement of the same run in `config.strong`.

  value  LPs/s with the inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e    LPs/s through the C-ABI call that takes HOST buffers (pinned): H2D of A, b, c and D2H of the
         results inside the timed region
  roofline      the batched FP64 DMMA SYRK kernel (M = A diag(x/s) A^T), timed live with CUDA events;
                `whole_step_frac` = all FP64 work of the step against the same peak
  cpu_baseline  the reference AS WRITTEN (dense (m+2n)^2 KKT + LAPACK dgesv twice per iteration, main.py:13-21,
                185-194, 232-244) on a bounded sample, one process per host core: the UNMODIFIED reference from
                oracle/_ref when build() could copy it (kind "reference"), else the oracle's port (kind "port")
  --impl reference   that same CPU path as its own arm (rank 0 only)

The run FAILS (exit code 3, "parity": "FAILED" in the line) unless every LP of every rank converged and agrees with
the frozen oracle table (tests/golden/batch_256x512_oracle.npz) within +-1 iteration and 1e-8 relative objective.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_LP, N_LP = 256, 512
TOL = 1e-8
METRIC = "batched LPs/sec (8192 dense LPs m=256 n=512, tol 1e-8)"
F_LP_ITER = M_LP * M_LP * N_LP + M_LP ** 3 / 3 + 4 * M_LP ** 2 + 12 * M_LP * N_LP     # 4.098e7 flop


# ----------------------------------------------------------------------------------------------- CPU arm
def _ref_module_available():
    return (os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "main.py"))
            or os.path.isfile("/root/reference/main.py"))


def _cpu_worker(args):
    """One LP on one host core (runs in a worker process, 1 BLAS/OpenMP thread).
    linear = "ref"   : the UNMODIFIED reference's `interior` functions (oracle/_ref, loop replayed by ref_harness so
                       that k and the objective come back; main.py:718-751)
             "kkt"   : the oracle's port of the same (dense KKT + dgesv)
             "normal": the oracle's normal-equations iteration (the elimination the GPU runs)"""
    seed, linear = args
    from oracle import ipm_oracle as orc      # the one place bench.py executes oracle/ (CPU baseline legs)
    A, b, c = orc.synthetic_dense_lp(M_LP, N_LP, seed)
    t0 = time.perf_counter()
    if linear == "ref":
        from oracle import ref_harness as rh
        with rh.quiet():
            r = rh.replay_interior_dense(A, b, c, tol=TOL)
    elif linear == "ref_driver":
        from oracle import ref_harness as rh
        rh.call_interior_dense(A, b, c, tol=TOL)          # main.interior itself; prints, returns None
        r = {"k": -1, "obj": float("nan")}
    else:
        r = orc.solve(A, b, c, tol=TOL, max_iter=50000, y0_is_one=False, linear=linear,
                      refine_thresh=1.0 if linear == "normal" else None)
    dt = time.perf_counter() - t0
    return r["k"], r["obj"], dt


@contextlib.contextmanager
def single_threaded_children():
    """Worker processes inherit the environment at spawn: one BLAS/OpenMP thread each, whatever the box sets
    (the oracle's helper library oversubscribed 16 workers x 16 OpenMP threads on the 1-GPU box)."""
    keys = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ[k] = "1"
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


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
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores, a bounded sample
    per step (one LP per core).  The UNMODIFIED reference driver `main.interior` from oracle/_ref when present
    (kind "reference"), else the oracle's port of it (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = host_cores()
    kind = "reference" if _ref_module_available() else "port"
    linear = "ref_driver" if kind == "reference" else "kkt"
    ctx = mp.get_context("spawn")
    with single_threaded_children(), ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_sample(pool, cores, 1, linear)
        t0 = time.perf_counter()
        n_lp = 0
        for k in range(args.steps):
            _, _, out = cpu_sample(pool, cores, 1, linear, first_seed=k * cores)
            n_lp += len(out)
        wall = time.perf_counter() - t0
    value = n_lp / wall
    what = ("UNMODIFIED reference main.interior (oracle/_ref, dense KKT + np.linalg.solve twice per iteration)"
            if kind == "reference" else "oracle port of the reference as written (dense KKT + dgesv)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "LPs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batch of 8192 synthetic dense LPs m=256 n=512 (BASELINE.json configs[4])",
                   "batch": args.batch, "m": M_LP, "n": N_LP, "tol": TOL,
                   "step": "bounded sample: %d LPs per step (one per host core), extrapolated as LPs/s" % cores},
        "cpu_baseline": {"value": value, "unit": "LPs/s", "cores": cores, "kind": kind,
                         "sample": "%d LPs per step x %d steps, %s, 1 BLAS thread per process" % (cores, args.steps, what)},
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


def _netlib_cpu_worker(name):
    """The reference's `interior_sparse` semantics (main.py:776-815: full KKT + SuperLU twice per iteration) on one
    Netlib LP, timed on one host core: the unmodified reference's own functions when oracle/_ref is present."""
    import interiorpointmethod_b200 as ipm
    from oracle import ipm_oracle as orc
    A, b, c, cTlb = ipm.load_golden_problem(name)
    t0 = time.perf_counter()
    if _ref_module_available():
        from oracle import ref_harness as rh
        import scipy.sparse as sp
        with rh.quiet():
            r = rh.replay_interior_sparse(sp.csc_matrix(A), b, c, cTlb, tol=TOL)
        kind = "reference"
    else:
        r = orc.solve(A, b, c, cTlb=cTlb, tol=TOL, linear="kkt")
        kind = "port"
    dt = time.perf_counter() - t0
    return {"kind": kind, "cores": 1, "iterations": int(r["k"]), "objective": float(r["obj"]), "solve_s": dt,
            "newton_it_per_s": r["k"] / dt if dt > 0 else None}


def extras_netlib(ipm, cpu_pool=None):
    """Newton iterations/s on the Netlib configs (BASELINE.json configs[0..2]); solve time to 1e-8, with the
    reference's CPU path on the same LP beside it where it terminates (AFIRO, SCSD8: App. C.1; 25FV47 NaN at k = 1,
    QAP15 > 40 min per iteration: SURVEY App. C.2)."""
    out = {}
    cpu_async = {}
    if cpu_pool is not None:
        for name in ("AFIRO", "SCSD8"):
            cpu_async[name] = cpu_pool.apply_async(_netlib_cpu_worker, (name,))
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
                if name == "QAP15":
                    # opt-in that is NOT in the reference either (ipm_detect_dependent_rows): QAP15's A is 10 % rank
                    # deficient; with the dependent rows removed from the normal equations the reference's own
                    # iteration (start x = s = 1, y = 1) reaches the Netlib optimum
                    nd = ns.detect_dependent_rows(1e-10)
                    ns.solve(tol=TOL, max_iter=10, cTlb=cTlb)
                    t0 = time.perf_counter()
                    r = ns.solve(tol=TOL, max_iter=300, cTlb=cTlb)
                    dt = time.perf_counter() - t0
                    opt = NETLIB_OPT[name]
                    out[name]["dependent_rows_removed"] = {
                        "dependent_rows": nd, "iterations": r.iterations, "status": r.status, "objective": r.objective,
                        "solve_s": dt, "newton_it_per_s": r.iterations / max(dt, 1e-12), "netlib_optimum": opt,
                        "rel_objective_diff": abs(r.objective - opt) / abs(opt)}
                    ns.detect_dependent_rows(0.0)
        except Exception as e:  # pragma: no cover
            out.setdefault(name, {})["error"] = str(e)[:200]
    for name, fut in cpu_async.items():
        try:
            out[name]["cpu_baseline"] = fut.get(timeout=120)
            g, cb = out[name], out[name]["cpu_baseline"]
            cb["gpu_vs_cpu"] = {"iteration_diff": g["iterations"] - cb["iterations"],
                                "rel_objective_diff": abs(g["objective"] - cb["objective"]) / max(1.0, abs(cb["objective"]))}
        except Exception as e:  # pragma: no cover
            out[name]["cpu_baseline"] = {"error": str(e)[:200]}
    for name, why in (("25FV47", "reference: NaN at k = 1 (empty row, SuperLU; SURVEY App. C.2)"),
                      ("QAP15", "reference: > 40 min per iteration (SuperLU on the 50880-order KKT matrix; SURVEY App. C.2)")):
        if name in out:
            out[name]["cpu_baseline"] = {"unavailable": why}
    # achieved HBM rates of the sparse / elementwise kernels come from committed ncu captures (profiles/)
    kp = os.path.join(ROOT, "profiles", "kernel_rates.json")
    if os.path.exists(kp):
        try:
            out["kernel_rates_from_ncu"] = json.load(open(kp))
        except Exception:
            pass
    return out


def extras_dense_big(ipm, lib, peak_tf, m=16384, n=32768, cpu_iterations=1, full_solve=True):
    """BASELINE.json configs[3] as a real Newton-step config: a strictly feasible dense LP generated on the device
    (SURVEY 8(d) construction: b = A x^, c = A^T y^ + s^), loaded with ipm_load_dense_d and run through ipm_solve -
    SYRK on the DMMA pipe, blocked Cholesky with look-ahead, pipelined triangular solves, GEMVs, fused vector
    kernels.  F = m^2 n + m^3/3 + 4 m^2 + 12 m n flop per iteration."""
    import ctypes
    import torch
    res = {"m": m, "n": n}
    try:
        dev = torch.device("cuda:0")
        g = torch.Generator(device=dev).manual_seed(0)
        A = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
        xh = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
        sh = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
        yh = torch.randn(m, dtype=torch.float64, device=dev, generator=g)
        b = A @ xh
        c = A.t() @ yh + sh
        F = float(m) * m * n + m ** 3 / 3.0 + 4.0 * m * m + 12.0 * m * n
        h = ctypes.c_void_p()
        from interiorpointmethod_b200 import _lib
        _lib.check(lib.ipm_create(ctypes.byref(h), 0), None, "ipm_create")
        try:
            _lib.check(lib.ipm_load_dense_d(h, m, n, ctypes.c_void_p(A.data_ptr()), n, ctypes.c_void_p(b.data_ptr()),
                                            ctypes.c_void_p(c.data_ptr())), h, "ipm_load_dense_d")
            xs, ys, ss = np.empty(n), np.empty(m), np.empty(n)
            obj, it, st = ctypes.c_double(0), ctypes.c_int(0), ctypes.c_int(0)
            rs = (ctypes.c_double * 5)()

            def solve(max_iter):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = lib.ipm_solve(h, TOL, max_iter, 0, xs.ctypes.data_as(ctypes.c_void_p), ys.ctypes.data_as(ctypes.c_void_p),
                                   ss.ctypes.data_as(ctypes.c_void_p), ctypes.byref(obj), ctypes.byref(it), ctypes.byref(st), rs)
                _lib.check(rc, h, "ipm_solve")
                return time.perf_counter() - t0
            solve(1)                                   # warm-up: kernel attributes, graph capture on the next call
            x1, y1, s1 = xs.copy(), ys.copy(), ss.copy()
            t3 = solve(4)
            res.update({"iterations_timed": int(it.value), "s_per_iteration": t3 / max(1, it.value),
                        "newton_it_per_s": it.value / t3, "flop_per_iteration": F,
                        "tflops": F * it.value / t3 * 1e-12,
                        "frac_of_dmma_peak": F * it.value / t3 * 1e-12 / peak_tf if peak_tf else None,
                        "frac_of_nominal_40tf": F * it.value / t3 * 1e-12 / 40.0})
            if full_solve:
                tf = solve(200)
                res["full_solve"] = {"iterations": int(it.value), "status": int(st.value), "objective": float(obj.value),
                                     "solve_s": tf, "newton_it_per_s": it.value / tf,
                                     "rb_norm": rs[0], "rc_norm": rs[1], "gap": rs[2],
                                     "rb_threshold": TOL * (1 + rs[3]), "rc_threshold": TOL * (1 + rs[4]),
                                     "weak_duality_bound_cTxhat": float((c @ xh).item())}
            # kernel-level numbers at this shape (what round 1 reported), for continuity
            d = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.1
            Mw = torch.empty(m, m, dtype=torch.float64, device=dev)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            for rep in range(2):
                torch.cuda.synchronize()
                e0.record()
                rc = lib.ipm_syrk_d(0, m, n, ctypes.c_void_p(A.data_ptr()), n, ctypes.c_void_p(d.data_ptr()),
                                    ctypes.c_void_p(Mw.data_ptr()), m)
                e1.record()
                nf = ctypes.c_int(0)
                rc2 = lib.ipm_potrf_d(0, m, ctypes.c_void_p(Mw.data_ptr()), m, 1e-30, ctypes.byref(nf))
                e2.record()
                torch.cuda.synchronize()
                t_syrk, t_chol = e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3
                res["kernels"] = {"syrk_s": t_syrk, "potrf_s": t_chol, "syrk_tflops": m * m * n / t_syrk * 1e-12,
                                  "potrf_tflops": m ** 3 / 3 / t_chol * 1e-12,
                                  "syrk_plus_potrf_tflops": (m * m * n + m ** 3 / 3) / (t_syrk + t_chol) * 1e-12,
                                  "frac_of_dmma_peak": (m * m * n + m ** 3 / 3) / (t_syrk + t_chol) * 1e-12 / peak_tf
                                  if peak_tf else None, "rc": [rc, rc2]}
            del Mw
            # ---- parity pin + CPU baseline: the first iteration(s) restated on the host (LAPACK, all cores)
            if cpu_iterations > 0:
                from scipy.linalg import blas, cho_factor, cho_solve
                Ah = A.cpu().numpy()
                bh, ch = b.cpu().numpy(), c.cpu().numpy()
                x, y, s = np.ones(n), np.zeros(m), np.ones(n)
                t0 = time.perf_counter()
                rb = Ah @ x - bh
                rc_ = Ah.T @ y + s - ch
                dd = x / s
                As = Ah * np.sqrt(dd)[None, :]
                # dsyrk on the Fortran view of the C-ordered array: As (C, m x n) = As^T (F, n x m); trans=1 gives As As^T
                Mh = blas.dsyrk(1.0, As.T, trans=1, lower=1)
                t_syrk = time.perf_counter() - t0
                del As
                cf = cho_factor(Mh, lower=True, overwrite_a=True, check_finite=False)
                t_chol = time.perf_counter() - t0 - t_syrk

                def direction(rcomp):
                    tt = rc_ - rcomp / x
                    dy = cho_solve(cf, -rb - Ah @ (dd * tt), check_finite=False)
                    dx = dd * (Ah.T @ dy) + dd * tt
                    ds = -s * dx / x - rcomp / x
                    return dx, dy, ds

                def ratio(v, dv):
                    neg = dv < 0
                    return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0
                r3 = x * s
                dxa, dya, dsa = direction(r3)
                apa, ada = ratio(x, dxa), ratio(s, dsa)
                mu_aff = float((x + apa * dxa) @ (s + ada * dsa)) / n
                mu = float(x @ s) / n
                sigma = (mu_aff / mu) ** 3
                dx, dy, ds = direction(r3 + dxa * dsa - sigma * mu)
                ap, ad = min(1.0, 0.91 * ratio(x, dx)), min(1.0, 0.91 * ratio(s, ds))
                xc, yc, sc = x + ap * dx, y + ad * dy, s + ad * ds
                t_it = time.perf_counter() - t0
                del Mh, cf, Ah
                rel = lambda u, v: float(np.linalg.norm(u - v) / max(1e-300, np.linalg.norm(v)))
                res["cpu_baseline"] = {"kind": "port", "cores": host_cores(),
                                       "what": "CPU restatement of main.py:221-229 with BLAS/LAPACK (dsyrk, cho_factor, "
                                               "cho_solve, gemv), one full predictor-corrector iteration from the start point",
                                       "s_per_iteration": t_it, "syrk_s": t_syrk, "cho_factor_s": t_chol,
                                       "newton_it_per_s": 1.0 / t_it}
                res["parity_iteration_1"] = {"rel_diff_x": rel(x1, xc), "rel_diff_y": rel(y1, yc),
                                             "rel_diff_s": rel(s1, sc), "tolerance": 1e-9}
        finally:
            lib.ipm_destroy(h)
        del A
        torch.cuda.empty_cache()
        return res
    except Exception as e:  # pragma: no cover
        res["error"] = str(e)[:300]
        return res


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

    def pin_to_gpu_numa_node():
        """Host buffers of this rank on the NUMA node its GPU hangs off (first-touch happens right after): eight
        concurrent 8.6 GB H2D copies otherwise all pull through node 0's root complex."""
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            bus = pynvml.nvmlDeviceGetPciInfo(hnd).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]).read())
            if node < 0:
                return None
            cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
            ids = set()
            for part in cpus.split(","):
                lo, _, hi = part.partition("-")
                ids.update(range(int(lo), int(hi or lo) + 1))
            allowed = ids & set(os.sched_getaffinity(0))
            if allowed:
                os.sched_setaffinity(0, allowed)
                return node
        except Exception:
            return None
        return None

    numa_node = pin_to_gpu_numa_node() if world > 1 else None

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
    # The strong-scaling arm's shard is generated HERE, long before it is timed: the generator's worker threads (numpy /
    # BLAS) keep spinning for a while after they finish, and a 1024-LP lockstep iteration (2.8 ms) leaves the solver's host
    # thread no slack for sharing a core with them (tools/strong_shards.py: solves of 55 ms took 80-150 ms when they
    # followed the generation directly; round 2 call 5 reported 75 ms at N = 8 for that reason).
    dbs = None
    if world > 1 and args.scaling == "weak":
        sfirst, scount = shard_range(args.batch, rank, world)
        As, bs, cs = ipm.synthetic_dense_batch(sfirst, scount, M_LP, N_LP, threads=max(1, min(16, host_cores() // world)))
        dbs = DeviceBatch(torch.from_numpy(As).to(dev), torch.from_numpy(bs).to(dev), torch.from_numpy(cs).to(dev))
        del As, bs, cs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(batch=None):
        batch = batch or db
        batch.solve(tol=TOL)
        if world > 1:
            return gather_results(batch.obj, batch.iters, batch.status)
        return batch.obj, batch.iters, batch.status

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

    # ---- frozen oracle table of the generator: the parity gate of this run
    table = np.load(os.path.join(ROOT, "tests", "golden", "batch_256x512_oracle.npz"))
    tab_k, tab_obj = table["k"].astype(np.int64), table["obj"]

    def parity(first_seed, obj, it, st):
        """(converged, max |k - k_table|, max relative objective difference) of LPs first_seed.. against the table."""
        nn = len(obj)
        if first_seed + nn > len(tab_k) or (tab_k[first_seed:first_seed + nn] <= 0).any():
            return int((st == 0).sum()), -1, float("nan")
        dk = np.abs(it.astype(np.int64) - tab_k[first_seed:first_seed + nn])
        to = tab_obj[first_seed:first_seed + nn]
        dobj = np.abs(obj - to) / np.maximum(1.0, np.abs(to))
        return int((st == 0).sum()), int(dk.max()), float(np.nanmax(dobj)) if np.isfinite(dobj).any() else float("inf")

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
    handoffs = lib.ipm_batched_last_handoffs()
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
    my = slice(first, first + count) if world > 1 else slice(0, count)
    par_dev = parity(first, obj_all[my], it_all[my], st_all[my])

    # ---- end-to-end arm (host buffers through the C ABI)
    step_e2e()
    t_e2e, wall_e2e, out2 = timed(step_e2e, args.steps)
    per_rank_e2e = list(timed.per_rank_ms)
    obj2 = out2[0].cpu().numpy()
    par_e2e = parity(first, obj_h.numpy(), it_h.numpy(), st_h.numpy())

    # ---- strong scaling beside weak (N > 1): BASELINE.json's literal "8192 LPs sharded across N"
    strong = None
    if dbs is not None:
        for _ in range(args.warmup):
            step_device(dbs)
        t_s, _, outs = timed(lambda: step_device(dbs), args.steps)
        so, si, ss_ = (t.cpu().numpy() for t in outs)
        ps = parity(0, so, si, ss_)          # gathered: all args.batch LPs, seeds 0..batch-1
        strong = {"value": args.batch * args.steps / t_s, "unit": "LPs/s", "ms_per_step": t_s / args.steps * 1e3,
                  "batch": args.batch, "lps_per_gpu": scount, "per_rank_ms_per_step": [round(v, 2) for v in timed.per_rank_ms],
                  "converged": ps[0], "max_iteration_diff_vs_oracle_table": ps[1],
                  "max_rel_objective_diff_vs_oracle_table": ps[2]}
        del dbs

    # ---- parity gate over all ranks (table covers generator seeds 0..65535: every default run at N <= 8)
    table_covers = par_dev[1] >= 0 and par_e2e[1] >= 0
    bad_local = 0
    for (nc, dk, dobj) in (par_dev, par_e2e):
        if nc != count or (table_covers and (dk > 1 or not (dobj <= 1e-8))):
            bad_local = 1
    sums = torch.tensor([launches, bad_local, par_dev[0], count - par_e2e[0], 0 if table_covers else 1],
                        dtype=torch.float64, device=dev)
    worst_obj = max(par_dev[2], par_e2e[2]) if table_covers else 0.0
    maxs = torch.tensor([max(par_dev[1], par_e2e[1]), worst_obj if np.isfinite(worst_obj) else 1e300],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    parity_failed = sums[1].item() > 0
    if strong is not None:
        parity_failed = parity_failed or strong["converged"] != args.batch \
            or not (0 <= strong["max_iteration_diff_vs_oracle_table"] <= 1) \
            or not (strong["max_rel_objective_diff_vs_oracle_table"] <= 1e-8)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 3 if parity_failed else 0

    # ---- rank 0: roofline, CPU baseline, JSON line
    n_conv = int(sums[2].item())
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
    # rank 0's own share of the step: lp_it counts rank 0's LP-iterations over args.steps solves
    whole_tf = float(lp_it.value) * F_LP_ITER / (per_rank_dev[0] * 1e-3 * args.steps) * 1e-12
    roofline = {
        "kernel": "dmma_ws_kernel<0,true,16,true> (batched SYRK M = A diag(x/s) A^T, warp-specialised persistent, DMMA.8x8x4; "
                  "its diagonal tiles also form the predictor right-hand side; the timed phase includes the elementwise "
                  "kb_wvec launch that precedes it)",
        "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic,
        "peak_source": "FP64 DMMA issue-rate ceiling measured live on this GPU (ipm_measure_dmma_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 tensor = 40 TFLOP/s",
        "flops_per_launch": syrk_flops / max(1, calls[1]), "avg_launch_ms": ms[1] / max(1, calls[1]),
        "launches": int(calls[1]),
        "share_of_step": syrk_s / t_dev if t_dev > 0 else None,
        "whole_step_tflops": whole_tf, "whole_step_frac": whole_tf / peak_tf if peak_tf > 0 else None,
        "whole_step_flop_per_lp_iteration": F_LP_ITER,
        "full_batch_launch": (lambda v: None if not v else {
            "ms": sorted(v)[len(v) // 2], "tflops": count * M_LP * M_LP * N_LP / (sorted(v)[len(v) // 2] * 1e-3) * 1e-12,
            "frac": (count * M_LP * M_LP * N_LP / (sorted(v)[len(v) // 2] * 1e-3) * 1e-12 / peak_tf) if peak_tf > 0 else None,
            "what": "median SYRK phase of the lockstep iterations in which every LP of the batch is still active"})(
            [trace_ms[i] for i in range(min(ntr, 512)) if trace_ph[i] == 1][:12]),
        "last_step_syrk_ms_per_iteration": [round(trace_ms[i], 3) for i in range(min(ntr, 512)) if trace_ph[i] == 1],
        "last_step_cholesky_ms_per_iteration": [round(trace_ms[i], 3) for i in range(min(ntr, 512)) if trace_ph[i] == 2],
        "phase_ms_per_step": {"residual_pass": ms[0] / args.steps, "syrk": ms[1] / args.steps,
                              "cholesky": ms[2] / args.steps, "solves_and_update": ms[3] / args.steps,
                              "sum": phase_total / args.steps * 1e3},
    }
    cpu = None
    cpu_pool = None
    if not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = host_cores()
        ctx = mp.get_context("spawn")
        kind = "reference" if _ref_module_available() else "port"
        with single_threaded_children():
            cpu_pool = ctx.Pool(cores)
        v_kkt, wall_kkt, outk = cpu_sample(cpu_pool, cores, 1, "ref" if kind == "reference" else "kkt")
        v_ne, wall_ne, outn = cpu_sample(cpu_pool, cores, 4, "normal")
        # parity spot check of the GPU results against the CPU path on the sample it just solved
        nchk = min(len(outk), count)
        dk = max(abs(int(it_all[i]) - outk[i][0]) for i in range(nchk))
        dobj = max(abs(obj_all[i] - outk[i][1]) / abs(outk[i][1]) for i in range(nchk))
        what = ("UNMODIFIED reference functions from oracle/_ref in the order of main.interior (main.py:718-751)"
                if kind == "reference" else "oracle port of the reference as written")
        cpu = {"value": v_kkt, "unit": "LPs/s", "cores": cores, "kind": kind,
               "sample": "%d LPs (seeds 0..%d), %s: dense (m+2n)^2 KKT + dgesv twice per iteration "
                         "(main.py:13-21,185-194,232-244), one process per core, 1 BLAS thread each, "
                         "%.1f s wall" % (len(outk), len(outk) - 1, what, wall_kkt),
               "normal_equations_port": {"value": v_ne, "unit": "LPs/s",
                                         "sample": "%d LPs, same Newton step via main.py:221-229 + safeguarded "
                                                   "Cholesky + the refinement rule, 1 thread per process, %.1f s wall"
                                                   % (len(outn), wall_ne)},
               "gpu_vs_cpu_on_sample": {"max_iteration_diff": int(dk), "max_rel_objective_diff": float(dobj)}}
        if dk > 1 or not (dobj <= 1e-8):
            parity_failed = True

    h2d = int(B) * (M_LP * N_LP + M_LP + N_LP) * 8
    line = {
        "metric": METRIC, "value": B * args.steps / t_dev, "unit": "LPs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batch of 8192 synthetic dense LPs m=256 n=512 (BASELINE.json configs[4])"
                               + (" on every GPU, LP seeds 0..%d" % (B - 1) if world > 1 and args.scaling == "weak" else ""),
                   "batch": B, "m": M_LP, "n": N_LP, "tol": TOL, "partition": "static block, %d LPs per GPU" % count,
                   "l2": "inputs per GPU (%.2f GB) exceed L2, no flush" % (count * M_LP * N_LP * 8 / 1e9),
                   "newton_iterations_per_step": int(it_all.sum()), "lockstep_iterations": int(it_all.max()),
                   "converged": n_conv, "converged_e2e": int(B - sums[3].item()),
                   "max_iteration_diff_vs_oracle_table": int(maxs[0].item()) if sums[4].item() == 0 else None,
                   "max_rel_objective_diff_vs_oracle_table": float(maxs[1].item()) if sums[4].item() == 0 else None,
                   "parity_table": "tests/golden/batch_256x512_oracle.npz (all %d LPs of all ranks, both arms)" % B
                                   if sums[4].item() == 0 else "not covered by the frozen table: convergence gate only",
                   "handed_to_augmented_system_kernel_rank0": int(handoffs),
                   "strong": strong},
        "parity": "FAILED" if parity_failed else "ok",
        "newton_it_per_s": float(it_all.sum()) * args.steps / t_dev,
        "wall_ms_per_step": wall_dev / args.steps * 1e3,
        "per_rank_ms_per_step": [round(v, 2) for v in per_rank_dev],
        "converged": n_conv, "iterations_min_max": [int(it_all.min()), int(it_all.max())],
        "e2e": {"value": B * args.steps / t_e2e, "unit": "LPs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(B) * 16, "ms_per_step": t_e2e / args.steps * 1e3,
                "per_rank_ms_per_step": [round(v, 2) for v in per_rank_e2e],
                "h2d_gbs_per_rank_if_copy_bound": h2d / world / (t_e2e / args.steps) * 1e-9,
                "host_numa_node_rank0": numa_node,
                "max_abs_diff_vs_device_arm": float(np.max(np.abs(obj2 - obj_all)))},
        "gpu_launches": int(sums[0].item()),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "setup": {"host_generation_s": gen_s},
    }
    if not args.no_extras and world == 1:
        del db
        torch.cuda.empty_cache()
        line["extras"] = {"netlib": extras_netlib(ipm, cpu_pool), "dense_big": extras_dense_big(ipm, lib, peak_tf)}
    if cpu_pool is not None:
        cpu_pool.close()
        cpu_pool.join()
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_failed:
        sys.stderr.write("bench.py: PARITY GATE FAILED (see config.converged / max_iteration_diff / max_rel_objective_diff)\n")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
