"""Every LP of the reference's benchmarks/ (frozen: tests/golden/problems, pinned: tests/golden/netlib_all.json) through
the GPU path:  python tools/netlib_sweep.py [out.json] [--only NAME ...]
reference start (x = s = 1, y = 1; cap 5000 / 400 for m > 2500) where the unmodified reference converges, Mehrotra
start (cap 500) for all; objective against the reference's / HiGHS's, host-recomputed residuals."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import interiorpointmethod_b200 as ipm  # noqa: E402

tab = json.load(open(os.path.join(ROOT, "tests", "golden", "netlib_all.json")))["problems"]
only = [a for a in sys.argv[1:] if not a.startswith("--") and not a.endswith(".json")]
outp = next((a for a in sys.argv[1:] if a.endswith(".json")), None)
res = {}
for name, ent in tab.items():
    if only and name not in only:
        continue
    A, b, c, cTlb = ipm.load_golden_problem(name)
    row = dict(m=ent["m"], n=ent["n"], finite=ent["finite"])
    target = None
    if ent.get("reference"):
        target = ent["reference"]["obj"]
    elif ent.get("highs", {}).get("optimum") is not None:
        target = ent["highs"]["optimum"]
    row["target"] = target
    try:
        with ipm.NewtonStep(A, b, c) as ns:
            for start, cap in (("reference", 400 if ent["m"] > 2500 else 5000), ("mehrotra", 500)):
                if start == "reference" and not ent.get("reference") and ent["finite"]:
                    continue
                t0 = time.perf_counter()
                r = ns.solve(tol=1e-8, max_iter=cap, cTlb=0.0 if not ent["finite"] else cTlb, start=start)
                dt = time.perf_counter() - t0
                x = np.asarray(r.x).ravel()
                d = dict(status=r.status, k=int(r.iterations), obj=float(r.objective), s=round(dt, 4))
                if ent["finite"] and np.isfinite(x).all():
                    rb = A @ x - np.asarray(b).ravel()
                    d["rb_over_thresh"] = float(np.linalg.norm(rb) / (1e-8 * (1 + np.linalg.norm(b))))
                    d["min_x"] = float(x.min())
                if target is not None and np.isfinite(r.objective):
                    d["rel_err"] = float(abs(r.objective - target) / max(1.0, abs(target)))
                row[start] = d
    except Exception as e:
        row["error"] = str(e)[:200]
    res[name] = row
    print(name, json.dumps(row), flush=True)
if outp:
    json.dump(res, open(outp, "w"), indent=1)
