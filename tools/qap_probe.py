"""QAP8 / QAP12 / QAP15 (rank-deficient A) through the opt-in treatments: python tools/qap_probe.py [names...]
dependent-row elimination (ipm_detect_dependent_rows) x conditional refinement (ipm_set_refinement) x start point."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import interiorpointmethod_b200 as ipm  # noqa: E402

tab = json.load(open(os.path.join(ROOT, "tests", "golden", "netlib_all.json")))["problems"]
names = [a for a in sys.argv[1:]] or ["QAP8", "QAP12", "QAP15"]
for name in names:
    A, b, c, cTlb = ipm.load_golden_problem(name)
    e = tab[name]
    target = e["highs"]["optimum"] if e["highs"].get("optimum") is not None else e["netlib_optimum"]
    with ipm.NewtonStep(A, b, c) as ns:
        for dep in (None, 1e-10):
            nd = ns.detect_dependent_rows(dep if dep else 0.0)
            for refine in (None, 1.0, 0.1):
                ns.set_refinement(refine)
                for start in ("reference", "mehrotra"):
                    t0 = time.perf_counter()
                    r = ns.solve(tol=1e-8, max_iter=300, cTlb=cTlb, start=start)
                    dt = time.perf_counter() - t0
                    x = np.asarray(r.x).ravel()
                    rb = np.linalg.norm(A @ x - np.asarray(b).ravel()) / (1e-8 * (1 + np.linalg.norm(b))) if np.isfinite(x).all() else np.nan
                    print("%-6s dep %-6s (%4d rows) refine %-5s start %-9s: %-9s k %3d obj %.12g rel %.2e rb/thr %.2e  %.2fs %.0f it/s"
                          % (name, dep, nd, refine, start, r.status, r.iterations, r.objective,
                             abs(r.objective - target) / max(1, abs(target)), rb, dt, r.iterations / dt), flush=True)
