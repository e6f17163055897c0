"""TEST INFRASTRUCTURE.  Seeded inputs of the step_size goldens (tests/golden/step_size_cases.npz holds the outputs of the
UNMODIFIED reference on exactly these inputs plus a checksum of each input set; oracle/make_golden_step_size.py)."""
import numpy as np


def cases():
    rng = np.random.default_rng(20261019)
    out = []
    for n in (1, 7, 64, 513, 4099):
        for flavour in ("generic", "all_dx_positive", "all_dx_negative", "small_steps", "inf_bounds", "ds_positive"):
            lb = rng.uniform(-2.0, 0.5, n)
            ub = lb + rng.uniform(0.5, 4.0, n)
            x = lb + rng.uniform(0.05, 0.95, n) * (ub - lb)
            x = np.maximum(x, 1e-3)                      # the cases without lb assume x > 0
            ub = np.maximum(ub, x + 1e-3)
            s = rng.uniform(0.1, 2.0, n)
            dx = rng.standard_normal(n) * 3.0
            ds = rng.standard_normal(n) * 3.0
            if flavour == "all_dx_positive":
                dx = np.abs(dx) + 0.1
            elif flavour == "all_dx_negative":
                dx = -np.abs(dx) - 0.1
            elif flavour == "small_steps":
                dx *= 1e-3
                ds *= 1e-3
            elif flavour == "inf_bounds":
                ub = np.where(rng.uniform(size=n) < 0.7, np.inf, ub)
                lb = np.where(rng.uniform(size=n) < 0.3, -np.inf, lb)
                if n <= 7:
                    ub[:] = np.inf
            elif flavour == "ds_positive":
                ds = np.abs(ds) + 0.1
            out.append((n, flavour, x, s, dx, ds, lb, ub))
    return out




def checksum(x, s, dx, ds, lb, ub):
    f = lambda v: float(np.sum(np.where(np.isfinite(v), v, 0.0)))      # noqa: E731
    return f(x) + 2 * f(s) + 3 * f(dx) + 5 * f(ds) + 7 * f(lb) + 11 * f(ub)
