"""TEST INFRASTRUCTURE.  Freezes outputs of the UNMODIFIED reference's `step_size` (main.py:325-547; reached through
predicted_stepsize_lb_ub / full_stepsize_lb_ub, main.py:550-559, 629-660) on seeded inputs, for the four bound
configurations x predictor / corrector, with the edge cases the case split distinguishes: empty index sets, ratios
above 1, infinite bounds.  Run in the build container (needs /root/reference):

    python oracle/make_golden_step_size.py        ->  tests/golden/step_size_cases.npz
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from oracle.step_size_cases import cases, checksum  # noqa: E402

ref_main, _ = ref_harness.load_reference()


def main():
    outs, sums = [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for (n, flavour, x, s, dx, ds, lb, ub) in cases():
            col = lambda v: v.reshape(-1, 1).copy()          # noqa: E731  the reference works on (n,1) columns
            for bounds in ("none", "ub", "lb", "both"):
                L = col(lb) if bounds in ("lb", "both") else None
                U = col(ub) if bounds in ("ub", "both") else None
                with np.errstate(all="ignore"):
                    ap0, ad0 = ref_main.predicted_stepsize_lb_ub(col(dx), None, col(ds), col(x), col(s), L, U)
                    ap1, ad1 = ref_main.full_stepsize_lb_ub(col(x), None, col(s), col(dx), None, col(ds), None, None, None, L, U)
                outs.append([float(ap0), float(ad0), float(ap1), float(ad1)])
            sums.append(checksum(x, s, dx, ds, lb, ub))
    path = os.path.join(ROOT, "tests", "golden", "step_size_cases.npz")
    # out[case][bounds in (none, ub, lb, both)] = (alpha_p, alpha_d) predictor, (alpha_p, alpha_d) corrector
    np.savez_compressed(path, out=np.array(outs).reshape(len(sums), 4, 4), input_checksum=np.array(sums))
    print("wrote", path, len(sums), "input sets x 4 bound configurations,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
