"""A few Newton iterations of one golden LP (for ncu launch lists): python tools/one_lp.py NAME [iterations]"""
import sys

import interiorpointmethod_b200 as ipm

name = sys.argv[1] if len(sys.argv) > 1 else "QAP15"
its = int(sys.argv[2]) if len(sys.argv) > 2 else 3
A, b, c, cT = ipm.load_golden_problem(name)
with ipm.NewtonStep(A, b, c) as ns:
    r = ns.solve(tol=1e-8, max_iter=its, cTlb=cT)
    print(name, r.status, r.iterations, r.objective)
