"""Workload for the HBM-rate captures of the single-LP elementwise / sparse kernels (north star: "achieved HBM GB/s for
the sparse and elementwise kernels"; SURVEY 7.1 gate: >= 80 % of the measured copy bandwidth at n >= 2^24):
    python tools/elementwise_probe.py [log2_n] [m] [iterations]
A sparse LP with n = 2^24 columns, one entry per column (row j mod m), so that every vector kernel of the iteration
(k_resid_dual, k_make_w, k_direction, k_sigma, k_update) streams n-vectors from HBM and the SpMV / SpGEMM kernels
stream 16.7 M entries; M is diagonal.  Run it under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` and feed the csv to tools/kernel_rates.py."""
import sys
import time

import numpy as np
from scipy import sparse

import interiorpointmethod_b200 as ipm

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
its = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << lg
rng = np.random.default_rng(0)
j = np.arange(n, dtype=np.int64)
data = 1.0 + 0.25 * (j % 5)
A = sparse.csc_matrix((data, (j % m).astype(np.int32), np.arange(n + 1, dtype=np.int64)), shape=(m, n))
xh = rng.uniform(0.1, 1.1, n)
sh = rng.uniform(0.1, 1.1, n)
yh = rng.standard_normal(m)
b = A @ xh
c = A.T @ yh + sh
t0 = time.perf_counter()
with ipm.NewtonStep(A, b, c) as ns:
    t1 = time.perf_counter()
    r = ns.solve(tol=1e-8, max_iter=its)
    t2 = time.perf_counter()
print("n = 2^%d, m = %d, nnz = %d: load %.2f s, %d iterations in %.3f s, status %s" % (lg, m, A.nnz, t1 - t0, r.iterations,
                                                                                      t2 - t1, r.status))
