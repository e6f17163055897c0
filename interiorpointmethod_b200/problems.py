"""Problem sources for the hot path: the reference's `.mat` standard-form files, the frozen `.npz` copies that
travel to the GPU box, and the synthetic dense generator named in BASELINE.json."""
from __future__ import annotations

import os

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_PROBLEMS = os.path.join(_REPO, "tests", "golden", "problems")


def create_problem_from_mps(name, root="benchmarks"):
    """Same contract as sparse_interior.create_problem_from_mps (sparse_interior.py:139-167, 211-216):
    reads `<root>/<name>.mat` {f, b, cTlb, A{i,j,k}, num_variables, num_constraints} and returns
    (A csc_matrix, b (m,1), c (n,1), cTlb).  Unlike the reference the path is not cwd-relative-only and the
    shape is taken from num_constraints x num_variables instead of being inferred from the largest index."""
    from scipy import sparse
    from scipy.io import loadmat

    path = os.path.join(root, name if name.endswith(".mat") else name + ".mat")
    data = loadmat(path)
    i = data["A"]["i"][0][0][0].astype(np.int64)
    j = data["A"]["j"][0][0][0].astype(np.int64)
    k = data["A"]["k"][0][0][0].astype(np.float64)
    n = int(data["num_variables"][0][0])
    m = int(data["num_constraints"][0][0])
    A = sparse.csc_matrix((k, (i, j)), shape=(m, n))
    b = np.asarray(data["b"], dtype=np.float64).reshape(-1, 1)
    c = np.asarray(data["f"], dtype=np.float64).reshape(-1, 1)
    return A, b, c, float(data["cTlb"][0][0])


def load_golden_problem(name):
    """Frozen copy of the reference loader output (written by the golden-vector script, see tests/golden/)."""
    from scipy import sparse

    z = np.load(os.path.join(GOLDEN_PROBLEMS, name + ".npz"))
    m, n = int(z["m"]), int(z["n"])
    A = sparse.csc_matrix((z["data"], z["indices"], z["indptr"]), shape=(m, n))
    return A, z["b"].reshape(-1, 1), z["c"].reshape(-1, 1), float(z["cTlb"])


def synthetic_dense_lp(m, n, seed):
    """Strictly primal-dual feasible dense LP: A ~ N(0,1), b = A x^, c = A^T y^ + s^."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    xh = rng.uniform(0.1, 1.1, n)
    sh = rng.uniform(0.1, 1.1, n)
    yh = rng.standard_normal(m)
    return A, A @ xh, A.T @ yh + sh


def synthetic_dense_batch(first, count, m, n, out_A=None, out_b=None, out_c=None, threads=8):
    """LPs first..first+count-1 of the batch workload (LP i uses default_rng(i)) as contiguous
    A[count][m][n], b[count][m], c[count][n]."""
    from concurrent.futures import ThreadPoolExecutor

    A = out_A if out_A is not None else np.empty((count, m, n))
    b = out_b if out_b is not None else np.empty((count, m))
    c = out_c if out_c is not None else np.empty((count, n))

    def one(i):
        rng = np.random.default_rng(first + i)
        rng.standard_normal((m, n), out=A[i])
        xh = rng.uniform(0.1, 1.1, n)
        sh = rng.uniform(0.1, 1.1, n)
        yh = rng.standard_normal(m)
        np.matmul(A[i], xh, out=b[i])
        np.add(A[i].T @ yh, sh, out=c[i])

    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(count)))
    else:
        for i in range(count):
            one(i)
    return A, b, c
