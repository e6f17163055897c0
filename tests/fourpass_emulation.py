"""CPU emulation of the batched solver's four-pass iteration on top of the oracle's building
blocks - TEST INFRASTRUCTURE, not product code.  It restates in numpy the two identities the CUDA path uses to
drop two of the six passes over A of the literal iteration (main.py:725-751):

  * corrector right-hand side by linearity of main.py:150-152 in r4:
        rhs_c = rhs_p + A (d * dxa * dsa / x) - sigma mu A (d / x),      d = x / s
  * residuals of the new point from the step instead of from scratch:
        rb += alpha_p A dx,        rc += alpha_d (A^T dy + ds)
    with a from-scratch evaluation (check_optimality, main.py:169-173) every `refresh_every` iterations and before
    an LP is declared finished.
"""
import numpy as np

from oracle import ipm_oracle as O


def solve_fourpass(A, b, c, tol=1e-8, refresh_every=3, max_iter=200):
    """Returns (k, x, y, s, history); history rows: (k, |rb| recurred, |rb| true, |rc| recurred, |rc| true,
    relative error of the linearity right-hand side against the direct one)."""
    m, n = A.shape
    b = b.reshape(-1, 1)
    c = c.reshape(-1, 1)
    x, y, s = O.initial_point(m, n, y0_is_one=False)          # dense driver start, main.py:287-302
    nb, nc = np.linalg.norm(b), np.linalg.norm(c)

    def cont(rb, rc):
        return bool((tol * (1 + nb) < np.linalg.norm(rb)) or (tol * (1 + nc) < np.linalg.norm(rc))
                    or (tol < float((x.T @ s)[0, 0])))

    rb, rc = O.residuals(A, b, c, x, y, s)
    fresh, k, hist = True, 0, []
    while k < max_iter:
        if not cont(rb, rc):
            if fresh:
                break
            rb, rc = O.residuals(A, b, c, x, y, s)             # stop candidates are re-checked from scratch
            fresh = True
            continue
        d = x / s
        L, _ = O.cholesky_safeguarded(O.normal_matrix(A, x, s))
        r3 = x * s
        tp = rc - r3 / x
        rhs_p = -rb - A @ (d * tp)                             # main.py:225
        dya = O.solve_with_factor(L, rhs_p)
        dxa = d * (A.T @ dya) + d * tp                         # main.py:227
        dsa = (-s * dxa / x) - (r3 / x)                        # main.py:228
        _, mu, sigma = O.sigma_mu(x, s, dxa, dsa)
        r4 = r3 + dxa * dsa - sigma * mu                       # main.py:150-152
        rhs_c = rhs_p + A @ (d * dxa * dsa / x) - sigma * mu * (A @ (d / x))
        rhs_direct = -rb - A @ (d * (rc - r4 / x))
        lin_err = float(np.linalg.norm(rhs_c - rhs_direct) / max(np.linalg.norm(rhs_direct), 1e-300))
        dy = O.solve_with_factor(L, rhs_c)
        u = A.T @ dy
        tc = rc - r4 / x
        dx = d * u + d * tc
        ds = (-s * dx / x) - (r4 / x)
        ap, ad = O.full_stepsize(x, s, dx, ds)
        x, y, s = x + ap * dx, y + ad * dy, s + ad * ds        # main.py:694-696
        rb = rb + ap * (A @ dx)
        rc = rc + ad * (u + ds)
        k += 1
        rbt, rct = O.residuals(A, b, c, x, y, s)
        hist.append((k, float(np.linalg.norm(rb)), float(np.linalg.norm(rbt)), float(np.linalg.norm(rc)),
                     float(np.linalg.norm(rct)), lin_err))
        fresh = refresh_every > 0 and k % refresh_every == 0
        if fresh:
            rb, rc = rbt, rct
    return k, x, y, s, hist


def solve_refined(A, b, c, tol=1e-8, max_iter=150):
    """Literal iteration on the normal equations with ONE step of iterative refinement of the corrector - what the
    batched solver runs for an LP after a straggler restart (kb_refine_rhs):
        delta = -rb - A dx;  M ddy = delta;  dy += ddy;  dx += d (A^T ddy);  ds = -s dx / x - r4 / x.
    Returns (k, objective)."""
    m, n = A.shape
    b = b.reshape(-1, 1)
    c = c.reshape(-1, 1)
    x, y, s = O.initial_point(m, n, y0_is_one=False)
    k = 0
    while O.continue_flag(A, b, c, x, y, s, tol, tol, tol) and k < max_iter:
        rb, rc = O.residuals(A, b, c, x, y, s)
        r3 = x * s
        L, _ = O.cholesky_safeguarded(O.normal_matrix(A, x, s))
        dxa, dya, dsa = O.direction_normal(A, L, x, s, rb, rc, r3)
        _, mu, sigma = O.sigma_mu(x, s, dxa, dsa)
        r4 = r3 + dxa * dsa - sigma * mu
        dx, dy, ds = O.direction_normal(A, L, x, s, rb, rc, r4)
        ddy = O.solve_with_factor(L, -rb - A @ dx)
        dy = dy + ddy
        dx = dx + (x / s) * (A.T @ ddy)
        ds = (-s * dx / x) - (r4 / x)
        ap, ad = O.full_stepsize(x, s, dx, ds)
        x, y, s = x + ap * dx, y + ad * dy, s + ad * ds
        k += 1
    return k, float((c.T @ x)[0, 0])
