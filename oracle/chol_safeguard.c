/* TEST INFRASTRUCTURE — CPU oracle helper, NOT product code.
 *
 * Safeguarded Cholesky used by oracle/ipm_oracle.py for matrices in the thousands.
 * Restates the factorisation that replaces the reference's linear-solve seam
 * `solve_linear` (main.py:176-182) on the normal-equations matrix of main.py:223-224,
 * with the tiny-pivot rule of SURVEY.md App. A.4:
 *     pivot p_j <= tau * max_i M_ii (or NaN)  ->  p_j = big.
 *
 * Row-major, in place, lower triangle; the strict upper triangle is zeroed on exit.
 * Left-looking by block column so the inner loops are contiguous dot products.
 */
#include <math.h>
#include <stddef.h>

#define NB 48

/* forced (nullable, m bytes): rows whose pivot is replaced whatever its value - the dependent rows of a
 * rank-deficient A found once by ipm_detect_dependent_rows (include/ipm_b200.h); replaced (nullable, m bytes): out,
 * 1 where the pivot was replaced (by the tiny-pivot rule or because it was forced). */
int oracle_chol_safeguard_masked(double *M, int m, double tau, double big, const unsigned char *forced,
                                 unsigned char *replaced, int *nfixed_out);

int oracle_chol_safeguard(double *M, int m, double tau, double big, int *nfixed_out)
{
    return oracle_chol_safeguard_masked(M, m, tau, big, NULL, NULL, nfixed_out);
}

int oracle_chol_safeguard_masked(double *M, int m, double tau, double big, const unsigned char *forced,
                                 unsigned char *replaced, int *nfixed_out)
{
    if (!M || m < 0) return -1;
    double maxdiag = -INFINITY;
    for (int i = 0; i < m; ++i) {
        double v = M[(size_t)i * m + i];
        if (v > maxdiag) maxdiag = v;            /* NaN diagonal entries are ignored like np.max would not; see below */
    }
    for (int i = 0; i < m; ++i) {                /* np.max propagates NaN: keep that behaviour */
        double v = M[(size_t)i * m + i];
        if (v != v) { maxdiag = v; break; }
    }
    const double thresh = tau * maxdiag;
    int nfixed = 0;

    for (int j0 = 0; j0 < m; j0 += NB) {
        const int j1 = (j0 + NB < m) ? j0 + NB : m;
        /* 1. left-looking update of block column [j0,j1) for all rows i >= j0 */
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = j0; i < m; ++i) {
            double *Li = M + (size_t)i * m;
            const int jmax = (i + 1 < j1) ? i + 1 : j1;
            for (int j = j0; j < jmax; ++j) {
                const double *Lj = M + (size_t)j * m;
                double acc = 0.0;
                for (int k = 0; k < j0; ++k) acc += Li[k] * Lj[k];
                Li[j] -= acc;
            }
        }
        /* 2. factor the diagonal block, unblocked, with the safeguard */
        for (int j = j0; j < j1; ++j) {
            double *Lj = M + (size_t)j * m;
            double p = Lj[j];
            for (int k = j0; k < j; ++k) p -= Lj[k] * Lj[k];
            const int bad = !(p > thresh) || (forced && forced[j]);
            if (bad) { p = big; ++nfixed; }
            if (replaced) replaced[j] = (unsigned char)bad;
            const double ljj = sqrt(p);
            Lj[j] = ljj;
            for (int i = j + 1; i < j1; ++i) {
                double *Li = M + (size_t)i * m;
                double v = Li[j];
                for (int k = j0; k < j; ++k) v -= Li[k] * Lj[k];
                Li[j] = v / ljj;
            }
        }
        /* 3. triangular solve of the rows below the block */
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = j1; i < m; ++i) {
            double *Li = M + (size_t)i * m;
            for (int j = j0; j < j1; ++j) {
                const double *Lj = M + (size_t)j * m;
                double v = Li[j];
                for (int k = j0; k < j; ++k) v -= Li[k] * Lj[k];
                Li[j] = v / Lj[j];
            }
        }
    }
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j) M[(size_t)i * m + j] = 0.0;
    if (nfixed_out) *nfixed_out = nfixed;
    return 0;
}
