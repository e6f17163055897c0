// Fused elementwise + reduction kernels of one predictor-corrector iteration (single LP).
// Each kernel names the reference lines it replaces.  Reductions are deterministic: fixed grid for a given
// length, fixed shuffle trees, partials combined in index order by the last CTA to finish.
#pragma once
#include "common.cuh"

namespace ipm {

constexpr int VEC_NT = 256;
constexpr int VEC_MAX_BLOCKS = 1184;     // 8 x 148

inline int vec_grid(int64_t len) {
    int64_t b = (len + VEC_NT * 2 - 1) / (VEC_NT * 2);
    if (b < 1) b = 1;
    if (b > VEC_MAX_BLOCKS) b = VEC_MAX_BLOCKS;
    return (int)b;
}

#ifdef __CUDACC__
// The streaming kernels below walk their vectors VEC_U elements per thread and trip: all loads of a trip are issued
// before the first division (FP64 divisions are long dependent chains; with one element per trip the next loads
// waited behind them and the kernels ran at 47-60 % of the HBM rate, ncu round 2).  Element-to-thread mapping and the
// order in which a thread accumulates are unchanged, so every result is bitwise what the one-element loop gave.
constexpr int VEC_U = 4;

// |v|_2 -> *out  (used once per problem for |b|, |c|: main.py:169-170)
static __global__ void k_norm2(const double* v, int len, double* out, double* partials, unsigned* counter) {
    __shared__ double sh[32];
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        acc[0] += v[i] * v[i];
    double tot[1];
    if (grid_reduce<1, RED_SUM>(acc, partials, counter, sh, tot) && threadIdx.x == 0) *out = sqrt(tot[0]);
}

// rb = Ax - b and |rb|^2   (main.py:67, 169).  Ax comes from the mat-vec kernel.
static __device__ __forceinline__ void d_resid_primal(const double* Ax, const double* b, double* rb, int m, double* scal, double* partials,
                               unsigned* counter) {
    __shared__ double sh[32];
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double r = Ax[i] - b[i];
        rb[i] = r;
        acc[0] += r * r;
    }
    double tot[1];
    if (grid_reduce<1, RED_SUM>(acc, partials, counter, sh, tot) && threadIdx.x == 0) {
        scal[S_NRB2] = tot[0];
        scal[S_NRB] = sqrt(tot[0]);
    }
}
static __global__ void k_resid_primal(const double* Ax, const double* b, double* rb, int m, double* scal, double* partials,
                               unsigned* counter) { d_resid_primal(Ax, b, rb, m, scal, partials, counter); }

// rc = A^T y + s - c, |rc|^2, x^T s, c^T x, d = x/s   (main.py:70, 170-172, 223, 815) and the continue flag
// of check_optimality (main.py:169-173): strict '<', NaN => stop.  Runs after k_resid_primal on the same stream.
static __device__ __forceinline__ void d_resid_dual(const double* ATy, const double* s, const double* c, const double* x, double* rc,
                             double* d, int n, double tol, double* scal, double* partials, unsigned* counter) {
    __shared__ double sh[32];
    double acc[3] = {0.0, 0.0, 0.0};
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + (VEC_U - 1) * T < n; i += VEC_U * T) {
        double xv[VEC_U], sv[VEC_U], cv[VEC_U], av[VEC_U];
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) { xv[u] = x[i + u * T]; sv[u] = s[i + u * T]; cv[u] = c[i + u * T]; av[u] = ATy[i + u * T]; }
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            const double r = av[u] + sv[u] - cv[u];
            rc[i + u * T] = r;
            d[i + u * T] = xv[u] / sv[u];
            acc[0] += r * r;
            acc[1] += xv[u] * sv[u];
            acc[2] += xv[u] * cv[u];
        }
    }
    for (; i < n; i += T) {
        const double xi = x[i], si = s[i], ci = c[i];
        const double r = ATy[i] + si - ci;
        rc[i] = r;
        d[i] = xi / si;
        acc[0] += r * r;
        acc[1] += xi * si;
        acc[2] += xi * ci;
    }
    double tot[3];
    if (grid_reduce<3, RED_SUM, RED_SUM, RED_SUM>(acc, partials, counter, sh, tot) && threadIdx.x == 0) {
        scal[S_NRC2] = tot[0];
        scal[S_NRC] = sqrt(tot[0]);
        scal[S_XS] = tot[1];
        scal[S_OBJ] = tot[2];
        const double nrb = scal[S_NRB], nrc = sqrt(tot[0]);
        const bool primal = tol * (1.0 + scal[S_NB]) < nrb;
        const bool dual = tol * (1.0 + scal[S_NC]) < nrc;
        const bool gap = tol < tot[1];
        scal[S_CONT] = (primal || dual || gap) ? 1.0 : 0.0;
    }
}
static __global__ void k_resid_dual(const double* ATy, const double* s, const double* c, const double* x, double* rc,
                             double* d, int n, double tol, double* scal, double* partials, unsigned* counter) { d_resid_dual(ATy, s, c, x, rc, d, n, tol, scal, partials, counter); }

// Complementarity right-hand side and the eliminated vector of the normal equations:
//   kind 0: rcomp = x*s                                  (main.py:72)
//   kind 1: rcomp = x*s + dxa*dsa - sigma*mu             (main.py:150-152)
//   rcx = rcomp / x ;  w = d * (rc - rcx)                (main.py:225: D^2 (r1 - r3/x))
static __device__ __forceinline__ void d_make_w(int kind, const double* x, const double* s, const double* rc, const double* d,
                         const double* dxa, const double* dsa, const double* scal, double* rcx, double* w, int n) {
    const double sigma_mu = kind ? scal[S_SIGMA_MU] : 0.0;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + (VEC_U - 1) * T < n; i += VEC_U * T) {
        double xv[VEC_U], rcomp[VEC_U], dv[VEC_U], rv[VEC_U];
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            xv[u] = x[i + u * T];
            rcomp[u] = xv[u] * s[i + u * T];
            if (kind) rcomp[u] = rcomp[u] + dxa[i + u * T] * dsa[i + u * T] - sigma_mu;
            dv[u] = d[i + u * T];
            rv[u] = rc[i + u * T];
        }
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            const double q = rcomp[u] / xv[u];
            rcx[i + u * T] = q;
            w[i + u * T] = dv[u] * (rv[u] - q);
        }
    }
    for (; i < n; i += T) {
        const double xi = x[i];
        double rcomp = xi * s[i];
        if (kind) rcomp = rcomp + dxa[i] * dsa[i] - sigma_mu;
        const double q = rcomp / xi;
        rcx[i] = q;
        w[i] = d[i] * (rc[i] - q);
    }
}
static __global__ void k_make_w(int kind, const double* x, const double* s, const double* rc, const double* d,
                         const double* dxa, const double* dsa, const double* scal, double* rcx, double* w, int n) { d_make_w(kind, x, s, rc, d, dxa, dsa, scal, rcx, w, n); }

// rhs = -rb - A w   (main.py:225)
static __device__ __forceinline__ void d_make_rhs(const double* rb, const double* Aw, double* rhs, int m) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        rhs[i] = -rb[i] - Aw[i];
}
static __global__ void k_make_rhs(const double* rb, const double* Aw, double* rhs, int m) { d_make_rhs(rb, Aw, rhs, m); }

// dx = d*(A^T dy) + w ; ds = -s*dx/x - rcx   (main.py:227-228) fused with the ratio test
//   alpha = min({-v_i/dv_i : dv_i < 0} U {1})            (main.py:308-319)
//   kind 1 additionally alpha = min(1, eta*alpha)        (main.py:616-623)
static __device__ __forceinline__ void d_direction(int kind, const double* ATdy, const double* d, const double* w, const double* rcx,
                            const double* x, const double* s, double* dx, double* ds, int n, double eta, double* scal,
                            double* partials, unsigned* counter) {
    __shared__ double sh[32];
    double acc[2] = {1.0, 1.0};
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + (VEC_U - 1) * T < n; i += VEC_U * T) {
        double xv[VEC_U], sv[VEC_U], dxv[VEC_U], qv[VEC_U];
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            xv[u] = x[i + u * T]; sv[u] = s[i + u * T];
            dxv[u] = d[i + u * T] * ATdy[i + u * T] + w[i + u * T];
            qv[u] = rcx[i + u * T];
        }
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            const double dxi = dxv[u];
            const double dsi = (-sv[u] * dxi / xv[u]) - qv[u];
            dx[i + u * T] = dxi;
            ds[i + u * T] = dsi;
            if (dxi < 0.0) acc[0] = fmin(acc[0], -xv[u] / dxi);
            if (dsi < 0.0) acc[1] = fmin(acc[1], -sv[u] / dsi);
        }
    }
    for (; i < n; i += T) {
        const double xi = x[i], si = s[i];
        const double dxi = d[i] * ATdy[i] + w[i];
        const double dsi = (-si * dxi / xi) - rcx[i];
        dx[i] = dxi;
        ds[i] = dsi;
        if (dxi < 0.0) acc[0] = fmin(acc[0], -xi / dxi);
        if (dsi < 0.0) acc[1] = fmin(acc[1], -si / dsi);
    }
    double tot[2];
    if (grid_reduce<2, RED_MIN, RED_MIN>(acc, partials, counter, sh, tot) && threadIdx.x == 0) {
        scal[S_RAW_P] = tot[0];
        scal[S_RAW_D] = tot[1];
        if (kind == 0) {
            scal[S_AP_AFF] = tot[0];
            scal[S_AD_AFF] = tot[1];
        } else {
            scal[S_AP] = fmin(1.0, eta * tot[0]);
            scal[S_AD] = fmin(1.0, eta * tot[1]);
        }
    }
}
static __global__ void k_direction(int kind, const double* ATdy, const double* d, const double* w, const double* rcx,
                            const double* x, const double* s, double* dx, double* ds, int n, double eta, double* scal,
                            double* partials, unsigned* counter) { d_direction(kind, ATdy, d, w, rcx, x, s, dx, ds, n, eta, scal, partials, counter); }

// mu_aff = (x + ap dxa)^T (s + ad dsa)/n ; mu = x^T s/n ; sigma = (mu_aff/mu)^3   (main.py:582-584, 598-600)
static __device__ __forceinline__ void d_sigma(const double* x, const double* s, const double* dxa, const double* dsa, int n, double* scal,
                        double* partials, unsigned* counter) {
    __shared__ double sh[32];
    const double ap = scal[S_AP_AFF], ad = scal[S_AD_AFF];
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc[0] += (x[i] + ap * dxa[i]) * (s[i] + ad * dsa[i]);
    double tot[1];
    if (grid_reduce<1, RED_SUM>(acc, partials, counter, sh, tot) && threadIdx.x == 0) {
        const double mu_aff = tot[0] / (double)n;
        const double mu = scal[S_XS] / (double)n;
        const double r = mu_aff / mu;
        const double sigma = r * r * r;
        scal[S_MU_AFF] = mu_aff;
        scal[S_MU] = mu;
        scal[S_SIGMA] = sigma;
        scal[S_SIGMA_MU] = sigma * mu;
    }
}
static __global__ void k_sigma(const double* x, const double* s, const double* dxa, const double* dsa, int n, double* scal,
                        double* partials, unsigned* counter) { d_sigma(x, s, dxa, dsa, n, scal, partials, counter); }

// x += ap dx ; s += ad ds ; y += ad dy   (main.py:694-696).  alpha read from scal unless overridden (>= 0).
static __device__ __forceinline__ void d_update(double* x, double* y, double* s, const double* dx, const double* dy, const double* ds, int m,
                         int n, const double* scal, double ap_override, double ad_override) {
    const double ap = ap_override >= 0.0 ? ap_override : scal[S_AP];
    const double ad = ad_override >= 0.0 ? ad_override : scal[S_AD];
    const int len = m > n ? m : n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) {
            x[i] = x[i] + ap * dx[i];
            s[i] = s[i] + ad * ds[i];
        }
        if (i < m) y[i] = y[i] + ad * dy[i];
    }
}
static __global__ void k_update(double* x, double* y, double* s, const double* dx, const double* dy, const double* ds, int m,
                         int n, const double* scal, double ap_override, double ad_override) { d_update(x, y, s, dx, dy, ds, m, n, scal, ap_override, ad_override); }

// Conditional refinement of the corrector, single LP (same rule as the batched kbf_dir, ipm_batched_fused.cuh):
// delta = -rb - A dx; flag = |delta| > thresh |rb| (NaN compares false).  The solve M ddy = delta runs
// unconditionally (no host round trip inside the captured iteration), k_add_if applies it when the flag is set.
static __global__ void k_refine_delta(const double* rb, const double* Adx, double* delta, int m, double thresh, double* scal,
                                      double* partials, unsigned* counter) {
    __shared__ double sh[32];
    double acc[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double r = rb[i], dl = -r - Adx[i];
        delta[i] = dl;
        acc[0] += dl * dl;
        acc[1] += r * r;
    }
    double tot[2];
    if (grid_reduce<2, RED_SUM, RED_SUM>(acc, partials, counter, sh, tot) && threadIdx.x == 0) {
        const bool go = tot[0] > thresh * thresh * tot[1];
        scal[S_REFINE_FLAG] = go ? 1.0 : 0.0;
        if (go) scal[S_NREFINE] = scal[S_NREFINE] + 1.0;
    }
}
static __global__ void k_add_if(double* v, const double* add, int len, const double* flag) {
    if (!(*flag > 0.5)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = v[i] + add[i];
}

static __global__ void k_fill(double* v, int64_t len, double val) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = val;
}
// Starting point heuristic of Mehrotra (SIAM J. Optim. 2 (1992), sec. 7) - NOT in the reference, opt-in
// (SURVEY.md 8(f) row 4).  On entry x = A^T (A A^T)^-1 b and s = c - A^T y with y = (A A^T)^-1 A c; one CTA shifts
// both into the positive orthant:  dx = max(-1.5 min x, 0), ds likewise, then the second shift
// 0.5 (x^T s) / sum(s) resp. / sum(x) that balances the complementarity products.
static __global__ void __launch_bounds__(1024) k_mehrotra_shift(double* x, double* s, int n) {
    __shared__ double sh[32];
    __shared__ double bc[2];
    const int tid = threadIdx.x;
    double mx = red_identity<RED_MIN>(), ms = red_identity<RED_MIN>();
    for (int i = tid; i < n; i += blockDim.x) { mx = fmin(mx, x[i]); ms = fmin(ms, s[i]); }
    mx = block_red<RED_MIN>(mx, sh);
    if (tid == 0) bc[0] = fmax(-1.5 * mx, 0.0);
    ms = block_red<RED_MIN>(ms, sh);
    if (tid == 0) bc[1] = fmax(-1.5 * ms, 0.0);
    __syncthreads();
    const double dx = bc[0], ds = bc[1];
    double xs = 0.0, sx = 0.0, ss = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double xi = x[i] + dx, si = s[i] + ds;
        xs += xi * si; sx += xi; ss += si;
    }
    __shared__ double tot[3];
    xs = block_red<RED_SUM>(xs, sh);
    if (tid == 0) tot[0] = xs;
    sx = block_red<RED_SUM>(sx, sh);
    if (tid == 0) tot[1] = sx;
    ss = block_red<RED_SUM>(ss, sh);
    if (tid == 0) {
        // second shift: x gets 0.5 x^T s / sum(s), s gets 0.5 x^T s / sum(x)
        bc[0] = dx + 0.5 * tot[0] / fmax(ss, 1e-300);
        bc[1] = ds + 0.5 * tot[0] / fmax(tot[1], 1e-300);
    }
    __syncthreads();
    const double tx = bc[0], ts = bc[1];
    for (int i = tid; i < n; i += blockDim.x) {
        x[i] = fmax(x[i] + tx, 1e-10);
        s[i] = fmax(s[i] + ts, 1e-10);
    }
}
// out = c - v  (dual slack of the starting point)
static __global__ void k_sub(const double* c, const double* v, double* out, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = c[i] - v[i];
}

#endif

}  // namespace ipm
