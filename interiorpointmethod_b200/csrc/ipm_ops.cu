// Stateless op-level entry points on HOST vectors: the remaining pure-function seams of the reference
// (predicted_stepsize main.py:305-322, full_stepsize main.py:604-626, predicted/duality_gap main.py:562-601,
// corrected main.py:663-697, solve_linear main.py:176-182).  They exist for parity tests and for callers that keep
// the reference's own driver loop; the solve-level entry points never copy vectors per operation.
#include <vector>

#include "chol.cuh"
#include "potrf_auto.cuh"
#include "common.cuh"
#include "vec.cuh"

using namespace ipm;

namespace {

// min({-v_i/dv_i : dv_i < 0} U {1}) for two (v, dv) pairs at once -> out[0], out[1]
__global__ void k_ratio_pair(const double* x, const double* dx, const double* s, const double* ds, int n, double* out,
                             double* partials, unsigned* counter) {
    __shared__ double sh[32];
    double acc[2] = {1.0, 1.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double dxi = dx[i], dsi = ds[i];
        if (dxi < 0.0) acc[0] = fmin(acc[0], -x[i] / dxi);
        if (dsi < 0.0) acc[1] = fmin(acc[1], -s[i] / dsi);
    }
    double tot[2];
    if (grid_reduce<2, RED_MIN, RED_MIN>(acc, partials, counter, sh, tot) && threadIdx.x == 0) {
        out[0] = tot[0];
        out[1] = tot[1];
    }
}

// Bounded-variable ratio test (step_size, main.py:325-547): three minima and whether each set is empty
//   neg : min over dx_i < 0 of (lo_i - x_i)/dx_i   (lo = lb, or 0 when lb is null)
//   pos : min over dx_i > 0 of (ub_i - x_i)/dx_i   (only with ub)
//   dual: min over ds_i < 0 of -s_i/ds_i
// out[0..2] = minima (+inf when empty), out[3..5] = 1.0 where the set is not empty (the reference distinguishes an empty
// set from one whose ratios are all infinite: `min(...) if any(i) else 1`).
__global__ void k_ratio_bounded(const double* x, const double* dx, const double* s, const double* ds, const double* lb,
                                const double* ub, int n, double* out, double* partials, unsigned* counter) {
    __shared__ double sh[32];
    const double inf = red_identity<RED_MIN>();
    double mn[3] = {inf, inf, inf};
    double any[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double dxi = dx[i], dsi = ds[i], xi = x[i];
        if (dxi < 0.0) {
            const double lo = lb ? lb[i] : 0.0;
            mn[0] = fmin(mn[0], (lo - xi) / dxi);
            any[0] = 1.0;
        }
        if (ub && dxi > 0.0) {
            mn[1] = fmin(mn[1], (ub[i] - xi) / dxi);
            any[1] = 1.0;
        }
        if (dsi < 0.0) {
            mn[2] = fmin(mn[2], -s[i] / dsi);
            any[2] = 1.0;
        }
    }
    double tot[3];
    if (grid_reduce<3, RED_MIN, RED_MIN, RED_MIN>(mn, partials, counter, sh, tot) && threadIdx.x == 0) {
        out[0] = tot[0]; out[1] = tot[1]; out[2] = tot[2];
    }
    __syncthreads();
    if (grid_reduce<3, RED_MAX, RED_MAX, RED_MAX>(any, partials + 3 * VEC_MAX_BLOCKS, counter + 1, sh, tot) && threadIdx.x == 0) {
        out[3] = tot[0]; out[4] = tot[1]; out[5] = tot[2];
    }
}

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <typename T>
    int dev(T** p, size_t count) {
        IPM_CUDA_OK(cudaMalloc(p, std::max<size_t>(count, 1) * sizeof(T)));
        ptrs.push_back(*p);
        return IPM_OK;
    }
    int up(double** p, const double* h, size_t count) {
        IPM_TRY(dev(p, count));
        IPM_CUDA_OK(cudaMemcpy(*p, h, count * sizeof(double), cudaMemcpyHostToDevice));
        return IPM_OK;
    }
};

int reduction_scratch(Scratch& sc, double** partials, unsigned** counter) {
    IPM_TRY(sc.dev(partials, (size_t)VEC_MAX_BLOCKS * 4));
    IPM_TRY(sc.dev(counter, 1));
    IPM_CUDA_OK(cudaMemset(*counter, 0, sizeof(unsigned)));
    return IPM_OK;
}

}  // namespace

extern "C" {

int ipm_op_ratio_test(int device_ordinal, int n, const double* x, const double* dx, const double* s, const double* ds,
                      double eta, double alpha[2]) {
    if (!x || !dx || !s || !ds || !alpha) return IPM_ERR_ARG;
    if (n <= 0) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    Scratch sc;
    double *dxv, *ddx, *dsv, *dds, *out, *partials;
    unsigned* counter;
    IPM_TRY(sc.up(&dxv, x, n)); IPM_TRY(sc.up(&ddx, dx, n)); IPM_TRY(sc.up(&dsv, s, n)); IPM_TRY(sc.up(&dds, ds, n));
    IPM_TRY(sc.dev(&out, 2));
    IPM_TRY(reduction_scratch(sc, &partials, &counter));
    k_ratio_pair<<<vec_grid(n), VEC_NT>>>(dxv, ddx, dsv, dds, n, out, partials, counter);
    count_launch();
    IPM_TRY(launch_check());
    double h[2];
    IPM_CUDA_OK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    // eta <= 0: predicted_stepsize (main.py:305-322); eta > 0: full_stepsize min(1, eta*min(...)) (main.py:616-623)
    alpha[0] = eta > 0.0 ? fmin(1.0, eta * h[0]) : h[0];
    alpha[1] = eta > 0.0 ? fmin(1.0, eta * h[1]) : h[1];
    return IPM_OK;
}

int ipm_op_step_size_bounded(int device_ordinal, int n, const double* x, const double* dx, const double* s,
                             const double* ds, const double* lb, const double* ub, double eta, double alpha[2]) {
    if (!x || !dx || !s || !ds || !alpha) return IPM_ERR_ARG;
    if (n <= 0) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    Scratch sc;
    double *dxv, *ddx, *dsv, *dds, *dlb = nullptr, *dub = nullptr, *out, *partials;
    unsigned* counter;
    IPM_TRY(sc.up(&dxv, x, n)); IPM_TRY(sc.up(&ddx, dx, n)); IPM_TRY(sc.up(&dsv, s, n)); IPM_TRY(sc.up(&dds, ds, n));
    if (lb) IPM_TRY(sc.up(&dlb, lb, n));
    if (ub) IPM_TRY(sc.up(&dub, ub, n));
    IPM_TRY(sc.dev(&out, 6));
    IPM_TRY(sc.dev(&partials, (size_t)VEC_MAX_BLOCKS * 6));
    IPM_TRY(sc.dev(&counter, 2));
    IPM_CUDA_OK(cudaMemset(counter, 0, 2 * sizeof(unsigned)));
    k_ratio_bounded<<<vec_grid(n), VEC_NT>>>(dxv, ddx, dsv, dds, dlb, dub, n, out, partials, counter);
    count_launch();
    IPM_TRY(launch_check());
    double h[6];
    IPM_CUDA_OK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    // The case split of main.py:325-547, quirks included.  `one` = `min(np.append(ratios, 1)) if any else 1`,
    // `raw` = `min(ratios) if any else 1` (an all-infinite set stays infinite, an empty one gives 1).
    const bool corr = eta > 0.0;
    auto one = [&](int k) { return h[3 + k] > 0.5 ? fmin(h[k], 1.0) : 1.0; };
    auto raw = [&](int k) { return h[3 + k] > 0.5 ? h[k] : 1.0; };
    double ap, ad;
    if (!lb && !ub) {                       // main.py:328-341 (predictor), 441-455 (corrector)
        ap = corr ? fmin(1.0, eta * one(0)) : one(0);
        // corrector: the reference tests `delta_s_aff < 0` there, a name that is not bound in that branch; the bare
        // `except:` turns the error into alpha_dual = 1 (main.py:449-454)
        ad = corr ? 1.0 : one(2);
    } else if (!lb) {                       // 0 <= x <= ub: main.py:342-383, 456-493
        ap = corr ? fmin(fmin(eta * raw(0), eta * raw(1)), 1.0) : fmin(fmin(raw(0), raw(1)), 1.0);
        ad = corr ? fmin(eta * raw(2), 1.0) : fmin(raw(2), 1.0);
    } else if (!ub) {                       // lb <= x: main.py:384-413, 494-522 (only dx < 0 can hit a bound)
        ap = corr ? fmin(eta * raw(0), 1.0) : fmin(raw(0), 1.0);
        ad = corr ? fmin(eta * raw(2), 1.0) : fmin(raw(2), 1.0);
    } else {                                // lb <= x <= ub: main.py:414-438, 523-547
        ap = corr ? fmin(fmin(eta * one(1), eta * one(0)), 1.0) : fmin(fmin(one(1), one(0)), 1.0);
        ad = corr ? fmin(eta * one(2), 1.0) : one(2);
    }
    alpha[0] = ap;
    alpha[1] = ad;
    return IPM_OK;
}

int ipm_op_sigma(int device_ordinal, int n, const double* x, const double* s, const double* dx_aff,
                 const double* ds_aff, double out3[3]) {
    if (!x || !s || !dx_aff || !ds_aff || !out3) return IPM_ERR_ARG;
    if (n <= 0) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    Scratch sc;
    double *dxv, *dsv, *ddx, *dds, *scal, *partials, *rc, *d;
    unsigned* counter;
    IPM_TRY(sc.up(&dxv, x, n)); IPM_TRY(sc.up(&dsv, s, n)); IPM_TRY(sc.up(&ddx, dx_aff, n)); IPM_TRY(sc.up(&dds, ds_aff, n));
    IPM_TRY(sc.dev(&scal, S_COUNT)); IPM_TRY(sc.dev(&rc, n)); IPM_TRY(sc.dev(&d, n));
    IPM_CUDA_OK(cudaMemset(scal, 0, S_COUNT * sizeof(double)));
    IPM_TRY(reduction_scratch(sc, &partials, &counter));
    // step lengths of the predictor (main.py:305-322) -> S_AP_AFF/S_AD_AFF, x^T s -> S_XS, then main.py:582-600
    k_ratio_pair<<<vec_grid(n), VEC_NT>>>(dxv, ddx, dsv, dds, n, scal + S_AP_AFF, partials, counter);
    k_resid_dual<<<vec_grid(n), VEC_NT>>>(dxv /*unused A^T y*/, dsv, dxv, dxv, rc, d, n, 0.0, scal, partials, counter);
    k_sigma<<<vec_grid(n), VEC_NT>>>(dxv, dsv, ddx, dds, n, scal, partials, counter);
    count_launch(3);
    IPM_TRY(launch_check());
    double h[S_COUNT];
    IPM_CUDA_OK(cudaMemcpy(h, scal, sizeof(h), cudaMemcpyDeviceToHost));
    out3[0] = h[S_MU_AFF]; out3[1] = h[S_MU]; out3[2] = h[S_SIGMA];
    return IPM_OK;
}

int ipm_op_update(int device_ordinal, int m, int n, double* x, double* y, double* s, const double* dx,
                  const double* dy, const double* ds, double alpha_p, double alpha_d) {
    if (!x || !y || !s || !dx || !dy || !ds) return IPM_ERR_ARG;
    if (m <= 0 || n <= 0) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    Scratch sc;
    double *vx, *vy, *vs, *vdx, *vdy, *vds;
    IPM_TRY(sc.up(&vx, x, n)); IPM_TRY(sc.up(&vy, y, m)); IPM_TRY(sc.up(&vs, s, n));
    IPM_TRY(sc.up(&vdx, dx, n)); IPM_TRY(sc.up(&vdy, dy, m)); IPM_TRY(sc.up(&vds, ds, n));
    k_update<<<vec_grid(m > n ? m : n), VEC_NT>>>(vx, vy, vs, vdx, vdy, vds, m, n, nullptr, alpha_p, alpha_d);
    count_launch();
    IPM_TRY(launch_check());
    IPM_CUDA_OK(cudaMemcpy(x, vx, n * sizeof(double), cudaMemcpyDeviceToHost));
    IPM_CUDA_OK(cudaMemcpy(y, vy, m * sizeof(double), cudaMemcpyDeviceToHost));
    IPM_CUDA_OK(cudaMemcpy(s, vs, n * sizeof(double), cudaMemcpyDeviceToHost));
    return IPM_OK;
}

// Solve M z = rhs for a symmetric positive (semi)definite dense host matrix (lower triangle read) with the
// safeguarded Cholesky and the triangular sweeps: the GPU counterpart of solve_linear (main.py:176-182) on the
// normal-equations matrix of main.py:226.
int ipm_solve_spd(int device_ordinal, int m, const double* M_rowmajor, const double* rhs, double pivot_rel_thresh,
                  double* z, int* n_fixed) {
    if (!M_rowmajor || !rhs || !z) return IPM_ERR_ARG;
    if (m <= 0) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    Scratch sc;
    const int64_t ldm = round_up(m, 16);
    double *dM, *dr, *dt, *dz, *scal;
    IPM_TRY(sc.dev(&dM, (size_t)m * ldm)); IPM_TRY(sc.up(&dr, rhs, m)); IPM_TRY(sc.dev(&dt, m)); IPM_TRY(sc.dev(&dz, m));
    IPM_TRY(sc.dev(&scal, S_COUNT));
    IPM_CUDA_OK(cudaMemcpy2D(dM, ldm * sizeof(double), M_rowmajor, (size_t)m * sizeof(double),
                             (size_t)m * sizeof(double), m, cudaMemcpyHostToDevice));
    IPM_TRY((potrf_single_auto(dM, ldm, m, scal, pivot_rel_thresh, 0)));
    IPM_TRY(potrs_single(dM, ldm, m, dr, dt, dz, 0));
    IPM_CUDA_OK(cudaMemcpy(z, dz, m * sizeof(double), cudaMemcpyDeviceToHost));
    if (n_fixed) {
        double h[S_COUNT];
        IPM_CUDA_OK(cudaMemcpy(h, scal, sizeof(h), cudaMemcpyDeviceToHost));
        *n_fixed = (int)h[S_NFIXED];
    }
    return IPM_OK;
}

}  // extern "C"
