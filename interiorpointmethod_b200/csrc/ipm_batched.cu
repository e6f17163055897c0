// C ABI, batched part: B independent dense LPs of one shape solved in lockstep on one GPU
// (BASELINE.json config "batch of 8192 synthetic dense LPs m=256, n=512").  Every kernel takes the LP index
// from the grid and skips LPs whose `active` flag is 0, so LPs that converge early stop costing anything.
//
// One predictor-corrector iteration (main.py:725-751, normal-equations elimination main.py:221-229) is
//   kb_residual            one pass over A: A x, A^T y, rb, rc, d, norms, continue flag     (main.py:67-70,169-173)
//   dmma_nt_kernel         M = A diag(d) A^T, lower tiles, FP64 tensor pipe                 (main.py:224)
//   potrf_blocked<64>      safeguarded Cholesky, all LPs per launch                         (main.py:176-182)
//   kb_rhs(0), trsv, kb_dir(0)   predictor: rhs, solves, dx/ds, ratio test, mu_aff, sigma   (main.py:225-228,305-322,588-601)
//   kb_rhs(1), trsv, kb_dir(1)   corrector: same with r4, eta = 0.91, update of x, y, s     (main.py:142-159,604-626,694-696)
// = 6 passes over A per iteration, everything else is O(n) or lives in L2.
//
// For m <= 256 (the benchmark shape) the iteration is restructured to FOUR passes over A, three of them on the
// critical path (ipm_batched_fused.cuh):
//   dmma_ws_kernel         M = A diag(d) A^T                                                 (main.py:224)
//   kb_chol, kb_rhs(0), trsv, kbf_dir<0>   predictor direction + the corrector right-hand side by linearity
//   trsv, kbf_dir<1>            corrector direction, update, residuals of the new point by recurrence
//   kb_residual<.,true>         from-scratch check_optimality, only for LPs the recurrences declare finished
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <chrono>
#include <vector>

#include "chol.cuh"
#include "chol_batched.cuh"
#include "common.cuh"
#include "dmma_gemm.cuh"
#include "dmma_ws.cuh"
#include "ipm_batched_fused.cuh"
#include "kkt_dense.cuh"

using namespace ipm;

namespace {

constexpr int KB_NT = 512;
constexpr int KB_NW = KB_NT / 32;


// ---------------------------------------------------------------------------------------------
// Phase profiler (bench.py roofline): CUDA events on the solve stream around each phase of every lockstep
// iteration; elapsed times are read after the solve has drained, so the timed region is not perturbed.
enum Phase { PH_RESID = 0, PH_SYRK, PH_CHOL, PH_SOLVE, PH_COUNT };
struct Profiler {
    bool enabled = false;
    double ms[PH_COUNT] = {0, 0, 0, 0};
    int64_t calls[PH_COUNT] = {0, 0, 0, 0};
    int64_t lp_iterations = 0;            // sum over lockstep iterations of the number of active LPs
    std::vector<cudaEvent_t> pool;
    std::vector<int> marks;               // phase id that ENDS at event i+1 (event i starts it)
    size_t used = 0;
    cudaEvent_t next() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void begin(cudaStream_t st) { if (enabled) cudaEventRecord(next(), st); }
    void end_phase(int ph, cudaStream_t st) {
        if (!enabled) return;
        cudaEventRecord(next(), st);
        marks.push_back(ph);
    }
    // events were recorded as: begin, end_phase, end_phase, ... per segment; segments are delimited by begin()
    std::vector<size_t> seg_starts;
    std::vector<float> last_ms;           // per recorded interval of the most recent solve
    std::vector<int> last_ph;
    void collect() {
        last_ms.clear(); last_ph.clear();
        if (!enabled) { used = 0; marks.clear(); seg_starts.clear(); return; }
        size_t mi = 0;
        for (size_t sgi = 0; sgi < seg_starts.size(); ++sgi) {
            const size_t e0 = seg_starts[sgi];
            const size_t e1 = (sgi + 1 < seg_starts.size()) ? seg_starts[sgi + 1] : used;
            for (size_t e = e0; e + 1 < e1; ++e, ++mi) {
                float t = 0.f;
                if (cudaEventElapsedTime(&t, pool[e], pool[e + 1]) == cudaSuccess) {
                    ms[marks[mi]] += t;
                    calls[marks[mi]] += 1;
                    last_ms.push_back(t); last_ph.push_back(marks[mi]);
                }
            }
        }
        used = 0; marks.clear(); seg_starts.clear();
    }
    void segment(cudaStream_t st) {
        if (!enabled) return;
        seg_starts.push_back(used);
        cudaEventRecord(next(), st);
    }
};
Profiler g_prof;

// Process-wide options of the batched solver (ipm_batched_set_variant / ipm_batched_set_option): read once at the
// start of every solve, atomics so that setting them from another thread is not a data race.
struct BatchedOptions {
    std::atomic<int> fused{1};          // 0 forces the literal six-pass iteration (tests, A/B timing)
    std::atomic<int> fresh_every{12};   // four-pass path: residuals from scratch every 12th iteration (ipm_b200.h)
    std::atomic<int> refine{1};         // conditional refinement of the corrector (kbf_dir / kb_dir)
    std::atomic<int> syrk_rhs{1};       // four-pass path: predictor right-hand side formed by the SYRK's diagonal tiles
    std::atomic<int> strip_tma{1};      // four-pass path: strips of A through a tensor map (1) or a strip-major copy (0)
    std::atomic<int> handoff{1};        // LPs the refined corrector cannot fix go to the augmented-system kernel
};
std::atomic<int> g_last_handoffs{0};    // LPs handed off in the most recent batched solve (ipm_batched_last_handoffs)
BatchedOptions g_opt;

// IPM_DEBUG_SYNC=1: synchronise and check after every launch of the batched loop and name the launch that failed
// (launch errors are otherwise only collected once per lockstep iteration).
int debug_check(const char* what, cudaStream_t st) {
    static const bool on = getenv("IPM_DEBUG_SYNC") != nullptr;
    if (!on) return IPM_OK;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
        return IPM_ERR_CUDA;
    }
    return IPM_OK;
}

// ---------------------------------------------------------------------------------------------
// One pass over A_i: Ax (warp per row) and A^T y (column partial sums per warp, combined in warp order).
// CAND (3-pass path): only LPs flagged 2 are evaluated (the others just report themselves as active), and the
// predictor operands rcx = (x s)/x, w = d (rc - rcx) of main.py:72,225 are produced here as well.
template <int NPL, bool CAND>
__global__ void __launch_bounds__(KB_NT, (NPL <= 8) ? 2 : 1) kb_residual(const BatchArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* colred = smem;                 // [KB_NW][n]
    __shared__ double sh[32];
    __shared__ double s_nrb2;
    const int lp = blockIdx.x;
    const int flag = a.active[lp];
    if (flag == 0) return;
    if (CAND && flag == 1) {
        if (threadIdx.x == 0) a.act_list[atomicAdd(a.n_active, 1u)] = lp;
        return;
    }
    const int m = a.m, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* A = a.A + (size_t)lp * m * n;
    const double* x = a.x + (size_t)lp * n;
    const double* s = a.s + (size_t)lp * n;
    const double* c = a.c + (size_t)lp * n;
    const double* y = a.y + (size_t)lp * m;
    const double* b = a.b + (size_t)lp * m;
    double* rb = a.rb + (size_t)lp * m;
    double* rc = a.rc + (size_t)lp * n;
    double* d = a.d + (size_t)lp * n;
    double* scal = a.scal + (size_t)lp * S_COUNT;
    const int n2 = n >> 1;

    // x lives in shared memory (after colred) instead of registers: 32 registers less per thread lets two
    // CTAs share an SM, which is what this HBM-bound pass needs to keep enough loads in flight
    double* xsh = colred + (size_t)KB_NW * n;
    for (int k = tid; k < n; k += KB_NT) xsh[k] = x[k];
    __syncthreads();
    const double2* xs2 = reinterpret_cast<const double2*>(xsh);
    double2 ca[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) ca[j] = make_double2(0.0, 0.0);
    double nrb2 = 0.0;
    for (int r = warp; r < m; r += KB_NW) {
        const double2* row = reinterpret_cast<const double2*>(A + (size_t)r * n);
        const double yr = y[r];
        double dot0 = 0.0, dot1 = 0.0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int c2 = j * 32 + lane;
            if (c2 < n2) {
                const double2 v = row[c2];
                const double2 xv = xs2[c2];
                dot0 += v.x * xv.x;
                dot1 += v.y * xv.y;
                ca[j].x += v.x * yr;
                ca[j].y += v.y * yr;
            }
        }
        const double dot = warp_sum(dot0 + dot1);
        if (lane == 0) {
            const double r_b = dot - b[r];
            rb[r] = r_b;
            nrb2 += r_b * r_b;
        }
    }
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const int c2 = j * 32 + lane;
        if (c2 < n2) reinterpret_cast<double2*>(colred + (size_t)warp * n)[c2] = ca[j];
    }
    __syncthreads();
    double nrc2 = 0.0, xs = 0.0, obj = 0.0;
    for (int k = tid; k < n; k += KB_NT) {
        double aty = 0.0;
#pragma unroll
        for (int w = 0; w < KB_NW; ++w) aty += colred[(size_t)w * n + k];
        const double xi = x[k], si = s[k], ci = c[k];
        const double r = aty + si - ci;
        rc[k] = r;
        const double dk = xi / si;
        d[k] = dk;
        if (CAND) {
            const double q = (xi * si) / xi;
            a.rcx[(size_t)lp * n + k] = q;
            a.w[(size_t)lp * n + k] = dk * (r - q);
        }
        nrc2 += r * r;
        xs += xi * si;
        obj += xi * ci;
    }
    nrb2 = block_red<RED_SUM>(nrb2, sh);
    if (tid == 0) s_nrb2 = nrb2;
    nrc2 = block_red<RED_SUM>(nrc2, sh);
    xs = block_red<RED_SUM>(xs, sh);
    obj = block_red<RED_SUM>(obj, sh);
    if (tid == 0) {
        const double nrb = sqrt(s_nrb2), nrc = sqrt(nrc2);
        scal[S_NRB2] = s_nrb2; scal[S_NRB] = nrb; scal[S_NRC2] = nrc2; scal[S_NRC] = nrc;
        scal[S_XS] = xs; scal[S_OBJ] = obj;
        const bool cont = (a.tol * (1.0 + scal[S_NB]) < nrb) || (a.tol * (1.0 + scal[S_NC]) < nrc) || (a.tol < xs);
        scal[S_CONT] = cont ? 1.0 : 0.0;
        const bool go = cont && a.iters[lp] < a.max_iter;
        if (go) a.act_list[atomicAdd(a.n_active, 1u)] = lp;
        a.active[lp] = go ? 1 : 0;
    }
}

// |b|, |c| per LP (once per solve) and state initialisation x = s = 1, y = 0 (main.py:287-302)
__global__ void __launch_bounds__(256) kb_init(const BatchArgs a, int flag, int lp0) {
    __shared__ double sh[32];
    const int lp = lp0 + blockIdx.x, tid = threadIdx.x;
    const int m = a.m, n = a.n;
    double nb = 0.0, nc = 0.0;
    for (int i = tid; i < m; i += blockDim.x) {
        const double v = a.b[(size_t)lp * m + i];
        nb += v * v;
        a.y[(size_t)lp * m + i] = 0.0;
    }
    for (int i = tid; i < n; i += blockDim.x) {
        const double v = a.c[(size_t)lp * n + i];
        nc += v * v;
        a.x[(size_t)lp * n + i] = 1.0;
        a.s[(size_t)lp * n + i] = 1.0;
    }
    nb = block_red<RED_SUM>(nb, sh);
    nc = block_red<RED_SUM>(nc, sh);
    if (tid == 0) {
        double* scal = a.scal + (size_t)lp * S_COUNT;
        for (int i = 0; i < S_COUNT; ++i) scal[i] = 0.0;
        scal[S_NB] = sqrt(nb);
        scal[S_NC] = sqrt(nc);
        a.active[lp] = flag;
        a.iters[lp] = 0;
    }
}

// rcx = rcomp/x, w = d (rc - rcx), rhs = -rb - A w      (main.py:72, 150-152, 225)
template <int NPL>
__global__ void __launch_bounds__(KB_NT) kb_rhs(const BatchArgs a, int kind) {
    extern __shared__ __align__(16) double smem[];
    double* ws = smem;     // [n]
    const int lp = blockIdx.x;
    if (a.active[lp] == 0) return;
    const int m = a.m, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* A = a.A + (size_t)lp * m * n;
    const size_t on = (size_t)lp * n, om = (size_t)lp * m;
    const double sigma_mu = kind ? a.scal[(size_t)lp * S_COUNT + S_SIGMA_MU] : 0.0;
    for (int k = tid; k < n; k += KB_NT) {
        const double xi = a.x[on + k];
        double rcomp = xi * a.s[on + k];
        if (kind) rcomp = rcomp + a.dxa[on + k] * a.dsa[on + k] - sigma_mu;
        const double q = rcomp / xi;
        const double wv = a.d[on + k] * (a.rc[on + k] - q);
        a.rcx[on + k] = q;
        a.w[on + k] = wv;
        ws[k] = wv;
    }
    __syncthreads();
    const int n2 = n >> 1;
    double2 wr[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const int c2 = j * 32 + lane;
        wr[j] = (c2 < n2) ? reinterpret_cast<const double2*>(ws)[c2] : make_double2(0.0, 0.0);
    }
    for (int r = warp; r < m; r += KB_NW) {
        const double2* row = reinterpret_cast<const double2*>(A + (size_t)r * n);
        double dot0 = 0.0, dot1 = 0.0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int c2 = j * 32 + lane;
            if (c2 < n2) {
                const double2 v = row[c2];
                dot0 += v.x * wr[j].x;
                dot1 += v.y * wr[j].y;
            }
        }
        const double dot = warp_sum(dot0 + dot1);
        if (lane == 0) a.rhs[om + r] = -a.rb[om + r] - dot;
    }
}

// The vectors of the predictor right-hand side without the product (four-pass iteration with IPM_BOPT_SYRK_RHS):
// rcx = rcomp/x, v = rc - rcx, w = d v (main.py:72, 150-152); A w is formed by the SYRK's diagonal tiles (DmmaArgs::vvec).
__global__ void __launch_bounds__(256) kb_wvec(const BatchArgs a, double* __restrict__ v) {
    const int lp = blockIdx.x;
    if (a.active[lp] == 0) return;
    const size_t on = (size_t)lp * a.n;
    for (int k = threadIdx.x; k < a.n; k += 256) {
        const double xi = a.x[on + k];
        const double q = (xi * a.s[on + k]) / xi;
        const double vv = a.rc[on + k] - q;
        a.rcx[on + k] = q;
        a.w[on + k] = a.d[on + k] * vv;
        v[on + k] = vv;
    }
}

// u = A^T dy; dx = d u + w; ds = -s dx/x - rcx; ratio test; then
//   kind 0: mu_aff, mu, sigma (main.py:582-600), predictor direction stored for the corrector rhs
//   kind 1: alpha = min(1, eta*min) (main.py:616-623), x += ap dx, y += ad dy, s += ad ds (main.py:694-696)
// dy is read from a.rhs (the batched triangular solve works in place).
// Conditional refinement of the corrector (kind 1, a.refine; the rule and its reason are stated at kbf_dir,
// ipm_batched_fused.cuh): a second sweep over A_i (from L2) forms delta = -rb - A dx; when |delta| > |rb| the LP
// is not updated, dx goes to a.dxc, delta to a.dy and the LP is flagged FLAG_REFINE; the host's next two launches
// solve ddy = M^-1 delta in place (a.dy) and run this kernel again with pass = 1 for the flagged LPs:
// dx = dx_old + d (A^T ddy), ds from dx, step in y along dy + ddy; if |delta| > |rb| even then, the LP is parked for
// the augmented-system kernel (a.handoff_list).
template <int NPL>
__global__ void __launch_bounds__(KB_NT) kb_dir(const BatchArgs a, int kind, int pass) {
    extern __shared__ __align__(16) double smem[];
    double* colred = smem;     // [KB_NW][n]
    __shared__ double sh[32];
    __shared__ double s_alpha[4];
    const int lp = blockIdx.x;
    {
        const int flag = a.active[lp];
        if (pass == 0 ? (flag == 0) : (flag != FLAG_REFINE)) return;
    }
    const int m = a.m, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* A = a.A + (size_t)lp * m * n;
    const size_t on = (size_t)lp * n, om = (size_t)lp * m;
    const double* dy = (pass == 0) ? a.rhs + om : a.dy + om;       // pass 1 streams ddy
    double* scal = a.scal + (size_t)lp * S_COUNT;
    const int n2 = n >> 1;

    double2 ca[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) ca[j] = make_double2(0.0, 0.0);
    for (int r = warp; r < m; r += KB_NW) {
        const double2* row = reinterpret_cast<const double2*>(A + (size_t)r * n);
        const double yr = dy[r];
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int c2 = j * 32 + lane;
            if (c2 < n2) {
                const double2 v = row[c2];
                ca[j].x += v.x * yr;
                ca[j].y += v.y * yr;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const int c2 = j * 32 + lane;
        if (c2 < n2) reinterpret_cast<double2*>(colred + (size_t)warp * n)[c2] = ca[j];
    }
    __syncthreads();
    // each thread owns columns tid, tid + KB_NT (n <= 2*KB_NT)
    double dxv[2], dsv[2], xv[2], sv[2];
    double minp = 1.0, mind = 1.0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = tid + q * KB_NT;
        dxv[q] = dsv[q] = 0.0; xv[q] = sv[q] = 1.0;
        if (k < n) {
            double u = 0.0;
#pragma unroll
            for (int w = 0; w < KB_NW; ++w) u += colred[(size_t)w * n + k];
            const double xi = a.x[on + k], si = a.s[on + k];
            const double dxi = a.d[on + k] * u + ((pass == 0) ? a.w[on + k] : a.dxc[on + k]);
            const double dsi = (-si * dxi / xi) - a.rcx[on + k];
            dxv[q] = dxi; dsv[q] = dsi; xv[q] = xi; sv[q] = si;
            if (dxi < 0.0) minp = fmin(minp, -xi / dxi);
            if (dsi < 0.0) mind = fmin(mind, -si / dsi);
        }
    }
    minp = block_red<RED_MIN>(minp, sh);
    if (tid == 0) s_alpha[0] = minp;
    mind = block_red<RED_MIN>(mind, sh);
    if (tid == 0) s_alpha[1] = mind;
    __syncthreads();
    double ap = s_alpha[0], ad = s_alpha[1];
    if (kind == 0) {
        double part = 0.0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = tid + q * KB_NT;
            if (k < n) {
                a.dxa[on + k] = dxv[q];
                a.dsa[on + k] = dsv[q];
                part += (xv[q] + ap * dxv[q]) * (sv[q] + ad * dsv[q]);
            }
        }
        part = block_red<RED_SUM>(part, sh);
        if (tid == 0) {
            const double mu_aff = part / (double)n, mu = scal[S_XS] / (double)n;
            const double r = mu_aff / mu, sigma = r * r * r;
            scal[S_AP_AFF] = ap; scal[S_AD_AFF] = ad; scal[S_MU_AFF] = mu_aff; scal[S_MU] = mu;
            scal[S_SIGMA] = sigma; scal[S_SIGMA_MU] = sigma * mu;
        }
    } else {
        if (a.refine) {
            // every thread is past the barriers above, so the column partials are dead: dx takes their place
            double* dxs = colred;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = tid + q * KB_NT;
                if (k < n) dxs[k] = dxv[q];
            }
            __syncthreads();
            double2 wr[NPL];
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int c2 = j * 32 + lane;
                wr[j] = (c2 < n2) ? reinterpret_cast<const double2*>(dxs)[c2] : make_double2(0.0, 0.0);
            }
            double* dsm = dxs + n;                     // [m] delta (second row of the dead partials)
            double nd2 = 0.0, nr2 = 0.0;
            for (int r = warp; r < m; r += KB_NW) {
                const double2* row = reinterpret_cast<const double2*>(A + (size_t)r * n);
                double dot0 = 0.0, dot1 = 0.0;
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const int c2 = j * 32 + lane;
                    if (c2 < n2) {
                        const double2 v = row[c2];
                        dot0 += v.x * wr[j].x;
                        dot1 += v.y * wr[j].y;
                    }
                }
                const double dot = warp_sum(dot0 + dot1);
                if (lane == 0) {
                    const double rbr = a.rb[om + r], dl = -rbr - dot;
                    dsm[r] = dl;
                    nd2 += dl * dl;
                    nr2 += rbr * rbr;
                }
            }
            nd2 = block_red<RED_SUM>(nd2, sh);
            if (tid == 0) s_alpha[2] = nd2;
            nr2 = block_red<RED_SUM>(nr2, sh);
            if (tid == 0) {
                const double fl = 1e-3 * a.tol * (1.0 + scal[S_NB]);            // see kbf_dir
                // pass 0: refine when |delta| > 0.1 |rb|; pass 1: hand off when the refined step still has |delta| > |rb|
                const double lim2 = (pass == 0) ? KF_REFINE_THRESH * KF_REFINE_THRESH * nr2 : nr2;
                s_alpha[3] = (s_alpha[2] > lim2 && s_alpha[2] > fl * fl) ? 1.0 : 0.0;   // NaN compares false
                if ((pass == 0 && a.refine == 2) || (pass != 0 && a.handoff == 2)) s_alpha[3] = 1.0;    // test hooks
            }
            __syncthreads();
            if (s_alpha[3] != 0.0) {
                if (pass == 0) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int k = tid + q * KB_NT;
                        if (k < n) a.dxc[on + k] = dxv[q];
                    }
                    for (int i = tid; i < m; i += KB_NT) a.dy[om + i] = dsm[i];
                    if (tid == 0) {
                        a.active[lp] = FLAG_REFINE;
                        scal[S_NREFINE] = scal[S_NREFINE] + 1.0;
                    }
                    return;
                }
                if (a.handoff) {
                    if (tid == 0) {
                        const unsigned slot = atomicAdd(a.n_handoff, 1u);
                        a.handoff_list[slot] = lp;
                        scal[S_HANDOFF] = 1.0;
                        a.active[lp] = 0;
                    }
                    return;
                }
            }
        }
        ap = fmin(1.0, a.eta * ap);
        ad = fmin(1.0, a.eta * ad);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = tid + q * KB_NT;
            if (k < n) {
                a.x[on + k] = xv[q] + ap * dxv[q];
                a.s[on + k] = sv[q] + ad * dsv[q];
            }
        }
        for (int i = tid; i < m; i += KB_NT) {
            const double dyi = (pass == 0) ? dy[i] : a.rhs[om + i] + dy[i];      // pass 1: dy + ddy
            a.y[om + i] = a.y[om + i] + ad * dyi;
        }
        if (tid == 0) {
            scal[S_AP] = ap; scal[S_AD] = ad;
            a.iters[lp] += 1;
            a.active[lp] = 1;
        }
    }
}

__global__ void kb_finalize(const BatchArgs a, int B, double* obj, int* iters, int* status) {
    const int lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= B) return;
    const double* sc = a.scal + (size_t)lp * S_COUNT;
    const bool finite = isfinite(sc[S_NRB]) && isfinite(sc[S_NRC]) && isfinite(sc[S_XS]) && isfinite(sc[S_OBJ]);
    int st = IPM_STATUS_CONVERGED;
    if (!finite) st = IPM_STATUS_NAN;
    else if (sc[S_CONT] > 0.5) st = IPM_STATUS_MAX_ITER;
    if (obj) obj[lp] = sc[S_OBJ];
    if (iters) iters[lp] = a.iters[lp];
    if (status) status[lp] = st;
}

// ---------------------------------------------------------------------------------------------
struct Workspace {
    BatchArgs a;
    double* M;
    int64_t ldm;
    double* ka_work;     // ka_slots * ka_work_doubles
    unsigned* h_nact;    // pinned, 8 words: [0,1] active LPs per check slot, [2,3] hand-off count per slot, [4] final
};

int64_t at_doubles(int B, int m, int n) {        // strip-major copy of A (3-pass path with IPM_BOPT_STRIP_TMA = 0 only)
    if (m > KF_MAX_M || g_opt.strip_tma.load() != 0) return 0;
    return (int64_t)B * ceil_div(n, KF_W) * (32 * kf_nrp(m)) * KF_W;
}

// slots of the augmented-system kernel (one matrix of order n + m each): LPs handed off beyond that run in rounds
int ka_slots(int B, int m, int n) { return (m + n > KA_MAX_N) ? 0 : std::min(B, 48); }

int64_t ws_bytes(int B, int m, int n) {
    const int64_t ldm = round_up(m, 16);
    int64_t doubles = (int64_t)B * (10 * (int64_t)n + 4 * (int64_t)m + S_COUNT) + (int64_t)B * m * ldm + at_doubles(B, m, n) +
                      (int64_t)ka_slots(B, m, n) * ka_work_doubles(m, n);
    int64_t bytes = doubles * 8 + (int64_t)B * 4 * sizeof(int) + 256 + 1024;
    return round_up(bytes, 256);
}

void carve(Workspace& w, void* base, int B, int m, int n) {
    double* p = reinterpret_cast<double*>(base);
    auto take = [&](int64_t len) { double* q = p; p += len; return q; };
    const int64_t bn = (int64_t)B * n, bm = (int64_t)B * m;
    w.ldm = round_up(m, 16);
    w.M = take((int64_t)B * m * w.ldm);
    w.a.At = take(at_doubles(B, m, n));
    w.a.x = take(bn); w.a.s = take(bn); w.a.rc = take(bn); w.a.d = take(bn); w.a.w = take(bn); w.a.rcx = take(bn);
    w.a.dxa = take(bn); w.a.dsa = take(bn); w.a.dxc = take(bn); w.a.dsc = take(bn);
    w.a.y = take(bm); w.a.rb = take(bm); w.a.dy = take(bm); w.a.rhs = take(bm);
    w.a.scal = take((int64_t)B * S_COUNT);
    w.ka_work = take((int64_t)ka_slots(B, m, n) * ka_work_doubles(m, n));
    int* ip = reinterpret_cast<int*>(p);
    w.a.active = ip; ip += B;
    w.a.iters = ip; ip += B;
    w.a.handoff_list = ip; ip += B;
    w.a.act_list = ip; ip += B;
    uintptr_t u = (reinterpret_cast<uintptr_t>(ip) + 63) & ~(uintptr_t)63;
    w.a.n_active = reinterpret_cast<unsigned*>(u);          // [0], [16]: the two check slots; [32]: hand-off count
    w.a.n_handoff = w.a.n_active + 32;
    w.a.kf_ctr = w.a.n_active + 48;                         // [48], [49]: work counter of the direction kernels
    w.a.m = m; w.a.n = n;
}

// Host-buffer entry point: the batch arrives in chunks on a copy stream while the lockstep loop is already
// running; a chunk's LPs join the loop (kb_init) at the first iteration after its copy has landed.  Every LP keeps
// its own iteration counter, so joining late changes nothing for it.
struct Arrival {
    int nchunks = 0;
    const int* first = nullptr;
    const int* count = nullptr;
    cudaEvent_t* landed = nullptr;      // recorded on the copy stream behind chunk k
    int next = 0;                       // first chunk that has not joined yet
};

template <int KIND, int NRP, int SRC>
int launch_kbf_one(int pass, const BatchArgs& a, const CUtensorMap& tm, int B, int m, cudaStream_t st) {
    const int nbuf = kf_nbuf(m, a.n);
    const size_t sm = kf_smem_bytes(m, a.n, nbuf);
    IPM_TRY(ensure_dyn_smem(kbf_dir<KIND, NRP, SRC>, sm));
    const int grid = std::min(B, kNumSMs);        // persistent: one CTA per SM, LPs from the work counter a.kf_ctr
    kbf_dir<KIND, NRP, SRC><<<grid, KF_NTT, sm, st>>>(a, pass, B, nbuf, tm);
    return IPM_OK;
}
template <int SRC>
int launch_kbf_dir_src(int kind, int pass, const BatchArgs& a, const CUtensorMap& tm, int B, int m, cudaStream_t st) {
    switch (kf_nrp(m) * 2 + kind) {
        case 2: return launch_kbf_one<0, 1, SRC>(pass, a, tm, B, m, st);
        case 3: return launch_kbf_one<1, 1, SRC>(pass, a, tm, B, m, st);
        case 4: return launch_kbf_one<0, 2, SRC>(pass, a, tm, B, m, st);
        case 5: return launch_kbf_one<1, 2, SRC>(pass, a, tm, B, m, st);
        case 8: return launch_kbf_one<0, 4, SRC>(pass, a, tm, B, m, st);
        case 9: return launch_kbf_one<1, 4, SRC>(pass, a, tm, B, m, st);
        case 16: return launch_kbf_one<0, 8, SRC>(pass, a, tm, B, m, st);
        default: return launch_kbf_one<1, 8, SRC>(pass, a, tm, B, m, st);
    }
}
// (the launch raises the kernel's dynamic shared-memory limit itself, for the instantiation and size it is about to use:
// one map lookup, ensure_dyn_smem)
int launch_kbf_dir(int kind, int pass, bool tma, const BatchArgs& a, const CUtensorMap& tm, int B, int m, cudaStream_t st) {
    if (tma) return launch_kbf_dir_src<1>(kind, pass, a, tm, B, m, st);
    return launch_kbf_dir_src<0>(kind, pass, a, tm, B, m, st);
}

// Host-side resources of one lockstep loop - the pinned page the "still active" counters are read back into, the events
// of the check slots, the high-priority streams of the augmented-system kernel - come from a per-device pool and go
// back to it: cudaMallocHost / cudaFreeHost and stream creation per solve cost little most of the time and hundreds of
// milliseconds now and then (page locking and stream creation serialise on the driver; with eight processes on one box
// one such stall per few dozen solves made every step wait for the slowest rank: tools/block_times.py).
constexpr int KA_STREAMS = 6;
struct LoopRes {
    int dev = -1;
    unsigned* h_nact = nullptr;                 // [8] pinned
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st_kas[KA_STREAMS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void destroy() {
        if (h_nact) cudaFreeHost(h_nact);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& q : st_kas) if (q) cudaStreamDestroy(q);
        *this = LoopRes();
    }
};
static std::mutex g_loopres_mu;
static std::vector<LoopRes*> g_loopres_free[16];
static int acquire_loopres(LoopRes** out) {
    int dev = 0;
    IPM_CUDA_OK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_loopres_mu);
        auto& fr = g_loopres_free[dev & 15];
        for (size_t i = 0; i < fr.size(); ++i)
            if (fr[i]->dev == dev) { *out = fr[i]; fr.erase(fr.begin() + (long)i); return IPM_OK; }
    }
    LoopRes* r = new LoopRes();
    r->dev = dev;
    auto fail = [&](cudaError_t e) { g_last_error = cudaGetErrorString(e); r->destroy(); delete r; return IPM_ERR_CUDA; };
    cudaError_t e = cudaMallocHost(&r->h_nact, 8 * sizeof(unsigned));
    if (e != cudaSuccess) return fail(e);
    for (int i = 0; i < 4; ++i)
        if ((e = cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    int lo = 0, hi = 0;
    if ((e = cudaDeviceGetStreamPriorityRange(&lo, &hi)) != cudaSuccess) return fail(e);
    for (int i = 0; i < KA_STREAMS; ++i)
        if ((e = cudaStreamCreateWithPriority(&r->st_kas[i], cudaStreamNonBlocking, hi)) != cudaSuccess) return fail(e);
    *out = r;
    return IPM_OK;
}
static void release_loopres(LoopRes* r, bool healthy) {
    if (!r) return;
    if (!healthy) { r->destroy(); delete r; return; }          // a failed solve may have left work on the streams
    std::lock_guard<std::mutex> lk(g_loopres_mu);
    g_loopres_free[r->dev & 15].push_back(r);
}
struct LoopResLease {
    LoopRes* r = nullptr;
    bool ok = false;
    ~LoopResLease() { release_loopres(r, ok); }
};
static void drop_loopres_pool() {
    std::lock_guard<std::mutex> lk(g_loopres_mu);
    for (auto& fr : g_loopres_free) {
        for (LoopRes* r : fr) { cudaSetDevice(r->dev); r->destroy(); delete r; }
        fr.clear();
    }
}

template <int NPL>
int run_batched(Workspace& w, int B, int m, int n, double tau, cudaStream_t st, int* iterations_run, Arrival* arr) {
    BatchArgs& a = w.a;
    const size_t smem_col = std::max((size_t)KB_NW * n, (size_t)n + m) * sizeof(double);
    const size_t smem_res = smem_col + (size_t)n * sizeof(double);
    const size_t smem_w = (size_t)n * sizeof(double);
    DmmaArgs g;
    g.P = a.A; g.ldp = n; g.strideP = (int64_t)m * n;
    g.Q = a.A; g.ldq = n; g.strideQ = (int64_t)m * n;
    g.dvec = a.d; g.strideD = n;
    g.C = w.M; g.ldc = w.ldm; g.strideC = (int64_t)m * w.ldm;
    g.rowsP = m; g.rowsQ = m; g.K = n; g.lower_only = 1; g.active = a.active;
    // 3-pass iteration (ipm_batched_fused.cuh) when one CTA can hold a column strip of A_i
    const bool fused = g_opt.fused.load() != 0 && m <= KF_MAX_M && (n % 2 == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    const bool tma = fused && g_opt.strip_tma.load() != 0;
    // predictor right-hand side inside the SYRK (one pass over A less): v = rc - rcx lives in dxc until the corrector
    // pass of the same iteration overwrites it
    const bool syrk_rhs = fused && g_opt.syrk_rhs.load() != 0 && ws_eligible(g) &&
                          ((reinterpret_cast<uintptr_t>(a.dxc) & 15) == 0);
    if (syrk_rhs) { g.vvec = a.dxc; g.strideV = n; g.rbvec = a.rb; g.rhs = a.rhs; g.strideR = m; }
    const bool refine = g_opt.refine.load() != 0;
    a.refine = refine ? g_opt.refine.load() : 0;
    a.handoff = (refine && g_opt.handoff.load() != 0 && ka_slots(B, m, n) > 0) ? g_opt.handoff.load() : 0;
    IPM_CUDA_OK(cudaMemsetAsync(a.n_handoff, 0, sizeof(unsigned), st));
    IPM_CUDA_OK(cudaMemsetAsync(a.kf_ctr, 0, 2 * sizeof(unsigned), st));
    {
        IPM_TRY(ensure_dyn_smem(kb_residual<NPL, false>, 131072 + 8192));
        IPM_TRY(ensure_dyn_smem(kb_residual<NPL, true>, 131072 + 8192));
        IPM_TRY(ensure_dyn_smem(kb_dir<NPL>, 131072 + 16384));
        IPM_TRY(ensure_dyn_smem(k_trsv_batched, 65536));
        IPM_TRY(ensure_dyn_smem(k_trsv_batched_inv, trsv_batched_inv_smem(32 * TRSVI_MAX_BLK)));
    }
    CUtensorMap tmapA;
    memset(&tmapA, 0, sizeof(tmapA));
    if (tma) IPM_TRY(kf_make_strip_tmap(&tmapA, a.A, B, m, n, 32 * kf_nrp(m)));
    IPM_TRY(debug_check("before the first launch (stale error)", st));
    // Every per-LP kernel of the loop takes LP = block index and skips LPs that are not active.  With host buffers the LPs
    // join in index order while their copies land, so the grids cover the joined prefix only: during the ramp an
    // iteration on the first few hundred LPs does not pay for thousands of CTAs that start only to find a zero flag
    // (eleven launches per iteration, 28-55 waves each).
    int Bg = 0;
    auto join = [&](int lp0, int cnt) {             // LPs lp0 .. lp0+cnt-1 enter the loop
        Bg = std::max(Bg, lp0 + cnt);
        kb_init<<<cnt, 256, 0, st>>>(a, fused ? 2 : 1, lp0);
        count_launch();
        if (fused && !tma) {
            const int nstrips = ceil_div(n, KF_W);
            kbf_repack<<<dim3(nstrips, cnt), 256, 0, st>>>(a, 32 * kf_nrp(m), nstrips, lp0);
            count_launch();
        }
    };
    auto join_landed = [&](bool block_for_one) -> int {
        // chunks whose copy has completed (cudaEventQuery: the loop never waits for data it can do without)
        while (arr->next < arr->nchunks) {
            cudaError_t q = block_for_one ? cudaEventSynchronize(arr->landed[arr->next]) : cudaEventQuery(arr->landed[arr->next]);
            if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; }     // not an error: clear it
            IPM_CUDA_OK(q);
            IPM_CUDA_OK(cudaStreamWaitEvent(st, arr->landed[arr->next], 0));
            join(arr->first[arr->next], arr->count[arr->next]);
            ++arr->next;
            block_for_one = false;
        }
        return IPM_OK;
    };
    if (arr) {
        IPM_CUDA_OK(cudaMemsetAsync(a.active, 0, (size_t)B * sizeof(int), st));
        IPM_TRY(join_landed(true));
    } else {
        join(0, B);
    }
    // The host reads the "LPs still active" counter of check k only after check k+1 has been enqueued, so the
    // GPU never idles on the host round trip; the price is one empty iteration (every kernel skips inactive LPs)
    // after the last LP has converged.  Every LP is bounded by max_iter, so the loop is.
    unsigned* nact_base = a.n_active;
    LoopResLease lease;
    IPM_TRY(acquire_loopres(&lease.r));
    cudaEvent_t* ev = lease.r->ev;                                  // [0,1]: checks; [2,3]: "SYRK of iteration it done"
    if (!w.h_nact) w.h_nact = lease.r->h_nact;
    cudaEvent_t last_syrk = nullptr;            // most recent SYRK enqueued on st
    int it = 0, bodies = 0;
    bool all_joined[2] = {true, true};          // per check slot: had every chunk joined when the check was enqueued?
    bool joined_pending = false;                // a chunk joined after the last check was enqueued
    const bool small_m = m <= 32 * TRSVI_MAX_BLK;
    // LPs parked by the corrector pass (a.handoff_list) are continued by the augmented-system kernel on a second,
    // higher-priority stream WHILE the lockstep loop runs on: a parked LP is off the loop's books (active = 0), its
    // iterate is final, and one CTA per LP for a few milliseconds hides behind the remaining iterations.
    // (several streams: LPs are parked in different iterations, and launches on one stream would run one after the other)
    cudaStream_t* st_kas = lease.r->st_kas;
    int ka_next_stream = 0;
    int ka_pending[KA_STREAMS] = {0, 0, 0, 0, 0, 0};      // CTAs launched on each stream and not known to have finished
    const int slots = ka_slots(B, m, n);
    int ka_launched = 0;
    KktArgs kk;
    kk.A = a.A; kk.b = a.b; kk.c = a.c; kk.x = a.x; kk.s = a.s; kk.y = a.y; kk.scal = a.scal; kk.iters = a.iters;
    kk.m = m; kk.n = n; kk.tol = a.tol; kk.eta = a.eta; kk.max_iter = a.max_iter;
    auto ka_sync_all = [&]() -> int {
        for (int i = 0; i < KA_STREAMS; ++i) IPM_CUDA_OK(cudaStreamSynchronize(st_kas[i]));
        return IPM_OK;
    };
    auto launch_parked = [&](int upto) -> int {          // list entries [ka_launched, upto) whose state is final
        upto = std::min(upto, slots);                    // the workspace has `slots` matrices; the rest waits
        if (upto <= ka_launched) return IPM_OK;
        kk.list = a.handoff_list + ka_launched;
        kk.work = w.ka_work + (size_t)ka_launched * ka_work_doubles(m, n);
        // The host learns of a parked LP at the moment the GPU passes from the residual check to the SYRK that is
        // ALREADY enqueued with the full grid.  If the augmented-system CTAs (high priority) win that race - they do
        // whenever a short kernel precedes the SYRK - one SYRK CTA finds no SM and runs after another has finished:
        // the launch takes twice as long (measured: 18.6 instead of 9.55 ms).  So they start behind that SYRK; the
        // next one is launched with the SMs they hold left out (below).
        if (last_syrk) IPM_CUDA_OK(cudaStreamWaitEvent(st_kas[ka_next_stream], last_syrk, 0));
        IPM_TRY(ka_launch(kk, upto - ka_launched, st_kas[ka_next_stream]));
        ka_pending[ka_next_stream] += upto - ka_launched;
        ka_next_stream = (ka_next_stream + 1) % KA_STREAMS;
        ka_launched = upto;
        return IPM_OK;
    };
    static const size_t trsv_pad_kb = [] {       // A/B: extra dynamic shared memory caps the CTAs per SM (IPM_TRSV_SMEM_KB)
        const char* e = getenv("IPM_TRSV_SMEM_KB");
        const int v = e ? atoi(e) : 0;
        return (size_t)((v >= 0 && v <= 220) ? v : 0);
    }();
    const size_t trsv_smem = std::max(trsv_batched_inv_smem(m), trsv_pad_kb * 1024);
    if (trsv_pad_kb) IPM_TRY(ensure_dyn_smem(k_trsv_batched_inv, std::max(trsv_batched_inv_smem(32 * TRSVI_MAX_BLK), trsv_pad_kb * 1024)));
    auto launch_trsv = [&](const TrsvBatchedArgs& t) {
        if (small_m) k_trsv_batched_inv<<<Bg, TRSVB_NT, trsv_smem, st>>>(t);
        else k_trsv_batched<<<Bg, TRSVB_NT, trsv_batched_smem(m), st>>>(t);
        count_launch();
    };
    static const bool trace_host = getenv("IPM_TRACE_HOST") != nullptr;      // where the host side of one solve spends its time
    const auto th0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    double th_wait_max = 0.0, th_parked = 0.0;
    int th_wait_it = -1;
    for (;; ++it) {
        const int slot = it & 1;
        if (arr && it > 0) {
            const int before = arr->next;
            IPM_TRY(join_landed(false));
            if (arr->next != before) joined_pending = true;
        }
        const bool counted_join = joined_pending;       // those LPs pass through this iteration's check first
        joined_pending = false;
        all_joined[slot] = !arr || arr->next == arr->nchunks;
        a.n_active = nact_base + slot * 16;
        IPM_CUDA_OK(cudaMemsetAsync(a.n_active, 0, sizeof(unsigned), st));
        g_prof.segment(st);
        if (fused) kb_residual<NPL, true><<<Bg, KB_NT, smem_res, st>>>(a);
        else kb_residual<NPL, false><<<Bg, KB_NT, smem_res, st>>>(a);
        count_launch();
        IPM_TRY(debug_check("kb_init / kb_residual", st));
        g_prof.end_phase(PH_RESID, st);
        IPM_CUDA_OK(cudaMemcpyAsync(w.h_nact + slot, a.n_active, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        if (a.handoff)
            IPM_CUDA_OK(cudaMemcpyAsync(w.h_nact + 2 + slot, a.n_handoff, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        IPM_CUDA_OK(cudaEventRecord(ev[slot], st));
        if (it > 0) {
            const auto tw = std::chrono::steady_clock::now();
            IPM_CUDA_OK(cudaEventSynchronize(ev[slot ^ 1]));
            if (trace_host) { const double d = since(tw); if (d > th_wait_max) { th_wait_max = d; th_wait_it = it; } }
            const unsigned cnt = w.h_nact[slot ^ 1];
            // everything the stream did before that check has completed: LPs parked by then can start now
            const auto tp = std::chrono::steady_clock::now();
            if (a.handoff) IPM_TRY(launch_parked((int)w.h_nact[2 + (slot ^ 1)]));
            if (trace_host) th_parked += since(tp);
            if (cnt == 0) {
                if (all_joined[slot ^ 1]) break;
                if (!counted_join) {
                    // nothing is active until the next chunk lands (only when the copies are slower than the
                    // solves): wait for it, let it join, and send it through the next check before any body runs
                    if (arr->next < arr->nchunks) IPM_TRY(join_landed(true));
                    joined_pending = true;
                    continue;
                }
            }
            ++bodies;
            if (g_prof.enabled) g_prof.lp_iterations += cnt;
        }
        g_prof.segment(st);
        if (a.handoff) {
            // The SYRK is persistent with one CTA per SM and a static tile schedule: an SM held by an augmented-system
            // CTA would make one SYRK CTA wait for another to finish (twice the launch time).  Leave those SMs out.
            int busy = 0;
            for (int i = 0; i < KA_STREAMS; ++i) {
                if (ka_pending[i] == 0) continue;
                if (cudaStreamQuery(st_kas[i]) == cudaSuccess) ka_pending[i] = 0;
                else (void)cudaGetLastError();            // cudaErrorNotReady is not an error
                busy += ka_pending[i] * ka_cluster_ctas().load();
            }
            g.max_ctas = (busy > 0 && busy < kNumSMs / 2) ? kNumSMs - busy : 0;
        }
        if (syrk_rhs) {
            kb_wvec<<<Bg, 256, 0, st>>>(a, a.dxc);
            count_launch();
        }
        IPM_TRY((dmma_syrk_auto<0>(g, Bg, st)));
        if (a.handoff) {
            last_syrk = ev[2 + slot];
            IPM_CUDA_OK(cudaEventRecord(last_syrk, st));
        }
        IPM_TRY(debug_check("syrk", st));
        g_prof.end_phase(PH_SYRK, st);
        if (m <= KBC_MAX_M_BIG)
            IPM_TRY(potrf_batched_fused(w.M, w.ldm, (int64_t)m * w.ldm, m, Bg, a.scal, S_COUNT, tau, a.active, st));
        else
            IPM_TRY((potrf_blocked<64, 256, 128>(w.M, w.ldm, (int64_t)m * w.ldm, m, Bg, a.scal, S_COUNT, tau,
                                                 a.active, st)));
        IPM_TRY(debug_check("cholesky", st));
        g_prof.end_phase(PH_CHOL, st);
        TrsvBatchedArgs t;
        t.L = w.M; t.ldm = w.ldm; t.strideM = (int64_t)m * w.ldm; t.v = a.rhs; t.strideV = m; t.m = m;
        t.active = a.active;
        if (fused) t.out = a.dy;           // the right-hand side survives: the corrector's is built on top of it
        for (int kind = 0; kind < 2; ++kind) {
            if (!fused || (kind == 0 && !syrk_rhs)) {
                kb_rhs<NPL><<<Bg, KB_NT, smem_w, st>>>(a, kind);
                count_launch();
            }
            IPM_TRY(debug_check("kb_rhs", st));
            launch_trsv(t);
            IPM_TRY(debug_check("trsv", st));
            if (!fused) kb_dir<NPL><<<Bg, KB_NT, smem_col, st>>>(a, kind, 0);
            else IPM_TRY(launch_kbf_dir(kind, 0, tma, a, tmapA, Bg, m, st));
            count_launch();
            IPM_TRY(debug_check(kind ? "direction (corrector)" : "direction (predictor)", st));
        }
        if (refine) {
            // LPs whose corrector asked for a refinement (flag FLAG_REFINE; about one in seventy per solve): delta is in
            // a.rhs (fused) / a.dy (six-pass); ddy = M^-1 delta in place on the same factor, then the corrector pass
            // again, which applies the step incrementally
            TrsvBatchedArgs t2 = t;
            t2.only_flag = FLAG_REFINE; t2.out = nullptr;
            t2.v = fused ? a.rhs : a.dy;
            launch_trsv(t2);
            if (!fused) kb_dir<NPL><<<Bg, KB_NT, smem_col, st>>>(a, 1, 1);
            else IPM_TRY(launch_kbf_dir(1, 1, tma, a, tmapA, Bg, m, st));
            count_launch();
            IPM_TRY(debug_check("refinement pass", st));
        }
        g_prof.end_phase(PH_SOLVE, st);
        IPM_TRY(launch_check());
    }
    a.n_active = nact_base;
    // ---- LPs parked late (or beyond the workspace's slots): the same kernel, now with the machine to itself
    int handed = 0;
    const double th_loop = since(th0);
    double th_drain = 0.0, th_ka = 0.0;
    if (a.handoff) {
        IPM_CUDA_OK(cudaMemcpyAsync(w.h_nact + 4, a.n_handoff, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        IPM_CUDA_OK(cudaStreamSynchronize(st));
        th_drain = since(th0) - th_loop;
        handed = (int)w.h_nact[4];
        IPM_TRY(launch_parked(handed));
        IPM_TRY(ka_sync_all());
        IPM_TRY(debug_check("augmented-system kernel", st_kas[0]));
        for (int first = ka_launched; first < handed; first += slots) {       // more LPs than slots: in rounds
            kk.list = a.handoff_list + first;
            kk.work = w.ka_work;
            IPM_TRY(ka_launch(kk, std::min(slots, handed - first), st_kas[0]));
            IPM_CUDA_OK(cudaStreamSynchronize(st_kas[0]));
        }
    }
    th_ka = since(th0) - th_loop - th_drain;
    g_last_handoffs.store(handed);
    IPM_CUDA_OK(cudaStreamSynchronize(st));
    lease.ok = true;                            // nothing of this solve is left on the pooled streams and events
    if (trace_host)
        fprintf(stderr, "[ipm host trace] loop enqueued %.1f ms (longest wait for a check %.1f ms at iteration %d, launch_parked %.1f ms), "
                        "stream drained +%.1f ms, augmented-system tail +%.1f ms, %d handed off, %d of them started by the loop or its tail\n",
                th_loop, th_wait_max, th_wait_it, th_parked, th_drain, th_ka, handed, ka_launched);
    if (iterations_run) *iterations_run = bodies;
    return IPM_OK;
}

int solve_on_device(int B, int m, int n, const double* A_d, const double* b_d, const double* c_d, double tol,
                    int max_iter, double* obj_d, int* iters_d, int* status_d, double* x_d, void* work_d,
                    unsigned* h_nact, cudaStream_t st, int* iterations_run, Arrival* arr = nullptr) {
    Workspace w;
    carve(w, work_d, B, m, n);
    w.h_nact = h_nact;
    w.a.A = A_d; w.a.b = b_d; w.a.c = c_d;
    w.a.tol = tol; w.a.eta = 0.91; w.a.max_iter = max_iter;
    w.a.fresh_every = g_opt.fresh_every.load();
    const double tau = 1e-30;
    if (n <= 512) IPM_TRY(run_batched<8>(w, B, m, n, tau, st, iterations_run, arr));
    else IPM_TRY(run_batched<16>(w, B, m, n, tau, st, iterations_run, arr));
    kb_finalize<<<ceil_div(B, 256), 256, 0, st>>>(w.a, B, obj_d, iters_d, status_d);
    count_launch();
    if (x_d) IPM_CUDA_OK(cudaMemcpyAsync(x_d, w.a.x, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return launch_check();
}

int check_shape(int B, int m, int n) {
    if (B <= 0 || m <= 0 || n <= 0) return IPM_ERR_SHAPE;
    if ((n & 1) || n > 2 * KB_NT || m > 2048) {
        g_last_error = "batched path needs even n <= 1024 and m <= 2048";
        return IPM_ERR_SHAPE;
    }
    return IPM_OK;
}

}  // namespace

extern "C" {

int ipm_batched_set_variant(int three_pass, int refresh_every) {
    if (refresh_every < 0) return IPM_ERR_ARG;
    g_opt.fused.store(three_pass != 0);
    g_opt.fresh_every.store(refresh_every);
    return IPM_OK;
}

int ipm_batched_set_option(int option, int value) {
    switch (option) {
        case IPM_BOPT_REFINE: g_opt.refine.store(value < 0 ? 0 : (value > 2 ? 2 : value)); return IPM_OK;
        case IPM_BOPT_STRIP_TMA: g_opt.strip_tma.store(value != 0); return IPM_OK;
        case IPM_BOPT_SYRK_RHS: g_opt.syrk_rhs.store(value != 0); return IPM_OK;
        case IPM_BOPT_HANDOFF: g_opt.handoff.store(value < 0 ? 0 : (value > 2 ? 2 : value)); return IPM_OK;
        default: return IPM_ERR_ARG;
    }
}

int ipm_batched_last_handoffs(void) { return g_last_handoffs.load(); }

static long long g_kkt_prof[16];
int ipm_kkt_last_profile(int64_t out[16]) {
    if (!out) return IPM_ERR_ARG;
    for (int i = 0; i < 16; ++i) out[i] = g_kkt_prof[i];
    return IPM_OK;
}

int ipm_set_kkt_cluster(int ctas) {
    if (ctas != 1 && ctas != 2 && ctas != 4 && ctas != 8) return IPM_ERR_ARG;
    ka_cluster_ctas().store(ctas);
    return IPM_OK;
}

int ipm_solve_dense_kkt(int device_ordinal, int m, int n, const double* A, const double* b, const double* c, double tol,
                        int max_iter, double* x, double* y, double* s, double* obj, int* iters, int* status) {
    if (!A || !b || !c || max_iter < 0) return IPM_ERR_ARG;
    if (m <= 0 || n <= 0) return IPM_ERR_SHAPE;
    if (m + n > KA_MAX_N) {
        g_last_error = "ipm_solve_dense_kkt: n + m exceeds " + std::to_string(KA_MAX_N);
        return IPM_ERR_SHAPE;
    }
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    struct Bufs {
        double *A = nullptr, *vec = nullptr, *work = nullptr, *scal = nullptr;
        int* ints = nullptr;
        ~Bufs() { cudaFree(A); cudaFree(vec); cudaFree(work); cudaFree(scal); cudaFree(ints); }
    } d;
    const size_t nv = (size_t)2 * m + 3 * n;                 // b, y | c, x, s
    IPM_CUDA_OK(cudaMalloc(&d.A, (size_t)m * n * sizeof(double)));
    IPM_CUDA_OK(cudaMalloc(&d.vec, nv * sizeof(double)));
    IPM_CUDA_OK(cudaMalloc(&d.work, (size_t)ka_work_doubles(m, n) * sizeof(double)));
    IPM_CUDA_OK(cudaMalloc(&d.scal, S_COUNT * sizeof(double)));
    IPM_CUDA_OK(cudaMalloc(&d.ints, 4 * sizeof(int) + 16 * sizeof(long long)));
    IPM_CUDA_OK(cudaMemset(d.ints, 0, 4 * sizeof(int) + 16 * sizeof(long long)));
    double *db = d.vec, *dy = db + m, *dc = dy + m, *dx = dc + n, *ds = dx + n;
    IPM_CUDA_OK(cudaMemcpy(d.A, A, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice));
    IPM_CUDA_OK(cudaMemcpy(db, b, (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
    IPM_CUDA_OK(cudaMemcpy(dc, c, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    BatchArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.b = db; ba.c = dc; ba.x = dx; ba.s = ds; ba.y = dy; ba.scal = d.scal; ba.active = d.ints; ba.iters = d.ints + 1;
    ba.m = m; ba.n = n;
    kb_init<<<1, 256>>>(ba, 1, 0);                           // |b|, |c|; x = s = 1, y = 0 (main.py:287-302)
    count_launch();
    KktArgs k;
    k.A = d.A; k.b = db; k.c = dc; k.x = dx; k.s = ds; k.y = dy; k.scal = d.scal; k.iters = d.ints + 1; k.list = nullptr;
    k.work = d.work; k.m = m; k.n = n; k.tol = tol; k.eta = 0.91; k.max_iter = max_iter;
    k.prof = reinterpret_cast<long long*>(d.ints + 4);
    IPM_TRY(ka_launch(k, 1, 0));
    double hs[S_COUNT];
    int hi[4];
    IPM_CUDA_OK(cudaMemcpy(hs, d.scal, sizeof(hs), cudaMemcpyDeviceToHost));
    IPM_CUDA_OK(cudaMemcpy(hi, d.ints, sizeof(hi), cudaMemcpyDeviceToHost));
    IPM_CUDA_OK(cudaMemcpy(g_kkt_prof, d.ints + 4, sizeof(g_kkt_prof), cudaMemcpyDeviceToHost));
    if (x) IPM_CUDA_OK(cudaMemcpy(x, dx, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (y) IPM_CUDA_OK(cudaMemcpy(y, dy, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost));
    if (s) IPM_CUDA_OK(cudaMemcpy(s, ds, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    const bool finite = std::isfinite(hs[S_NRB]) && std::isfinite(hs[S_NRC]) && std::isfinite(hs[S_XS]) && std::isfinite(hs[S_OBJ]);
    if (obj) *obj = hs[S_OBJ];
    if (iters) *iters = hi[1];
    if (status) *status = !finite ? IPM_STATUS_NAN : (hs[S_CONT] > 0.5 ? IPM_STATUS_MAX_ITER : IPM_STATUS_CONVERGED);
    return IPM_OK;
}

int ipm_profile_enable(int on) {
    g_prof.collect();
    g_prof.enabled = on != 0;
    for (int i = 0; i < PH_COUNT; ++i) { g_prof.ms[i] = 0.0; g_prof.calls[i] = 0; }
    g_prof.lp_iterations = 0;
    return IPM_OK;
}

int ipm_profile_read(double ms[4], int64_t calls[4], int64_t* lp_iterations) {
    for (int i = 0; i < PH_COUNT; ++i) {
        if (ms) ms[i] = g_prof.ms[i];
        if (calls) calls[i] = g_prof.calls[i];
    }
    if (lp_iterations) *lp_iterations = g_prof.lp_iterations;
    return IPM_OK;
}

int ipm_syrk_batched_d(int device_ordinal, int B, int m, int n, const double* A_d, const double* d_d, double* M_d,
                       int64_t ldm) {
    if (!A_d || !M_d) return IPM_ERR_ARG;
    if (B <= 0 || m <= 0 || n <= 0 || ldm < m) return IPM_ERR_SHAPE;
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    DmmaArgs g;
    g.P = A_d; g.ldp = n; g.strideP = (int64_t)m * n;
    g.Q = A_d; g.ldq = n; g.strideQ = (int64_t)m * n;
    g.dvec = d_d; g.strideD = n;
    g.C = M_d; g.ldc = ldm; g.strideC = (int64_t)m * ldm;
    g.rowsP = m; g.rowsQ = m; g.K = n; g.lower_only = 1; g.active = nullptr;
    return dmma_syrk_auto<0>(g, B, 0);
}

int ipm_potrf_batched_d(int device_ordinal, int B, int m, double* M_d, int64_t ldm, int64_t strideM,
                        double pivot_rel_thresh, int* n_fixed_total) {
    if (!M_d) return IPM_ERR_ARG;
    if (B <= 0 || m <= 0 || ldm < m || (ldm & 1) || (strideM & 1) || strideM < (int64_t)m * ldm ||
        (reinterpret_cast<uintptr_t>(M_d) & 15))
        return IPM_ERR_SHAPE;          // rows must start on 16-byte boundaries (vector loads, bulk prefetch)
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    double* scal = nullptr;
    IPM_CUDA_OK(cudaMalloc(&scal, (size_t)B * S_COUNT * sizeof(double)));
    int rc = (m <= KBC_MAX_M_BIG)
                 ? potrf_batched_fused(M_d, ldm, strideM, m, B, scal, S_COUNT, pivot_rel_thresh, nullptr, 0)
                 : potrf_blocked<64, 256, 128>(M_d, ldm, strideM, m, B, scal, S_COUNT, pivot_rel_thresh, nullptr, 0);
    if (rc == IPM_OK) {
        std::vector<double> hs((size_t)B * S_COUNT);
        cudaError_t e = cudaMemcpy(hs.data(), scal, hs.size() * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); rc = IPM_ERR_CUDA; }
        else if (n_fixed_total) {
            int tot = 0;
            for (int i = 0; i < B; ++i) tot += (int)hs[(size_t)i * S_COUNT + S_NFIXED];
            *n_fixed_total = tot;
        }
    }
    cudaFree(scal);
    return rc;
}

int ipm_profile_last(double* ms, int* phase, int cap) {
    const int n = (int)g_prof.last_ms.size();
    for (int i = 0; i < n && i < cap; ++i) {
        if (ms) ms[i] = g_prof.last_ms[i];
        if (phase) phase[i] = g_prof.last_ph[i];
    }
    return n;
}

int64_t ipm_batched_workspace_bytes(int B, int m, int n) {
    if (B <= 0 || m <= 0 || n <= 0) return 0;
    return ws_bytes(B, m, n);
}

int ipm_solve_batched_dense_d(int device_ordinal, int B, int m, int n, const double* A_d, const double* b_d,
                              const double* c_d, double tol, int max_iter, double* obj_d, int* iters_d,
                              int* status_d, double* x_d, void* work_d, int* iterations_run) {
    if (!A_d || !b_d || !c_d || max_iter < 0) return IPM_ERR_ARG;
    if (work_d && (reinterpret_cast<uintptr_t>(work_d) & 15)) {
        g_last_error = "work_d must be 16-byte aligned";
        return IPM_ERR_ARG;
    }
    IPM_TRY(check_shape(B, m, n));
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    void* own = nullptr;
    if (!work_d) {
        IPM_CUDA_OK(cudaMalloc(&own, (size_t)ws_bytes(B, m, n)));
        work_d = own;
    }
    int rc = [&]() -> int {
        // (the pinned read-back page comes from the loop's resource pool: no cudaMallocHost / cudaFreeHost per call)
        IPM_TRY(solve_on_device(B, m, n, A_d, b_d, c_d, tol, max_iter, obj_d, iters_d, status_d, x_d, work_d, nullptr,
                                0, iterations_run));
        IPM_CUDA_OK(cudaStreamSynchronize(0));
        return IPM_OK;
    }();
    g_prof.collect();
    if (own) cudaFree(own);
    return rc;
}

// Device-side staging of the host-buffer entry point, cached per process (one context per device) so that
// repeated calls do not pay cudaMalloc / cudaMallocHost / stream creation again.
struct HostPathCtx {
    int dev = -1;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    std::vector<cudaEvent_t> landed;                 // one per chunk: its H2D copy has completed
    double *dA = nullptr, *db = nullptr, *dc = nullptr;
    double *d_obj = nullptr, *d_x = nullptr;
    int *d_it = nullptr, *d_st = nullptr;
    void* work = nullptr;
    unsigned* h_nact = nullptr;
    int64_t capA = 0, capb = 0, capc = 0, capres = 0, capx = 0, capwork = 0;   // bytes
    void release() {
        if (dev < 0) return;
        cudaSetDevice(dev);
        for (cudaEvent_t e : landed) cudaEventDestroy(e);
        landed.clear();
        void* ptrs[] = {dA, db, dc, d_obj, d_it, d_st, d_x, work};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        if (h_nact) cudaFreeHost(h_nact);
        if (s_copy) cudaStreamDestroy(s_copy);
        if (s_comp) cudaStreamDestroy(s_comp);
        *this = HostPathCtx();
    }
};
static HostPathCtx g_host_ctx[16];

static int ensure_bytes(void** p, int64_t* cap, int64_t need) {
    if (*cap >= need) return IPM_OK;
    if (*p) { IPM_CUDA_OK(cudaFree(*p)); *p = nullptr; *cap = 0; }
    IPM_CUDA_OK(cudaMalloc(p, (size_t)need));
    *cap = need;
    return IPM_OK;
}

int ipm_release_cached(void) {
    for (auto& c : g_host_ctx) c.release();
    drop_loopres_pool();
    ipm_pattern_cache_clear();
    return IPM_OK;
}

int ipm_solve_batched_dense(int device_ordinal, int B, int m, int n, const double* A, const double* b,
                            const double* c, double tol, int max_iter, double* obj, int* iters, int* status,
                            double* x) {
    if (!A || !b || !c || max_iter < 0) return IPM_ERR_ARG;
    IPM_TRY(check_shape(B, m, n));
    if (device_ordinal < 0 || device_ordinal >= 16) return IPM_ERR_ARG;
    // one solve at a time per device through this entry point: the staging buffers and streams of a device are
    // shared by every caller (distinct devices run concurrently from distinct threads)
    static std::mutex host_ctx_mu[16];
    std::lock_guard<std::mutex> host_ctx_lock(host_ctx_mu[device_ordinal]);
    IPM_CUDA_OK(cudaSetDevice(device_ordinal));
    // The whole batch gets device buffers; it is copied in chunks of 128 LPs on a copy stream (small chunks: an LP can
    // join 2.5 ms after its predecessor instead of 20, so the loop has more LPs to work on while the copy is ahead of
    // it by little - 1024-LP chunks cost 9 ms per 8192-LP solve, tools/e2e_trace.py) and ONE lockstep loop runs from
    // the moment the first chunk has landed:
    // the LPs of a chunk join the loop at the first iteration after their copy has completed (Arrival).  Solving
    // chunk after chunk instead paid the latency-bound last iterations of a lockstep solve once per chunk.
    std::vector<int> first_of, count_of;
    {
        static const int chunk_lps = [] {
            const char* e = getenv("IPM_E2E_CHUNK");           // LPs per chunk after the first (A/B measurements)
            const int v = e ? atoi(e) : 0;
            return (v >= 32 && v <= 65536) ? v : 128;
        }();
        int next = std::min(B, std::min(chunk_lps, 256)), at = 0;
        while (at < B) {
            const int cnt = std::min(next, B - at);
            first_of.push_back(at);
            count_of.push_back(cnt);
            at += cnt;
            next = chunk_lps;
        }
    }
    const int nchunks = (int)first_of.size();
    HostPathCtx& C = g_host_ctx[device_ordinal];
    int rc = [&]() -> int {
        if (C.dev < 0) {
            C.dev = device_ordinal;
            IPM_CUDA_OK(cudaStreamCreateWithFlags(&C.s_copy, cudaStreamNonBlocking));
            IPM_CUDA_OK(cudaStreamCreateWithFlags(&C.s_comp, cudaStreamNonBlocking));
            IPM_CUDA_OK(cudaMallocHost(&C.h_nact, 8 * sizeof(unsigned)));
        }
        while ((int)C.landed.size() < nchunks) {
            cudaEvent_t e = nullptr;
            IPM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            C.landed.push_back(e);
        }
        IPM_TRY(ensure_bytes((void**)&C.dA, &C.capA, (int64_t)B * m * n * 8));
        IPM_TRY(ensure_bytes((void**)&C.db, &C.capb, (int64_t)B * m * 8));
        IPM_TRY(ensure_bytes((void**)&C.dc, &C.capc, (int64_t)B * n * 8));
        if (C.capres < (int64_t)B * 8) {
            if (C.d_obj) cudaFree(C.d_obj);
            if (C.d_it) cudaFree(C.d_it);
            if (C.d_st) cudaFree(C.d_st);
            C.d_obj = nullptr; C.d_it = C.d_st = nullptr; C.capres = 0;
            IPM_CUDA_OK(cudaMalloc(&C.d_obj, (size_t)B * 8));
            IPM_CUDA_OK(cudaMalloc(&C.d_it, (size_t)B * sizeof(int)));
            IPM_CUDA_OK(cudaMalloc(&C.d_st, (size_t)B * sizeof(int)));
            C.capres = (int64_t)B * 8;
        }
        if (x) IPM_TRY(ensure_bytes((void**)&C.d_x, &C.capx, (int64_t)B * n * 8));
        IPM_TRY(ensure_bytes(&C.work, &C.capwork, ws_bytes(B, m, n)));
        // the previous call's solve has drained (the call synchronises before returning): the buffers are free
        for (int k = 0; k < nchunks; ++k) {
            const size_t first = (size_t)first_of[k], cnt = (size_t)count_of[k];
            IPM_CUDA_OK(cudaMemcpyAsync(C.dA + first * m * n, A + first * m * n, cnt * m * n * 8, cudaMemcpyHostToDevice,
                                        C.s_copy));
            IPM_CUDA_OK(cudaMemcpyAsync(C.db + first * m, b + first * m, cnt * m * 8, cudaMemcpyHostToDevice, C.s_copy));
            IPM_CUDA_OK(cudaMemcpyAsync(C.dc + first * n, c + first * n, cnt * n * 8, cudaMemcpyHostToDevice, C.s_copy));
            IPM_CUDA_OK(cudaEventRecord(C.landed[k], C.s_copy));
        }
        Arrival arr;
        arr.nchunks = nchunks; arr.first = first_of.data(); arr.count = count_of.data(); arr.landed = C.landed.data();
        IPM_TRY(solve_on_device(B, m, n, C.dA, C.db, C.dc, tol, max_iter, C.d_obj, C.d_it, C.d_st, x ? C.d_x : nullptr,
                                C.work, C.h_nact, C.s_comp, nullptr, &arr));
        if (obj) IPM_CUDA_OK(cudaMemcpyAsync(obj, C.d_obj, (size_t)B * 8, cudaMemcpyDeviceToHost, C.s_comp));
        if (iters) IPM_CUDA_OK(cudaMemcpyAsync(iters, C.d_it, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, C.s_comp));
        if (status) IPM_CUDA_OK(cudaMemcpyAsync(status, C.d_st, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, C.s_comp));
        if (x) IPM_CUDA_OK(cudaMemcpyAsync(x, C.d_x, (size_t)B * n * 8, cudaMemcpyDeviceToHost, C.s_comp));
        IPM_CUDA_OK(cudaStreamSynchronize(C.s_comp));
        IPM_CUDA_OK(cudaStreamSynchronize(C.s_copy));
        return IPM_OK;
    }();
    g_prof.collect();
    if (rc != IPM_OK) C.release();
    return rc;
}

}  // extern "C"
