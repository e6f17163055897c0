// Batched direction kernels of the 3-pass iteration (m <= 256): one CTA per LP walks A_i in column strips of 16
// (m x 16 doubles, a ring of six buffers in shared memory filled by bulk copies, four strips in flight), so that ONE trip of A_i through HBM serves
// both products every direction needs:
//     u = A^T dy          (column sums over the strip, needed first)
//     q = A e             (row sums, where e depends on u through the elementwise direction formulas)
// kind 0 (predictor, main.py:225-228, 305-322, 582-600):
//     dxa = d u + w, dsa = -s dxa / x - rcx, ratio test, mu_aff, sigma, and the corrector right-hand side by
//     linearity of main.py:150-152 in r4 = x s + dxa dsa - sigma mu:
//         rhs_corr = -rb - A d (rc - r4/x) = rhs_pred + A (d dxa dsa / x) - sigma mu A (d / x)
// kind 1 (corrector, main.py:142-159, 604-626, 694-696):
//     dx, ds, eta-damped ratio test, x += ap dx, y += ad dy, s += ad ds, and the residuals of the NEW point from
//     q = A dx and u:   rb += ap q,  rc += ad (u + ds)   (exact identities; check_optimality main.py:169-173 is
//     re-evaluated from scratch by kb_residual for every LP these recurrences declare converged),
//     then d = x/s and the predictor operand w = d (rc - (x s)/x) for the next assembly.
//     Conditional refinement (a.refine): the reference solves the unreduced system, whose second block row IS
//     A dx = -rb (main.py:13-21); the normal equations satisfy it only as well as the factor of M allows, and once
//     max d / min d passes 1e19 that is not well enough for |rb| to keep falling - an LP that has not met
//     check_optimality by then stays trapped at the boundary (generator LPs 16893, 31186).  q = A dx is already
//     here, so delta = -rb - q is free: when |delta| > 0.1 |rb| - a tenth of what the step is meant to remove
//     is put back by the solve's error - the LP is not updated; delta goes to a.rhs, the LP is flagged FLAG_REFINE, and the host's next two
//     launches (triangular solve ddy = M^-1 delta in place, then this kernel again with pass = 1, both for flagged
//     LPs only) apply the step INCREMENTALLY: dx += d (A^T ddy), ds re-formed from dx, dy += ddy.  (Re-forming dx
//     from the refined dy would bring back the cancellation noise eps d_max |t| of the large-d columns, which is
//     what the step removes.)  If the refined direction STILL has |delta| > |rb|, the normal equations have broken
//     down for this LP - a safeguarded pivot on a row that is not dependent: numerically degenerate vertex - and
//     the LP is parked (a.handoff_list) for the augmented-system kernel (kkt_dense.cuh), which continues it from
//     its current iterate.  On the CPU restatement the step is taken by 1.4 % of the generator's LPs, 0.1 % are
//     handed off, and all 65536 finish in the reference's iteration count +-1 (tests/golden/batch_256x512_*).
// With the predictor right-hand side (kb_rhs, one pass) an iteration reads A four times instead of six.
#pragma once
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)

#include "common.cuh"
#include "dmma_ws.cuh"

namespace ipm {

constexpr int KF_NT = 512;                  // streaming threads
constexpr int KF_NTT = KF_NT + 32;          // + warp E (elementwise formulas)
constexpr int KF_W = 16;                    // strip width (columns)
constexpr int KF_RG = KF_NT / KF_W;         // 32 row groups
constexpr int KF_RPT = 8;                   // rows per thread
constexpr int KF_MAX_M = KF_RG * KF_RPT;    // 256

struct BatchArgs {
    const double* A;   // [B][m][n]
    double* At;        // [B][ceil(n/16)][32*nrp][16]  strip-major copy of A (3-pass path, kbf_repack)
    const double* b;   // [B][m]
    const double* c;   // [B][n]
    double *x, *s, *rc, *d, *w, *rcx, *dxa, *dsa;      // [B][n]
    double *dxc, *dsc;                                 // [B][n] corrector direction (the predictor's stays intact
                                                       // for a refinement pass)
    double *y, *rb, *dy, *rhs;                         // [B][m]
    double* scal;      // [B][S_COUNT]
    int* active;       // [B]   0 = finished, 1 = iterating, 2 = to be checked from scratch (3-pass path)
    int* iters;        // [B]
    unsigned* n_active;
    int m, n;
    double tol, eta;
    int max_iter;
    int fresh_every;   // 3-pass path: residuals are re-evaluated from scratch every fresh_every-th iteration (0 = only
                       // when the recurrences report convergence)
    int refine;        // 1: conditional refinement of the corrector (see kbf_dir / kb_dir), 0: off
    int handoff;       // 1: an LP whose refined corrector still violates A dx = -rb leaves the loop (handoff_list)
    int* handoff_list; // [B] LP indices parked for the augmented-system kernel (kkt_dense.cuh)
    unsigned* n_handoff;
    int* act_list;     // [B] LPs the last residual check left active, in the order their CTAs got there (n_active entries)
    unsigned* kf_ctr;  // [2] work counter of the persistent direction kernels: next LP to hand out, CTAs that have left
                       // (zero between launches: the last CTA to leave resets both)
};
// Refinement threshold: |delta| > KF_REFINE_THRESH |rb|.  Measured against the UNMODIFIED reference on 517 generator LPs
// (CPU restatement): 1.0 -> objectives within 1.2e-8 relative, 0.3 -> 1.7e-9, 0.1 -> 1.6e-10 (iteration counts equal in all
// three); the step is then taken 0.06 times per LP (3.5 % of the LPs).
constexpr double KF_REFINE_THRESH = 0.1;
constexpr int FLAG_REFINE = 3;    // active[] value between the corrector pass that asks for a refinement and the one that
                                  // applies it

// Strip buffers: row sums of s-2, s-1 (pending), column sums of s, and nbuf - 2 loads ahead.  Six buffers (four strips =
// 128 KB on the wire per SM) when they fit beside the per-column coefficients, five otherwise (m > 128 with n > 640).
constexpr int KF_NBUF_MAX = 6;
constexpr size_t KF_SMEM_CAP = 227 * 1024 - 2048;      // opt-in maximum per block minus the kernel's static shared memory

inline int kf_nrp(int m) { return m <= 32 ? 1 : m <= 64 ? 2 : m <= 128 ? 4 : 8; }     // rows per thread
inline size_t kf_smem_bytes(int m, int n, int nbuf) {
    const int mr = 32 * kf_nrp(m);
    const int npad = (n + KF_W - 1) / KF_W * KF_W;
    // (the row sums q1, q2 of the epilogue live in the buffer of the LP's last strip, which nobody reads by then)
    return (size_t)(nbuf * mr * KF_W + mr + 4 * (KF_NT / 32) * KF_W + 8 * KF_W + 5 * npad) * sizeof(double);
}
inline int kf_nbuf(int m, int n) { return kf_smem_bytes(m, n, KF_NBUF_MAX) <= KF_SMEM_CAP ? KF_NBUF_MAX : KF_NBUF_MAX - 1; }

#ifdef __CUDACC__
// One CTA per SM, 17 warps.  Thread 0 keeps KF_NBUF - 2 strips (four of 32 KB at the benchmark shape) in flight with
// bulk copies into a ring of KF_NBUF buffers (six when they fit, else five: kf_nbuf), each with its own mbarrier.  Per strip s the 16 streaming
// warps form the column sums (-> part[s & 3], mbarrier pfull), warp E turns them into dx, ds and the row-sum
// operand e (-> ev[s & 3], mbarrier efull) and the streaming warps add the row sums of strip s-2, so E's chain
// of dependent FP64 operations (long latency on this part) is two strips off the critical path.
// NRP = rows per thread is a template parameter (rows are padded to 32 NRP with zeros), which turns every
// shared-memory offset of the inner loops into an immediate: the first version was issue-bound on index
// arithmetic (ncu: 300 warp instructions per strip and warp, 75 % of them integer).
// SRC = 1 (default): the strips come straight from the caller's row-major A[B][m][n] through a 3-D tensor map
// (box 16 columns x MR rows x 1 LP, cp.async.bulk.tensor -> UTMALDG; rows >= m and columns >= n are zero-filled by
// the hardware), SRC = 0: from a strip-major copy of A made once per solve by kbf_repack (one untiled bulk copy per
// strip, UBLKCP) - kept for A/B measurements, it costs a second copy of A in HBM and 4.4 ms per 8192 LPs.
// PERSISTENT: one CTA per SM takes LPs from a work counter (a.kf_ctr) over the list of active LPs, so that (1) a
// lockstep iteration on a few hundred active LPs does not start thousands of CTAs that only find a zero flag, and (2) the first KF_AHEAD strips of the NEXT LP are
// requested during the last strips of the current one: the bulk copies are in flight across the epilogue (ratio
// test, update, norms) and the next prologue, where a one-LP-per-CTA launch left HBM idle for 17 % of its time.
// The ring position (buffer, mbarrier phases) simply continues from one LP to the next.
template <int KIND, int NRP, int SRC>
__global__ void __launch_bounds__(KF_NTT, 1) kbf_dir(const BatchArgs a, const int pass, const int nlp, const int KF_NBUF,
                                                     const __grid_constant__ CUtensorMap tmapA) {
    const int KF_AHEAD = KF_NBUF - 2;             // strips in flight
    constexpr int MR = 32 * NRP;                  // padded rows
    constexpr int SB = MR * KF_W;                 // doubles per strip buffer
    extern __shared__ __align__(128) double smem_kf[];        // strip buffers first: TMA destinations, 128-byte aligned
    __shared__ double sh[32];
    __shared__ double s_val[4];
    __shared__ __align__(8) uint64_t full[KF_NBUF_MAX], pfull[4], efull[4];
    __shared__ int s_next;
    const int m = a.m, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int mr = MR;
    double* const dxo = (KIND == 0) ? a.dxa : a.dxc;       // where this pass leaves its direction
    double* const dso = (KIND == 0) ? a.dsa : a.dsc;
    double* strip = smem_kf;                      // [KF_NBUF][MR][16]
    double* dys = strip + KF_NBUF * SB;           // [MR]   dy
    double* part = dys + MR;                      // [4][16 warps][16] column partial sums (ring of 4 strips)
    double* ev1 = part + 4 * (KF_NT / 32) * KF_W; // [4][16]
    double* ev2 = ev1 + 4 * KF_W;                 // [4][16]
    double* gsm = ev2 + 4 * KF_W;                 // [5][npad] per-column coefficients that do not depend on u
    const int c = tid & (KF_W - 1), rg = tid >> 4;
    const int nstrips = (n + KF_W - 1) / KF_W;

    // strip s of this LP is one contiguous block of the strip-major copy: a single bulk (TMA) copy per strip,
    // completion signalled on the buffer's mbarrier
    if (tid == 0) {
        for (int k = 0; k < KF_NBUF; ++k) mbar_init(full + k, 1);
        for (int k = 0; k < 4; ++k) { mbar_init(pfull + k, KF_NT / 32); mbar_init(efull + k, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int lp, int sidx, int boff_d, int slot) {     // thread 0 only
        const double* At = (SRC == 0) ? a.At + (size_t)lp * nstrips * SB : nullptr;
        const uint32_t bar = smem_u32(full + slot);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SB * 8)) : "memory");
        if (SRC == 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(strip + boff_d)), "l"(At + (size_t)sidx * SB), "r"((uint32_t)(SB * 8)), "r"(bar)
                         : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(strip + boff_d)), "l"(reinterpret_cast<uint64_t>(&tmapA)), "r"(sidx * KF_W), "r"(0),
                           "r"(lp), "r"(bar)
                         : "memory");
        }
    };
    // ---- work distribution (warp E, which has slack at the head of an LP).  pass 0: one entry at a time from the list
    // of active LPs the residual check of this iteration compiled (a.act_list, *a.n_active entries): no scan, and a
    // launch on a few LPs spreads them over the SMs.  pass 1 (refinement, about one LP in seventy): the flags are
    // scanned, eight indices per atomic, the LPs of a group that ask for the pass are kept in a bit mask.
    int grp_base = 0;
    unsigned grp_mask = 0u;
    const unsigned nact = (pass == 0) ? *a.n_active : 0u;
    auto grab = [&]() -> int {                       // all lanes of warp E; -1: no LP left.  Blocks on two round trips.
        if (pass == 0) {
            int v = -1;
            if (lane == 0) {
                const unsigned idx = atomicAdd(a.kf_ctr, 1u);
                if (idx < nact) v = a.act_list[idx];
                if (v >= nlp) v = -1;
            }
            return __shfl_sync(0xffffffffu, v, 0);
        }
        for (;;) {
            if (grp_mask) {
                const int bit = __ffs(grp_mask) - 1;
                grp_mask &= grp_mask - 1;
                return grp_base + bit;
            }
            unsigned idx = (lane == 0) ? atomicAdd(a.kf_ctr, 8u) : 0u;
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= (unsigned)nlp) return -1;
            bool want = false;
            if (lane < 8 && (int)idx + lane < nlp) want = a.active[idx + lane] == FLAG_REFINE;
            grp_mask = __ballot_sync(0xffffffffu, want);
            grp_base = (int)idx;
        }
    };
    // pass 0, enough strips: the two round trips of the grab are spread over the first strips of warp E's loop (atomic
    // at strip 0, list entry read at strip 3, published at strip 6), so that nobody ever waits for them
    const bool lazy_grab = (pass == 0) && nstrips >= 12;
    // ---- ring state, carried from one LP to the next
    auto nextb = [&](int bo) { return (bo + SB == KF_NBUF * SB) ? 0 : bo + SB; };
    auto wrap = [&](int k) { return k >= KF_NBUF ? k - KF_NBUF : k; };            // k < 2 KF_NBUF
    int b_prev2 = (KF_NBUF - 2) * SB, b_prev = (KF_NBUF - 1) * SB, b_cur = 0;      // buffers of strips s-2, s-1, s
    int slot = 0;                                 // mbarrier of strip s (= its buffer index)
    uint32_t parity = 0;
    uint32_t gs = 0;                              // strips this CTA has been through: position in the part / ev rings
    bool prefetched = false;                      // the first KF_AHEAD strips of the coming LP are already on their way
    const bool xlp = nstrips >= 12;               // (s_next, out at strip 6, must be public KF_AHEAD strips before the end)
    if (warp == KF_NT / 32) {
        const int v = grab();
        if (lane == 0) s_next = v;
    }
    __syncthreads();
    int lp = s_next;
    __syncthreads();                              // everyone has read it: warp 0 writes the next one right away

    auto one_lp = [&](const int lp, int& next_lp) {
    const size_t on = (size_t)lp * n, om = (size_t)lp * m;
    double* scal = a.scal + (size_t)lp * S_COUNT;
    if (tid == 0 && !prefetched) {
        for (int k = 0; k < KF_AHEAD && k < nstrips; ++k) issue(lp, k, wrap(slot + k) * SB, wrap(slot + k));
    }
    // pass 1 streams ddy (the in-place refinement solve left it in a.rhs); dy itself is only needed for y
    const double* dy_src = (pass == 0) ? a.dy : a.rhs;
    for (int i = tid; i < mr; i += KF_NTT) dys[i] = (i < m) ? dy_src[om + i] : 0.0;
    const double sigma_mu = (KIND == 1) ? scal[S_SIGMA_MU] : 0.0;
    // Everything of the elementwise formulas that does not depend on u (it holds the divisions), for all columns,
    // by all threads, while the first strips are on their way:
    //   kind 0: dx = d u + w,        ds = g2 dx - rcx,   e1 = g3 dx ds, e2 = g3      (g2 = -s/x, g3 = d/x)
    //   kind 1: dx = d u + w_corr,   ds = g2 dx - rcx_corr
    const int npad = nstrips * KF_W;
    for (int k = tid; k < npad; k += KF_NTT) {
        double di = 0.0, v0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
        if (k < n) {
            const double xi = a.x[on + k], si = a.s[on + k];
            di = a.d[on + k];
            g2 = -si / xi;
            if (KIND == 0) {
                v0 = a.w[on + k];
                g1 = a.rcx[on + k];
                g3 = di / xi;
            } else {
                const double rcomp = xi * si + a.dxa[on + k] * a.dsa[on + k] - sigma_mu;
                g1 = rcomp / xi;
                v0 = (pass == 0) ? di * (a.rc[on + k] - g1) : a.dxc[on + k];      // pass 1: dx = dx_old + d (A^T ddy)
            }
        }
        gsm[k] = di; gsm[npad + k] = v0; gsm[2 * npad + k] = g1; gsm[3 * npad + k] = g2; gsm[4 * npad + k] = g3;
    }
    __syncthreads();

    double acc1[NRP], acc2[NRP];
#pragma unroll
    for (int i = 0; i < NRP; ++i) acc1[i] = acc2[i] = 0.0;

    const int toff = rg * KF_W + c;               // this thread's element inside a strip buffer
    // Role split.  Warps 0..15 (512 threads) stream the strips: column sums of strip s, row sums of strip s-2.
    // Warp 16 (E) turns the column sums of a strip into the direction entries of its 16 columns and runs up to two
    // strips behind the streaming warps, so its chain of dependent FP64 operations never stalls them.
    auto stream_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(KF_NT) : "memory"); };

    if (warp == KF_NT / 32) {
        // ------------------------------------------------------------------ E: u -> dx, ds, e
        // ... and the LP after this one -> s_next (read by thread 0 behind an efull wait, by everyone behind the
        // barrier that ends the streaming)
        unsigned g_idx = 0u;
        int g_lp = -1;
        if (!lazy_grab) {
            const int v = grab();
            if (lane == 0) s_next = v;
        }
        for (int sidx = 0; sidx < nstrips; ++sidx) {
            if (lazy_grab && lane == 0) {
                if (sidx == 0) g_idx = atomicAdd(a.kf_ctr, 1u);
                if (sidx == 3 && g_idx < nact) g_lp = a.act_list[g_idx];
                if (sidx == 6) s_next = (g_lp >= 0 && g_lp < nlp) ? g_lp : -1;
            }
            const uint32_t ge = gs + (uint32_t)sidx;
            const int j = ge & 3;
            mbar_wait(pfull + j, (ge >> 2) & 1u);
            // u: 16 warp partials per column, two lanes per column (8 each, pairwise), fixed order
            const double* pp = part + j * (KF_NT / 32) * KF_W + (lane >> 4) * 8 * KF_W + (lane & 15);
            const double t0 = pp[0] + pp[KF_W], t1 = pp[2 * KF_W] + pp[3 * KF_W];
            const double t2 = pp[4 * KF_W] + pp[5 * KF_W], t3 = pp[6 * KF_W] + pp[7 * KF_W];
            double u = (t0 + t1) + (t2 + t3);
            u += __shfl_xor_sync(0xffffffffu, u, 16);
            const int col = sidx * KF_W + lane;
            if (lane < KF_W) {
                const double* gq = gsm + col;
                const double di = gq[0], v0 = gq[npad], g1 = gq[2 * npad], g2 = gq[3 * npad], g3 = gq[4 * npad];
                const double dxi = di * u + v0;
                const double dsi = g2 * dxi - g1;
                double e1, e2 = 0.0;
                if (KIND == 0) { e1 = g3 * (dxi * dsi); e2 = g3; }
                else e1 = dxi;
                if (col < n) {
                    // A^T dy + ds (change of rc per unit dual step); pass 1: A^T (dy + ddy) = (w_old - ds_old) + u
                    if (KIND == 1) a.w[on + col] = (pass == 0) ? (u + dsi) : ((a.w[on + col] - a.dsc[on + col]) + u + dsi);
                    dxo[on + col] = dxi;
                    dso[on + col] = dsi;
                } else {
                    e1 = e2 = 0.0;
                }
                ev1[j * KF_W + lane] = e1;
                ev2[j * KF_W + lane] = e2;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(efull + j);
        }
    } else {
        // ------------------------------------------------------------------ streaming warps
        auto row_sums = [&](int boff_d, int sidx) {
            const uint32_t ge = gs + (uint32_t)sidx;
            const int j = ge & 3;
            mbar_wait(efull + j, (ge >> 2) & 1u);
            const double* sb = strip + boff_d + toff;
            const double f1 = ev1[j * KF_W + c], f2 = ev2[j * KF_W + c];
#pragma unroll
            for (int i = 0; i < NRP; ++i) {
                const double v = sb[i * KF_RG * KF_W];
                acc1[i] = fma(v, f1, acc1[i]);
                if (KIND == 0) acc2[i] = fma(v, f2, acc2[i]);
            }
        };
        // buffer offsets (doubles) of strips s-2, s-1, s and s+KF_AHEAD rotate through the ring
        for (int sidx = 0; sidx < nstrips; ++sidx) {
            mbar_wait(full + slot, parity);                     // strip sidx has landed
            {
                const double* sb = strip + b_cur + toff;
                const double* dq = dys + rg;
                double p0 = 0.0, p1 = 0.0;                      // two chains: FP64 latency is long on this part
#pragma unroll
                for (int i = 0; i < NRP; i += 2) {
                    p0 = fma(sb[i * KF_RG * KF_W], dq[KF_RG * i], p0);
                    if (i + 1 < NRP) p1 = fma(sb[(i + 1) * KF_RG * KF_W], dq[KF_RG * (i + 1)], p1);
                }
                double p = p0 + p1;
                p += __shfl_xor_sync(0xffffffffu, p, 16);       // the warp's two row groups
                const int jp = (gs + (uint32_t)sidx) & 3;
                if (lane < KF_W) part[jp * (KF_NT / 32) * KF_W + warp * KF_W + lane] = p;
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull + jp);
            }
            if (sidx >= 2) row_sums(b_prev2, sidx - 2);
            stream_bar();                                       // strip sidx-2 is no longer read by anyone
            if (tid == 0) {
                const int ns = sidx + KF_AHEAD;
                if (ns < nstrips) issue(lp, ns, b_prev2, wrap(slot + KF_AHEAD));
                else if (xlp) {
                    const int nl = *reinterpret_cast<volatile int*>(&s_next);     // public since strip 6 (efull waits)
                    if (nl >= 0) issue(nl, ns - nstrips, b_prev2, wrap(slot + KF_AHEAD));
                }
            }
            b_prev2 = b_prev; b_prev = b_cur; b_cur = nextb(b_cur);
            if (++slot == KF_NBUF) { slot = 0; parity ^= 1u; }
        }
        if (nstrips >= 2) row_sums(b_prev2, nstrips - 2);
        row_sums(b_prev, nstrips - 1);
    }
    __syncthreads();
    next_lp = s_next;
    // row sums of the whole LP: in the buffer of its last strip (free now; the next LP's first strips land in others)
    double* const q1s = strip + (size_t)((gs + (uint32_t)nstrips - 1u) % (uint32_t)KF_NBUF) * SB;     // [MR]   A e1
    double* const q2s = q1s + MR;                                                                  // [MR]   A e2 (kind 0)
    // row sums: combine the 16 column lanes of every row group
#pragma unroll
    for (int i = 0; i < NRP; ++i) {
        double v = acc1[i], v2 = acc2[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            if (KIND == 0) v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        const int r = rg + KF_RG * i;
        if (tid < KF_NT && c == 0 && r < mr) { q1s[r] = v; q2s[r] = v2; }
    }
    // ratio test over all columns (main.py:305-322)
    double minp = 1.0, mind = 1.0;
    for (int k = tid; k < n; k += KF_NTT) {
        const double dxi = dxo[on + k], dsi = dso[on + k];
        if (dxi < 0.0) minp = fmin(minp, -a.x[on + k] / dxi);
        if (dsi < 0.0) mind = fmin(mind, -a.s[on + k] / dsi);
    }
    minp = block_red<RED_MIN>(minp, sh);
    if (tid == 0) s_val[0] = minp;
    mind = block_red<RED_MIN>(mind, sh);
    if (tid == 0) s_val[1] = mind;
    __syncthreads();
    double ap = s_val[0], ad = s_val[1];
    if (KIND == 0) {
        double partsum = 0.0;
        for (int k = tid; k < n; k += KF_NTT)
            partsum += (a.x[on + k] + ap * a.dxa[on + k]) * (a.s[on + k] + ad * a.dsa[on + k]);
        partsum = block_red<RED_SUM>(partsum, sh);
        if (tid == 0) {
            const double mu_aff = partsum / (double)n, mu = scal[S_XS] / (double)n;
            const double r = mu_aff / mu, sigma = r * r * r;
            scal[S_AP_AFF] = ap; scal[S_AD_AFF] = ad; scal[S_MU_AFF] = mu_aff; scal[S_MU] = mu;
            scal[S_SIGMA] = sigma; scal[S_SIGMA_MU] = sigma * mu;
            s_val[2] = sigma * mu;
        }
        __syncthreads();
        const double sm = s_val[2];
        for (int i = tid; i < m; i += KF_NTT) a.rhs[om + i] = a.rhs[om + i] + q1s[i] - sm * q2s[i];
    } else {
        if (a.refine) {
            // delta = -rb - A dx against rb (both squared norms in index order: deterministic)
            double nd2 = 0.0, nr2 = 0.0;
            for (int i = tid; i < m; i += KF_NTT) {
                const double r = a.rb[om + i], dl = -r - q1s[i];
                nd2 += dl * dl;
                nr2 += r * r;
            }
            nd2 = block_red<RED_SUM>(nd2, sh);
            if (tid == 0) s_val[2] = nd2;
            nr2 = block_red<RED_SUM>(nr2, sh);
            if (tid == 0) {
                // NaN compares false: no refinement.  floor: 1e-3 of the stopping threshold of |rb| (main.py:170) -
                // below it the primal residual is converged whatever delta does
                const double fl = 1e-3 * a.tol * (1.0 + scal[S_NB]);
                // pass 0: refine when |delta| > 0.1 |rb|; pass 1: hand off when the refined step still has |delta| > |rb|
                const double lim2 = (pass == 0) ? KF_REFINE_THRESH * KF_REFINE_THRESH * nr2 : nr2;
                s_val[3] = (s_val[2] > lim2 && s_val[2] > fl * fl) ? 1.0 : 0.0;
                // test hooks (ipm_batched_set_option value 2): refine every corrector / hand every LP off
                if ((pass == 0 && a.refine == 2) || (pass != 0 && a.handoff == 2)) s_val[3] = 1.0;
            }
            __syncthreads();
            if (s_val[3] != 0.0) {
                if (pass == 0) {
                    for (int i = tid; i < m; i += KF_NTT) a.rhs[om + i] = -a.rb[om + i] - q1s[i];
                    if (tid == 0) {
                        a.active[lp] = FLAG_REFINE;
                        scal[S_NREFINE] = scal[S_NREFINE] + 1.0;
                    }
                    return;
                }
                if (a.handoff) {
                    // the refined corrector still does not restore A dx = -rb: park the LP (no update)
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
        double nrc2 = 0.0, xs = 0.0, obj = 0.0, nrb2 = 0.0;
        for (int k = tid; k < n; k += KF_NTT) {
            const double xn = a.x[on + k] + ap * a.dxc[on + k];
            const double sn = a.s[on + k] + ad * a.dsc[on + k];
            const double rcn = a.rc[on + k] + ad * a.w[on + k];
            const double dn = xn / sn;
            const double q = (xn * sn) / xn;
            a.x[on + k] = xn; a.s[on + k] = sn; a.rc[on + k] = rcn; a.d[on + k] = dn;
            a.rcx[on + k] = q;
            a.w[on + k] = dn * (rcn - q);
            nrc2 += rcn * rcn;
            xs += xn * sn;
            obj += xn * a.c[on + k];
        }
        for (int i = tid; i < m; i += KF_NTT) {
            double dyi = dys[i];
            if (pass != 0) {                       // dys holds ddy: the step is dy + ddy
                dyi = a.dy[om + i] + dyi;
                a.dy[om + i] = dyi;
            }
            a.y[om + i] = a.y[om + i] + ad * dyi;
            const double r = a.rb[om + i] + ap * q1s[i];
            a.rb[om + i] = r;
            nrb2 += r * r;
        }
        nrb2 = block_red<RED_SUM>(nrb2, sh);
        if (tid == 0) s_val[2] = nrb2;
        nrc2 = block_red<RED_SUM>(nrc2, sh);
        xs = block_red<RED_SUM>(xs, sh);
        obj = block_red<RED_SUM>(obj, sh);
        if (tid == 0) {
            nrb2 = s_val[2];
            const double nrb = sqrt(nrb2), nrc = sqrt(nrc2);
            scal[S_NRB2] = nrb2; scal[S_NRB] = nrb; scal[S_NRC2] = nrc2; scal[S_NRC] = nrc;
            scal[S_XS] = xs; scal[S_OBJ] = obj; scal[S_AP] = ap; scal[S_AD] = ad;
            const bool cont = (a.tol * (1.0 + scal[S_NB]) < nrb) || (a.tol * (1.0 + scal[S_NC]) < nrc) || (a.tol < xs);
            scal[S_CONT] = cont ? 1.0 : 0.0;
            const int it = a.iters[lp] + 1;
            a.iters[lp] = it;
            // stop candidates (and the iteration cap) are decided by kb_residual on residuals computed from scratch
            a.active[lp] = (!cont || it >= a.max_iter || (a.fresh_every > 0 && it % a.fresh_every == 0)) ? 2 : 1;
        }
    }
    };      // one_lp

    while (lp >= 0) {
        int next_lp = -1;
        one_lp(lp, next_lp);
        gs += (uint32_t)nstrips;
        prefetched = xlp && next_lp >= 0;
        lp = next_lp;
        __syncthreads();                          // shared memory of this LP (and s_next) is free
    }
    if (tid == 0) {                               // the last CTA to leave re-arms the counter for the next launch
        const unsigned left = atomicAdd(a.kf_ctr + 1, 1u);
        if (left == gridDim.x - 1) {
            a.kf_ctr[0] = 0u;
            a.kf_ctr[1] = 0u;
        }
    }
}
// Strip-major copy of A for kbf_dir: At[lp][s][r][c] = A[lp][r][16 s + c], rows padded to mr = 32 nrp and columns to
// a multiple of 16 with zeros, so that a column strip is ONE contiguous block (DRAM pages are read whole and a
// strip is a single bulk copy).  One CTA per (strip, LP), one thread per row; runs once per solve.
__global__ void __launch_bounds__(256) kbf_repack(const BatchArgs a, int mr, int nstrips, int lp0) {
    const int s = blockIdx.x, lp = lp0 + blockIdx.y;
    const int m = a.m, n = a.n;
    const double* A = a.A + (size_t)lp * m * n;
    double* dst = a.At + ((size_t)lp * nstrips + s) * mr * KF_W;
    for (int r = threadIdx.x; r < mr; r += blockDim.x) {
        double2 v[KF_W / 2];
#pragma unroll
        for (int q = 0; q < KF_W / 2; ++q) {
            const int col = s * KF_W + 2 * q;
            v[q] = (r < m && col < n) ? *reinterpret_cast<const double2*>(A + (size_t)r * n + col) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int q = 0; q < KF_W / 2; ++q) reinterpret_cast<double2*>(dst + (size_t)r * KF_W)[q] = v[q];
    }
}

#endif

// 3-D tensor map over the caller's A[B][m][n] (row-major, f64): box = KF_W columns x mr rows x 1 LP.
inline int kf_make_strip_tmap(CUtensorMap* out, const double* A, int B, int m, int n, int mr) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        IPM_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            g_last_error = "cuTensorMapEncodeTiled is not available from this driver";
            return IPM_ERR_CUDA;
        }
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)m, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)n * 8, (cuuint64_t)m * n * 8};       // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)KF_W, (cuuint32_t)mr, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_error = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
        return IPM_ERR_CUDA;
    }
    return IPM_OK;
}

}  // namespace ipm
