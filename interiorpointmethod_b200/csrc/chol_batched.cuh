// Fused per-LP Cholesky for the batched workload: ONE CTA factors one m x m matrix (m <= KBC_MAX_M) end to
// end in 32-wide panels, so the factorisation of all LPs is a single launch and M makes one trip through HBM
// (read) and one back (L).  Replaces solve_linear (main.py:176-182) on M (main.py:223-224).
//
// Left-looking with look-ahead inside the CTA.  Warp 0 is the FACTOR warp, warps 1-7 are UPDATE warps.
// While warp 0 factors the 32x32 diagonal block of panel J (lane i keeps row i in registers, the pivot travels by
// warp shuffle, safeguard p <= tau*maxdiag or NaN -> 1e128), the update warps already build
// panel J+1:  acc = M[:, J+1] - sum_{K<J} L[:, K] L[J+1, K]^T  on the tensor pipe (DMMA.8x8x4; A fragments
// straight from global memory = L written by this CTA a moment ago, B fragments from a staged shared chunk).
// After the rows below the block are solved (one thread per row) and panel J is stored, the missing K = J term
// is added from SHARED memory (the freshly solved panel is both operands) and the accumulators become the raw
// panel J+1.  Per panel: diag || early update, trsm, store + late update - four block barriers.
#pragma once
#include "common.cuh"

namespace ipm {

constexpr int KBC_NT = 256;         // batched variant: 2 CTAs per SM
constexpr int KBC_NT_BIG = 512;     // single-matrix variant for 256 < m <= 512: 1 CTA per SM
// NT/32 - 1 update warps x 4 accumulator slots x 8 rows, plus the 32 rows of the diagonal block
constexpr int kbc_max_m(int nt) { return 32 + (nt / 32 - 1) * 4 * 8; }
constexpr int KBC_MAX_M = kbc_max_m(KBC_NT);          // 256
constexpr int KBC_MAX_M_BIG = kbc_max_m(KBC_NT_BIG);  // 512
constexpr int KBC_LD = 34;     // 2-way conflicts for both access patterns (MMA fragments (g,t) and one row per thread); 33 made the
                                // fragment reads 4-way (ncu: 53 M shared bank conflicts per 2048 LPs), 36 the row reads
constexpr int KBC_LDT = 34;     // transposed diagonal block: even so that 128-bit reads stay aligned
constexpr int KBC_KC = 64;      // columns of L[J+1, :] staged in shared memory per chunk of the update
constexpr int KBC_LDB = KBC_KC + 4;   // 68 = 4 mod 16: conflict-free 64-bit fragment reads

struct CholBatchedArgs {
    double* M; int64_t ldm; int64_t strideM;
    double* scal; int64_t strideScal;     // S_MAXDIAG, S_NFIXED written per LP (nullable)
    double tau;
    int m;
    const int* active;
    unsigned char* dep = nullptr;         // single matrix only: dependent-row mask (see CholArgs in chol.cuh)
    int dep_mode = 0;
    // Diagonal block of a LARGER matrix (potrf_blocked, chol.cuh): the safeguard threshold comes from the whole
    // matrix (scal[S_MAXDIAG], set by k_maxdiag), the count of replaced pivots is accumulated, S_MAXDIAG is left alone.
    int as_block = 0;
};

inline size_t kbc_smem_bytes(int m) {
    const int below = m > 32 ? m - 32 : 0;
    return (size_t)(32 * KBC_LD + 2 + 32 * KBC_LDT + 32 * KBC_LDB + (size_t)below * KBC_LD + 32) * sizeof(double);
}

#ifdef __CUDACC__
#ifdef KBC_PROFILE
__device__ long long kbc_prof[32];
#define KBC_T(slot) do { if (lp == 0 && lane == 0 && (warp == 1 || warp == 0)) { long long _t = clock64(); kbc_prof[(warp == 0 ? 16 : 0) + (slot)] += _t - t_last; t_last = _t; } } while (0)
#else
#define KBC_T(slot) do {} while (0)
#endif
template <int NT>
__device__ __forceinline__ void kbc_update_bar() {      // barrier among the update warps only
    asm volatile("bar.sync 1, %0;" ::"n"(NT - 32) : "memory");
}

// One chunk of the early update for the NTI row tiles of a warp.  A fragments come straight from global memory
// as 16-byte loads: lane (g,t) reads A[g][k+2t], A[g][k+2t+1] and feeds them to two consecutive MMAs, i.e. the
// contraction index is permuted inside every 8-column group - harmless because the B fragments (from the staged
// shared chunk) are read with the same permutation.  One group (= 2 MMA steps) is prefetched ahead.
template <int NTI>
__device__ __forceinline__ void kbc_update_chunk(double (&acc)[4][4][2], const double* __restrict__ Mb,
                                                 const int (&aoff2)[4], const bool (&aok)[4], const double* Bs,
                                                 int k0, int kc, int g, int t) {
    double2 af[NTI], afn[NTI];
#pragma unroll
    for (int i = 0; i < NTI; ++i)
        af[i] = aok[i] ? *reinterpret_cast<const double2*>(Mb + aoff2[i] + k0) : make_double2(0.0, 0.0);
    const double* bs = Bs + g * KBC_LDB + 2 * t;
#pragma unroll 2
    for (int kk = 0; kk < kc; kk += 8) {
#pragma unroll
        for (int i = 0; i < NTI; ++i)
            afn[i] = (aok[i] && kk + 8 < kc) ? *reinterpret_cast<const double2*>(Mb + aoff2[i] + k0 + kk + 8)
                                             : make_double2(0.0, 0.0);
        double2 bf[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            // acc holds +M - sum L L^T: the sign lives in the B fragments (shared memory, short latency), so that the
            // accumulators can be initialised by plain loads that nothing touches until the first MMA
            const double2 b2 = *reinterpret_cast<const double2*>(bs + ni * 8 * KBC_LDB + kk);
            bf[ni] = make_double2(-b2.x, -b2.y);
        }
#pragma unroll
        for (int ti = 0; ti < NTI; ++ti)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[ti][ni][0], acc[ti][ni][1], af[ti].x, bf[ni].x);
#pragma unroll
        for (int ti = 0; ti < NTI; ++ti)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[ti][ni][0], acc[ti][ni][1], af[ti].y, bf[ni].y);
#pragma unroll
        for (int i = 0; i < NTI; ++i) af[i] = afn[i];
    }
}

// K = J term of panel J+1 from shared memory: rows of the new panel are Ps[0..nrows), its own first 32 rows
// (= rows of the next diagonal block) are the B operand.
template <int NTI, int KBC_UW>
__device__ __forceinline__ void kbc_update_late(double (&acc)[4][4][2], const double* Ps, int uw, int nrows, int g,
                                                int t) {
#pragma unroll
    for (int kk = 0; kk < 32; kk += 4) {
        double af[NTI], bf[4];
#pragma unroll
        for (int i = 0; i < NTI; ++i) {
            const int r = (uw + i * KBC_UW) * 8 + g;
            af[i] = (r < nrows) ? Ps[(size_t)r * KBC_LD + kk + t] : 0.0;
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int r = ni * 8 + g;
            bf[ni] = (r < nrows) ? -Ps[(size_t)r * KBC_LD + kk + t] : 0.0;
        }
#pragma unroll
        for (int ti = 0; ti < NTI; ++ti)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[ti][ni][0], acc[ti][ni][1], af[ti], bf[ni]);
    }
}

template <int KBC_NT>
static __device__ __forceinline__ void d_kb_chol(const CholBatchedArgs& a) {
    constexpr int KBC_UW = KBC_NT / 32 - 1;   // update warps
    extern __shared__ __align__(16) double smem[];
    double* D = smem;                       // [32][33]  diagonal block, becomes L_JJ
    double* DT = smem + 32 * KBC_LD + 2;    // [32][34]  DT[k][j] = L_JJ[j][k]  (+2 doubles: 16-byte aligned)
    double* Bs = DT + 32 * KBC_LDT;         // [32][68]  chunk of the (J+1)-block rows of L (B operand, early update)
    double* Ps = Bs + 32 * KBC_LDB;         // [m-32][33] rows below the diagonal block
    __shared__ double sh[32];
    __shared__ double dg[32];               // 1 / diagonal of L_JJ
    __shared__ __align__(16) double colb[64];   // column j of L_JJ while it is being folded into the rows (two buffers)
    __shared__ double s_maxdiag;
    __shared__ int s_nfix;
    const int lp = blockIdx.x;
    if (a.active && a.active[lp] == 0) return;
    const int m = a.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int uw = warp - 1;                // index among the update warps (-1 for the factor warp)
    static_assert(KBC_NT == 256 || KBC_NT == 512, "two variants");
    double* Mb = a.M + (size_t)lp * a.strideM;
    const int64_t ldm = a.ldm;

    // The lower triangle of M_i goes to L2 now (one bulk prefetch per row): M was written by the SYRK of ALL LPs and
    // has left L2 long ago, so without this every first touch of a panel column (accumulator initialisation of the
    // early update) is an exposed DRAM round trip in the middle of the factorisation.
    for (int r = tid; r < m; r += KBC_NT) {
        const uint32_t bytes = (uint32_t)(((r + 1) * 8 + 15) & ~15);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Mb + (size_t)r * ldm), "r"(bytes) : "memory");
    }
    // max_i M_ii (threshold of the safeguard) and the raw first panel: rows 0..m-1, columns 0..31
    {
        double v = red_identity<RED_MAX>();
        for (int i = tid; i < m; i += KBC_NT) v = fmax(v, Mb[(size_t)i * ldm + i]);
        if (!(ldm & 1) && !(reinterpret_cast<uintptr_t>(Mb) & 15)) {        // aligned rows: 16 bytes per copy
#pragma unroll 8
            for (int idx = tid; idx < m * 16; idx += KBC_NT) {
                const int r = idx >> 4, c = (idx & 15) * 2;
                const double* src = Mb + (size_t)r * ldm + c;
                double2 val = make_double2(0.0, 0.0);
                if (c + 1 < m) val = *reinterpret_cast<const double2*>(src);
                else if (c < m) val.x = src[0];
                double* dst = (r < 32) ? D + r * KBC_LD + c : Ps + (size_t)(r - 32) * KBC_LD + c;
                *reinterpret_cast<double2*>(dst) = val;
            }
        } else {
#pragma unroll 8
            for (int idx = tid; idx < m * 32; idx += KBC_NT) {
                const int r = idx >> 5, c = idx & 31;
                const double val = (c < m) ? Mb[(size_t)r * ldm + c] : 0.0;
                if (r < 32) D[r * KBC_LD + c] = val;
                else Ps[(size_t)(r - 32) * KBC_LD + c] = val;
            }
        }
        v = block_red<RED_MAX>(v, sh);
        if (tid == 0) { s_maxdiag = v; s_nfix = 0; }
        __syncthreads();
    }
    const double thresh = a.tau * (a.as_block ? a.scal[(size_t)lp * a.strideScal + S_MAXDIAG] : s_maxdiag);
#ifdef KBC_PROFILE
    long long t_last = clock64();
#endif

    // ---- phases executed by every thread (called from both role loops below)
    auto phase_trsm = [&](int nrows1) {
        // rows below the block: x L_JJ^T = a, one thread per row, 8 columns at a time
        for (int r = tid; r < nrows1; r += KBC_NT) {
            double* pr = Ps + (size_t)r * KBC_LD;
#pragma unroll 1
            for (int jb = 0; jb < 32; jb += 8) {
                double x8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x8[q] = pr[jb + q];
#pragma unroll 4
                for (int k = 0; k < jb; ++k) {
                    const double xk = pr[k];
                    const double2* lpp = reinterpret_cast<const double2*>(DT + k * KBC_LDT + jb);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 lv = lpp[q];
                        x8[2 * q] -= xk * lv.x;
                        x8[2 * q + 1] -= xk * lv.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double* lrow = DT + (jb + q) * KBC_LDT + jb;
                    const double xv = x8[q] * dg[jb + q];
                    x8[q] = xv;
#pragma unroll
                    for (int q2 = q + 1; q2 < 8; ++q2) x8[q2] -= xv * lrow[q2];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) pr[jb + q] = x8[q];
            }
        }
    };
    auto phase_store = [&](int j0, int rows, int nb) {
        for (int idx = tid; idx < rows * 32; idx += KBC_NT) {
            const int pr = idx >> 5, c = idx & 31;
            if (c >= nb) continue;
            if (pr < 32) {
                if (c <= pr) Mb[(size_t)(j0 + pr) * ldm + j0 + c] = D[pr * KBC_LD + c];
            } else {
                Mb[(size_t)(j0 + pr) * ldm + j0 + c] = Ps[(size_t)(pr - 32) * KBC_LD + c];
            }
        }
    };

    // Panel J goes back to global memory from the factor warp alone (128-bit copies, two rows per instruction),
    // while the update warps apply the K = J term: the store is off the chain of the panel step (it was 15 % of it
    // with all threads storing before the K = J term).  Needs 16-byte aligned rows (guaranteed by the entry points;
    // a diagonal block of a larger matrix at an odd offset takes the all-threads path).
    const bool fw_store = !(ldm & 1) && !(reinterpret_cast<uintptr_t>(Mb) & 15);
    auto store_by_factor_warp = [&](int j0, int rows, int nb) {
        const int nrows1 = rows - 32;                                // rows below the block exist only when nb == 32
        const int half = lane >> 4, c2 = 2 * (lane & 15);
        double* dstp = Mb + (size_t)(j0 + 32 + half) * ldm + j0 + c2;
        const double* srcp = Ps + (size_t)half * KBC_LD + c2;
#pragma unroll 8
        for (int r = half; r < nrows1; r += 2) {
            *reinterpret_cast<double2*>(dstp) = *reinterpret_cast<const double2*>(srcp);
            dstp += 2 * ldm;
            srcp += 2 * KBC_LD;
        }
        const int nd = rows < 32 ? rows : 32;
        for (int pr = 0; pr < nd; ++pr)
            if (lane <= pr && lane < nb) Mb[(size_t)(j0 + pr) * ldm + j0 + lane] = D[pr * KBC_LD + lane];
    };

    if (warp == 0) {
        // =================================================================== FACTOR warp
        for (int j0 = 0; j0 < m; j0 += 32) {
            const int nb = (m - j0 < 32) ? (m - j0) : 32;
            const int rows = m - j0, nrows1 = rows - 32;
            {
                // diagonal block of panel J: register-resident, right-looking, warp shuffle pivot
                double arow[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) arow[c] = D[lane * KBC_LD + c];
                // The lane's own diagonal entry lives in a register of its own: its update needs no value of another
                // lane (l_ij of this lane, squared), so the next pivot is ready one multiply-add after l_ij - the trip
                // of column j through shared memory (colb) is off the pivot-to-pivot chain.  Same operations on the
                // same values as the update through colb: bitwise the same block.
                double mydiag = D[lane * KBC_LD + lane];
                int nfix = 0;
                double my_inv = 1.0;
                const bool masked = a.dep != nullptr && a.dep_mode == 1;
                const unsigned forced_bits = masked ? __ballot_sync(0xffffffffu, lane < nb && a.dep[j0 + lane] != 0) : 0u;
                bool my_bad = false;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < nb) {
                        double p = __shfl_sync(0xffffffffu, mydiag, j);
                        const bool bad = !(p > thresh) || ((forced_bits >> j) & 1u);
                        if (bad) p = kPivotBig;
                        const double inv = rsqrt(p);
                        const double lij = (lane == j) ? p * inv : arow[j] * inv;
                        mydiag = fma(-lij, lij, mydiag);                 // matters for lanes > j only
                        if (lane == j) { my_inv = inv; nfix += bad ? 1 : 0; my_bad = bad; }
                        arow[j] = lij;
                        double* cb = colb + (j & 1) * 32;                // two buffers: one warp barrier per column
                        cb[lane] = lij;
                        __syncwarp();
                        const double2* cb2 = reinterpret_cast<const double2*>(cb);
#pragma unroll
                        for (int k2 = (j + 1) >> 1; k2 < 16; ++k2) {     // k > lane unused
                            const double2 v = cb2[k2];
                            if (2 * k2 > j) arow[2 * k2] = fma(-lij, v.x, arow[2 * k2]);
                            arow[2 * k2 + 1] = fma(-lij, v.y, arow[2 * k2 + 1]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c <= lane) D[lane * KBC_LD + c] = arow[c];
                // the transposed copy the substitution reads (DT[k][j] = L_JJ[j][k]) straight from the registers: one
                // block barrier and a pass over the block by all threads less per panel
#pragma unroll
                for (int c = 0; c < 32; ++c) DT[c * KBC_LDT + lane] = (lane < nb && c <= lane) ? arow[c] : 0.0;
                dg[lane] = my_inv;
                if (a.dep != nullptr && a.dep_mode == 2 && lane < nb) a.dep[j0 + lane] = my_bad ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) nfix += __shfl_xor_sync(0xffffffffu, nfix, o);
                if (lane == 0) s_nfix += nfix;               // lane j counted its own pivot
            }
            KBC_T(0);
            __syncthreads();                                       // (S1)
            KBC_T(2);
            phase_trsm(nrows1);
            __syncthreads();                                       // (S3)
            KBC_T(3);
            if (fw_store) store_by_factor_warp(j0, rows, nb);
            else phase_store(j0, rows, nb);
            __syncthreads();                                       // (S4)
            __syncthreads();                                       // (S5)
            KBC_T(4);
        }
    } else {
        // =================================================================== UPDATE warps
        double acc[4][4][2];
        // acc <- +M[tile rows][j1 .. j1+31], the start of the early update of the panel at column j1.  Issued one
        // panel step ahead (before the loop, then right after the accumulators of the previous panel have gone to
        // shared memory), so that the loads are in flight across the barrier and the staging of the first chunk
        // instead of stalling the first MMA (ncu source view, round 2: 13.6 % of the kernel's stall samples).
        auto init_acc = [&](int j1) {
            const int nrows = m - j1;
            if (nrows <= 0) return;
            const int ntile = (nrows + 7) >> 3;
#pragma unroll
            for (int ti = 0; ti < 4; ++ti) {
                const int tile = uw + ti * KBC_UW;
                const int r = j1 + tile * 8 + g;
                const bool ok = (tile < ntile) && (r < m);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int c = j1 + ni * 8 + 2 * t;
                    double v0 = 0.0, v1 = 0.0;
                    if (ok) {                        // raw loads: no arithmetic on them before the first MMA
                        const double* src = Mb + (size_t)r * ldm + c;
                        if (fw_store && c + 1 < m) {
                            const double2 v = *reinterpret_cast<const double2*>(src);
                            v0 = v.x;
                            v1 = v.y;
                        } else {
                            if (c < m) v0 = src[0];
                            if (c + 1 < m) v1 = src[1];
                        }
                    }
                    acc[ti][ni][0] = v0;
                    acc[ti][ni][1] = v1;
                }
            }
        };
        init_acc(32);
        for (int j0 = 0; j0 < m; j0 += 32) {
            const int nb = (m - j0 < 32) ? (m - j0) : 32;
            const int rows = m - j0;                       // panel rows (diag block included)
            const int nrows1 = rows - 32;                  // rows of the NEXT panel (<= 0: this is the last one)
            const int ntile1 = nrows1 > 0 ? (nrows1 + 7) >> 3 : 0;
            const int nti = (ntile1 > uw) ? (ntile1 - uw + KBC_UW - 1) / KBC_UW : 0;
            if (nrows1 > 0) {
                // early update of panel J+1 (columns j1..j1+31, rows j1..m-1) with K < J
                const int j1 = j0 + 32;
                bool aok[4];
                int aoff[4];                     // 32-bit element offsets (m*ldm <= 2^16 here)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int tile = uw + i * KBC_UW;
                    const int ra = j1 + tile * 8 + g;
                    aok[i] = (tile < ntile1) && (ra < m);
                    aoff[i] = (aok[i] ? ra : j1) * (int)ldm + 2 * t;     // 16-byte aligned: ldm and k are even
                }
                for (int k0 = 0; k0 < j0; k0 += KBC_KC) {
                    const int kc = (j0 - k0 < KBC_KC) ? (j0 - k0) : KBC_KC;
                    kbc_update_bar<KBC_NT>();                                  // previous chunk fully consumed
                    const int kshift = (kc == KBC_KC) ? 6 : 5;         // kc is 64 or 32 (j0 is a multiple of 32)
                    if (fw_store) {                                    // aligned rows: 16 bytes per copy
#pragma unroll 5
                        for (int idx = tid - 32; idx < 16 * kc; idx += KBC_UW * 32) {
                            const int r = idx >> (kshift - 1), k = (idx & ((kc >> 1) - 1)) * 2;
                            const double2 v = (j1 + r < m)
                                                  ? *reinterpret_cast<const double2*>(Mb + (size_t)(j1 + r) * ldm + k0 + k)
                                                  : make_double2(0.0, 0.0);
                            *reinterpret_cast<double2*>(Bs + r * KBC_LDB + k) = v;
                        }
                    } else {
#pragma unroll 5
                        for (int idx = tid - 32; idx < 32 * kc; idx += KBC_UW * 32) {
                            const int r = idx >> kshift, k = idx & (kc - 1);
                            Bs[r * KBC_LDB + k] = (j1 + r < m) ? Mb[(size_t)(j1 + r) * ldm + k0 + k] : 0.0;
                        }
                    }
                    kbc_update_bar<KBC_NT>();
                    switch (nti) {
                        case 1: kbc_update_chunk<1>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                        case 2: kbc_update_chunk<2>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                        case 3: kbc_update_chunk<3>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                        case 4: kbc_update_chunk<4>(acc, Mb, aoff, aok, Bs, k0, kc, g, t); break;
                        default: break;
                    }
                }
            }
            KBC_T(0);
            __syncthreads();                                       // (S1) block factored (D, DT, dg), early update done
            KBC_T(2);
            phase_trsm(nrows1);
            __syncthreads();                                       // (S3) panel J final in shared memory
            KBC_T(3);
            if (!fw_store) phase_store(j0, rows, nb);
            KBC_T(4);
            if (nrows1 > 0) {
                // K = J term of panel J+1 straight from shared memory
                switch (nti) {
                    case 1: kbc_update_late<1, KBC_UW>(acc, Ps, uw, nrows1, g, t); break;
                    case 2: kbc_update_late<2, KBC_UW>(acc, Ps, uw, nrows1, g, t); break;
                    case 3: kbc_update_late<3, KBC_UW>(acc, Ps, uw, nrows1, g, t); break;
                    case 4: kbc_update_late<4, KBC_UW>(acc, Ps, uw, nrows1, g, t); break;
                    default: break;
                }
            }
            KBC_T(5);
            __syncthreads();                                       // (S4) everyone is done reading D / Ps
            KBC_T(6);
            if (nrows1 > 0) {
                // accumulators become the raw panel J+1
#pragma unroll
                for (int ti = 0; ti < 4; ++ti) {
                    const int tile = uw + ti * KBC_UW;
                    const int pr = tile * 8 + g;                // row inside the new panel
                    if (tile >= ntile1 || pr >= nrows1) continue;
                    double* dst = (pr < 32) ? (D + pr * KBC_LD) : (Ps + (size_t)(pr - 32) * KBC_LD);
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)           // (KBC_LD is even and the blocks start on 16-byte boundaries)
                        *reinterpret_cast<double2*>(dst + ni * 8 + 2 * t) = make_double2(acc[ti][ni][0], acc[ti][ni][1]);
                }
                init_acc(j0 + 64);                                 // for the next panel step, ahead of its barrier
            }
            __syncthreads();                                       // (S5)
            KBC_T(7);
        }
    }
    __syncthreads();
    if (tid == 0 && a.scal) {
        if (a.as_block) {
            if (s_nfix) a.scal[(size_t)lp * a.strideScal + S_NFIXED] += (double)s_nfix;
        } else {
            a.scal[(size_t)lp * a.strideScal + S_MAXDIAG] = s_maxdiag;
            a.scal[(size_t)lp * a.strideScal + S_NFIXED] = (double)s_nfix;
        }
    }
}

template <int KBC_NT>
static __global__ void __launch_bounds__(KBC_NT, KBC_NT == 256 ? 2 : 1) kb_chol(const CholBatchedArgs a) {
    d_kb_chol<KBC_NT>(a);
}

// The diagonal block M[j0 : j0+nb, j0 : j0+nb] (nb <= 256) of one matrix of order m, factored by the fused kernel: the
// panel kernel of the blocked factorisation (chol.cuh) for one large matrix.
inline int potrf_diag_block_fused(double* M, int64_t ldm, int j0, int nb, double* scal, double tau, cudaStream_t st,
                                  unsigned char* dep, int dep_mode) {
    IPM_TRY(ensure_dyn_smem(kb_chol<KBC_NT>, kbc_smem_bytes(kbc_max_m(KBC_NT))));
    CholBatchedArgs a;
    a.M = M + (size_t)j0 * ldm + j0; a.ldm = ldm; a.strideM = 0; a.scal = scal; a.strideScal = 0; a.tau = tau; a.m = nb;
    a.active = nullptr;
    a.dep = dep ? dep + j0 : nullptr; a.dep_mode = dep ? dep_mode : 0;
    a.as_block = 1;
    kb_chol<KBC_NT><<<1, KBC_NT, kbc_smem_bytes(nb), st>>>(a);
    count_launch();
    return launch_check();
}

template <int NT>
inline int potrf_batched_fused_nt(double* M, int64_t ldm, int64_t strideM, int m, int batch, double* scal,
                                  int64_t strideScal, double tau, const int* active, cudaStream_t st,
                                  unsigned char* dep = nullptr, int dep_mode = 0) {
    IPM_TRY(ensure_dyn_smem(kb_chol<NT>, kbc_smem_bytes(kbc_max_m(NT))));
    CholBatchedArgs a;
    a.M = M; a.ldm = ldm; a.strideM = strideM; a.scal = scal; a.strideScal = strideScal; a.tau = tau; a.m = m;
    a.active = active;
    a.dep = (batch == 1) ? dep : nullptr; a.dep_mode = (batch == 1) ? dep_mode : 0;
    kb_chol<NT><<<batch, NT, kbc_smem_bytes(m), st>>>(a);
    count_launch();
    return launch_check();
}
// m <= 256: 256-thread CTAs (two per SM); m <= 512: 512-thread CTAs
inline int potrf_batched_fused(double* M, int64_t ldm, int64_t strideM, int m, int batch, double* scal,
                               int64_t strideScal, double tau, const int* active, cudaStream_t st,
                               unsigned char* dep = nullptr, int dep_mode = 0) {
    if (m <= KBC_MAX_M)
        return potrf_batched_fused_nt<KBC_NT>(M, ldm, strideM, m, batch, scal, strideScal, tau, active, st, dep, dep_mode);
    return potrf_batched_fused_nt<KBC_NT_BIG>(M, ldm, strideM, m, batch, scal, strideScal, tau, active, st, dep, dep_mode);
}
#endif

}  // namespace ipm
