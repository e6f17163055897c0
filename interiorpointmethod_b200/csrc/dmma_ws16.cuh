// Variant of the warp-specialised NT product (dmma_ws.cuh) with SIXTEEN consumer warps of 16 x 64 accumulator blocks
// instead of eight of 32 x 64: four MMA-issuing warps per scheduler instead of two.  Same ring, same producers, same
// operations on every accumulator in the same order (bitwise the same C); only who holds which sub-tile changes.
//
// Why: ncu on the 8-consumer kernel shows the consumers waiting for the FP64 pipe (math-pipe throttle) while the pipe
// is 83 % busy - with two consumers per scheduler their non-MMA stretches (fragment loads, scaling, selects, stage
// hand-over) coincide often enough to leave it idle.  With 20 warps the launch is sized for 96 registers per thread;
// the producer warpgroup gives back down to 32, the four consumer warpgroups take 112: 64 accumulator registers, ONE
// fragment set (the other warps of the scheduler cover its latency), addresses.
//
// Diagonal tiles (lower triangle, 136 sub-tiles): warp W < 8 takes the nine static sub-tiles of the strip pair
// (W, 15-W) of the 8-consumer kernel (sub-tiles 0..7 of strip 15-W, sub-tile 0 of strip W) and forms the right-hand
// side; warp W + 8 takes its eight run-time slots.
#pragma once
#include "dmma_ws.cuh"

namespace ipm {

constexpr int WS16_CONSUMER_WARPS = 16;
constexpr int WS16_THREADS = (WS16_CONSUMER_WARPS + WS_PRODUCER_WARPS) * 32;     // 640
constexpr int WS16_REGS_CONSUMER = 112, WS16_REGS_PRODUCER = 32;

#ifdef __CUDACC__
template <int EPI, bool SCALE, bool RHS>
__device__ __forceinline__ void ws16_diag_tile(const DmmaArgs& a, const double* Ps, const double* Qs, const double* Ds,
                                               const double* Vs, uint64_t* full, uint64_t* empty, uint32_t& it, int nk, int z,
                                               int row0, int lane, int warp) {
    constexpr int BK = 16, LD = WsGeom<BK>::LD, S = WsGeom<BK>::STAGES;
    const bool fixed = warp < 8;              // role: nine static sub-tiles / eight slots
    const int W = warp & 7;
    const int R0 = 8 * W, R1 = 8 * (15 - W);
    const int g = lane >> 2, t = lane & 3;
    // accumulators: acc[0..7] = strip 15-W sub-tiles 0..7 (fixed) or the slots (variable); acc8 = strip W sub-tile 0 (fixed)
    double acc[8][2], acc8[2];
    int sub[8];
    bool lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lo[i] = !fixed && (i + 1 <= W);
        sub[i] = fixed ? i : (lo[i] ? i + 1 : 8 + i - W);
    }
    double* C = a.C + (size_t)z * a.strideC;
    auto init = [&](int r, int c, double& v0, double& v1) {
        v0 = v1 = 0.0;
        if (EPI == 1 && r < a.rowsP && c < a.rowsQ) {
            const double* cp = C + (size_t)r * a.ldc + c;
            if (c + 1 < a.rowsQ) { const double2 v = *reinterpret_cast<const double2*>(cp); v0 = v.x; v1 = v.y; }
            else v0 = cp[0];
        }
    };
#pragma unroll
    for (int i = 0; i < 8; ++i) init(row0 + (lo[i] ? R0 : R1) + g, row0 + 8 * sub[i] + 2 * t, acc[i][0], acc[i][1]);
    acc8[0] = acc8[1] = 0.0;
    if (fixed) init(row0 + R0 + g, row0 + 2 * t, acc8[0], acc8[1]);
    double dot0 = 0.0, dot1 = 0.0;
    const int pa0 = (R0 + g) * LD + t, pa1 = (R1 + g) * LD + t, pb = g * LD + t;
    for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % S;
        mbar_wait(full + s, (it / S) & 1);
        const double* ps = Ps + s * WS_BM * LD;
        const double* qs = Qs + s * WS_BN * LD + pb;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double x0 = ps[pa0 + kk], x1 = ps[pa1 + kk], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) b[i] = qs[sub[i] * 8 * LD + kk];
            if (SCALE) {
                const double dv = Ds[s * BK + t + kk];
                const double dk = (EPI == 1) ? -dv : dv;
                x0 *= dk;
                x1 *= dk;
            } else if (EPI == 1) {
                x0 = -x0;
                x1 = -x1;
            }
            if (RHS && fixed) {
                const double vv = Vs[s * BK + t + kk];
                dot0 = fma(x0, vv, dot0);
                dot1 = fma(x1, vv, dot1);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[i][0], acc[i][1], lo[i] ? x0 : x1, b[i]);
            if (fixed) dmma884(acc8[0], acc8[1], x0, b[0]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    auto store = [&](int r, int c, double v0, double v1) {
        if (r >= a.rowsP || c >= a.rowsQ) return;
        double* cp = C + (size_t)r * a.ldc + c;
        if (c + 1 < a.rowsQ) *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
        else cp[0] = v0;
    };
#pragma unroll
    for (int i = 0; i < 8; ++i) store(row0 + (lo[i] ? R0 : R1) + g, row0 + 8 * sub[i] + 2 * t, acc[i][0], acc[i][1]);
    if (fixed) store(row0 + R0 + g, row0 + 2 * t, acc8[0], acc8[1]);
    if (RHS && fixed) {
        dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1);
        dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1);
        dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
        dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
        if (t == 0) {
            const double* rb = a.rbvec + (size_t)z * a.strideR;
            double* out = a.rhs + (size_t)z * a.strideR;
            const int r0 = row0 + R0 + g, r1 = row0 + R1 + g;
            if (r0 < a.rowsP) out[r0] = -rb[r0] - dot0;
            if (r1 < a.rowsP) out[r1] = -rb[r1] - dot1;
        }
    }
}

template <int EPI, bool SCALE, bool RHS = false>
__global__ void __launch_bounds__(WS16_THREADS, 1) dmma_ws16_kernel(const DmmaArgs a, int ntri, int total_tiles) {
    constexpr int BK = 16, LD = WsGeom<BK>::LD, S = WsGeom<BK>::STAGES;
    extern __shared__ __align__(128) unsigned char ws_raw[];
    double* Ps = reinterpret_cast<double*>(ws_raw);            // [S][128][LD]
    double* Qs = Ps + S * WS_BM * LD;                          // [S][128][LD]
    double* Ds = Qs + S * WS_BN * LD;                          // [S][BK]
    double* Vs = Ds + S * BK;                                  // [S][BK]  (RHS only)
    uint64_t* full = reinterpret_cast<uint64_t*>(Vs + S * BK);
    uint64_t* empty = full + S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 32 * WS_PRODUCER_WARPS);
            mbar_init(empty + s, WS16_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nk = (a.K + BK - 1) / BK;
    if (warp >= WS16_CONSUMER_WARPS) {
        setmaxnreg_dec<WS16_REGS_PRODUCER>();
        ws_produce<SCALE, BK, RHS, WS_PRODUCER_WARPS>(a, Ps, Qs, Ds, Vs, full, empty, ntri, total_tiles,
                                                      warp - WS16_CONSUMER_WARPS, lane);
        return;
    }
    setmaxnreg_inc<WS16_REGS_CONSUMER>();
    // ---------------------------------------------------------------------- consumers: 8 x 2 warps, 16 x 64 each
    constexpr int MI = 2, NI = 8;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 1) * 16, wn0 = (warp & 1) * 64;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / ntri;
        if (a.active && a.active[z] == 0) continue;
        int bi, bj;
        if (a.col0_only) { bi = tile - z * ntri; bj = 0; }
        else tri_decode(tile - z * ntri, bi, bj);
        if (bi == bj) {
            ws16_diag_tile<EPI, SCALE, RHS>(a, Ps, Qs, Ds, Vs, full, empty, it, nk, z, bi * WS_BM, lane, warp);
            continue;
        }
        double acc[MI][NI][2];
        double* C = a.C + (size_t)z * a.strideC;
        const int row0 = bi * WS_BM, col0 = bj * WS_BN;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int r = row0 + wm0 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int c = col0 + wn0 + j * 8 + 2 * t;
                double v0 = 0.0, v1 = 0.0;
                if (EPI == 1 && r < a.rowsP && c < a.rowsQ) {
                    const double* cp = C + (size_t)r * a.ldc + c;
                    if (c + 1 < a.rowsQ) { const double2 v = *reinterpret_cast<const double2*>(cp); v0 = v.x; v1 = v.y; }
                    else v0 = cp[0];
                }
                acc[i][j][0] = v0;
                acc[i][j][1] = v1;
            }
        }
        for (int kt = 0; kt < nk; ++kt, ++it) {
            const int s = it % S;
            mbar_wait(full + s, (it / S) & 1);
            const double* ps = Ps + s * WS_BM * LD + (wm0 + g) * LD + t;
            const double* qs = Qs + s * WS_BN * LD + (wn0 + g) * LD + t;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double af[MI], bf[NI];
#pragma unroll
                for (int i = 0; i < MI; ++i) af[i] = ps[i * 8 * LD + kk];
#pragma unroll
                for (int j = 0; j < NI; ++j) bf[j] = qs[j * 8 * LD + kk];
                if (SCALE) {
                    const double dv = Ds[s * BK + t + kk];
                    const double dk = (EPI == 1) ? -dv : dv;
#pragma unroll
                    for (int i = 0; i < MI; ++i) af[i] *= dk;
                } else if (EPI == 1) {
#pragma unroll
                    for (int i = 0; i < MI; ++i) af[i] = -af[i];
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int r = row0 + wm0 + i * 8 + g;
            if (r >= a.rowsP) continue;
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int c = col0 + wn0 + j * 8 + 2 * t;
                if (c >= a.rowsQ) continue;
                double* cp = C + (size_t)r * a.ldc + c;
                if (c + 1 < a.rowsQ) *reinterpret_cast<double2*>(cp) = make_double2(acc[i][j][0], acc[i][j][1]);
                else cp[0] = acc[i][j][0];
            }
        }
    }
}

#endif
inline std::atomic<int>& ws_consumer_warps() {       // 8 (default) or 16: ipm_set_syrk_consumers, A/B measurements
    static std::atomic<int> n{8};
    return n;
}
#ifdef __CUDACC__

template <int EPI, bool SCALE, bool RHS>
inline int dmma_ws16_launch(const DmmaArgs& a, int batch, cudaStream_t st) {
    auto kern = dmma_ws16_kernel<EPI, SCALE, RHS>;
    constexpr size_t smem = WsGeom<16>::smem;
    IPM_TRY(ensure_dyn_smem(kern, smem));
    if (a.rowsP <= 0 || batch <= 0 || a.K <= 0) return IPM_OK;
    const int T = ceil_div(a.rowsP, WS_BM);
    const int ntri = a.col0_only ? T : T * (T + 1) / 2;
    const int64_t total = (int64_t)ntri * batch;
    if (total > 0x7fffffff) return IPM_ERR_SHAPE;
    const int cap = (a.max_ctas > 0 && a.max_ctas < kNumSMs) ? a.max_ctas : kNumSMs;
    const int grid = (int)std::min<int64_t>(total, cap);
    kern<<<grid, WS16_THREADS, smem, st>>>(a, ntri, (int)total);
    count_launch();
    return launch_check();
}
#endif

}  // namespace ipm
