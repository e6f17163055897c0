// Warp-specialised, persistent FP64 tensor-core "NT" product (second generation of dmma_gemm.cuh).
//
//   C[i][j] (op)= sum_k P[i][k] * d[k] * Q[j][k]          lower 128x128 tiles only, batch in the tile index
//
// One CTA per SM loops over output tiles.  Warps 8 and 9 are PRODUCERS: they stream the 128x16 slabs of the two
// operands (and the matching 16 entries of d) from global memory into a 4-stage shared-memory ring with cp.async (LDGSTS, 16 B
// per lane, zero-fill for ragged edges) and signals each stage through an mbarrier
// (cp.async.mbarrier.arrive.noinc).  Warps 0-7 are CONSUMERS: each owns a 32x64 block of the tile
// (64 DMMA.8x8x4 accumulators), waits on the stage's "full" barrier, applies diag(d) to its A fragments,
// issues the MMAs and releases the stage on its "empty" barrier.  No __syncthreads in steady state, the
// producer runs ahead across tile boundaries so the epilogue of tile i overlaps the loads of tile i+1.
// tcgen05/TMEM are not usable here: the f64 kind does not exist (ptxas rejects tcgen05.mma.kind::f64).
#pragma once
#include "common.cuh"
#include "dmma_gemm.cuh"

namespace ipm {

constexpr int WS_BM = 128, WS_BN = 128;
// Stage geometry: BK columns per stage, LD = BK + 4 doubles per row (conflict-free 64-bit fragment reads: the row
// stride is 8 banks mod 32 for both), as many stages as fit 227 KB.  BK = 16 x 5 stages is the round-1 kernel; BK = 32 x 3
// halves the number of stage boundaries (mbarrier wait + release per consumer warp) per tile.
template <int BK> struct WsGeom {
    static constexpr int LD = BK + 4;
    static constexpr int STAGES = (BK == 16) ? 5 : 3;
    static constexpr size_t smem = (size_t)STAGES * ((WS_BM + WS_BN) * LD + 2 * BK) * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
};
constexpr int WS_CONSUMER_WARPS = 8;
constexpr int WS_PRODUCER_WARPS = 4;    // warps 8, 10 stream the two 64-row halves of P (warp 8 also d, v), warps 9, 11 those of Q
// Three warpgroups: two of consumers, one of producers.
// ptxas sizes the launch for 384 threads (168 registers each, the same it assumed for 320: it rounds the CTA up to whole
// warpgroups); the producer group then hands registers back (setmaxnreg.dec) and the consumers take them
// (setmaxnreg.inc): 8 x 32 x 224 + 4 x 32 x 56 = 64512 = 384 x 168.  With 168 registers the 128 accumulator registers left
// no room to hold the fragments of the next k-step: 56 B of spills in the inner loop (ncu: 29.6 M local loads per 2048
// LPs) and every k-step began with an exposed shared-memory load -> scale -> first MMA chain.
constexpr int WS_THREADS = 384;
constexpr int WS_REGS_CONSUMER = 224, WS_REGS_PRODUCER = 56;


#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {      // one non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}

// tile index inside one matrix -> (bi, bj) with bj <= bi
__device__ __forceinline__ void tri_decode(int t, int& bi, int& bj) {
    if (t < 3) {                 // matrices of up to two tile rows (the batched workload): no double-precision sqrt per tile
        bi = t > 0;
        bj = t > 1;
        return;
    }
    int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    bi = i;
    bj = t - i * (i + 1) / 2;
}

// Diagonal tile (bi == bj): only the lower triangle is needed.  The 128 rows are cut into 16 strips of 8 rows;
// strip s needs the 8x8 sub-tiles 0..s.  Consumer warp W takes strips W and 15-W = (W+1) + (16-W) = 17 sub-tiles
// for every warp, against 32 in a full tile, so a diagonal tile costs 53 % of an off-diagonal one.
// ONE code path for all warps (W is a run-time value): strip 15-W always needs sub-tiles 0..7 and strip W sub-tile 0
// (nine static accumulators); the other eight accumulators are "slots": slot i <= W is sub-tile i of strip W, slot
// i > W is sub-tile 7+i-W of strip 15-W (i = 1..8).  A slot picks its A fragment with a select and its B fragment with a run-time row offset, the
// register indices stay static.  (Round 1 had eight instantiations behind a switch on the warp index; with the
// fragment prefetch below ptxas then rotated the loop-carried fragments through local memory.)
template <int EPI, bool SCALE, int WS_BK, bool RHS>
__device__ __forceinline__ void ws_diag_tile(const DmmaArgs& a, const double* Ps, const double* Qs, const double* Ds,
                                             const double* Vs, uint64_t* full, uint64_t* empty, uint32_t& it, int nk, int z,
                                             int row0, int lane, int W) {
    static_assert(!RHS || (SCALE && EPI == 0), "the right-hand side rides on the scaled product");
    constexpr int LD = WsGeom<WS_BK>::LD, S = WsGeom<WS_BK>::STAGES;
    constexpr int NS = 8, NV = 8;
    const int R0 = 8 * W, R1 = 8 * (15 - W);
    const int g = lane >> 2, t = lane & 3;
    double accS[NS][2], accV[NV][2], acc00[2];      // acc00: sub-tile 0 of strip W
    int subv[NV];                  // column sub-tile of slot i + 1
    bool lo[NV];                   // slot belongs to strip W
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        lo[i] = (i + 1 <= W);
        subv[i] = lo[i] ? i + 1 : 8 + i - W;
    }
    double* C = a.C + (size_t)z * a.strideC;
    // EPI == 1 (C -= P Q^T): the accumulators start from C, loaded here so that the latency hides behind the
    // first pipeline stages, and the A fragments are negated; the epilogue is then a plain store.
    auto init = [&](int r, int c, double& v0, double& v1) {
        v0 = v1 = 0.0;
        if (EPI == 1 && r < a.rowsP && c < a.rowsQ) {
            const double* cp = C + (size_t)r * a.ldc + c;
            if (c + 1 < a.rowsQ) { const double2 v = *reinterpret_cast<const double2*>(cp); v0 = v.x; v1 = v.y; }
            else v0 = cp[0];
        }
    };
#pragma unroll
    for (int j = 0; j < NS; ++j) init(row0 + R1 + g, row0 + 8 * j + 2 * t, accS[j][0], accS[j][1]);
#pragma unroll
    for (int i = 0; i < NV; ++i) init(row0 + (lo[i] ? R0 : R1) + g, row0 + 8 * subv[i] + 2 * t, accV[i][0], accV[i][1]);
    init(row0 + R0 + g, row0 + 2 * t, acc00[0], acc00[1]);
    // Fragments of the NEXT k-step (and, at a stage boundary, of the next stage) are loaded and scaled while the MMAs of
    // the current one issue.
    double a0 = 0.0, a1 = 0.0, bs[NS], bv[NV], a0n, a1n, bsn[NS], bvn[NV];
    // RHS: lane (g,t) sums its columns k = t (mod 4) of rows R0+g, R1+g.  ND partial sums per row: 1 measured faster than
    // one per k-step of a stage (SYRK phase 170.7 vs 178.9 ms per solve; the FMA of the previous k-step has long left
    // the FP64 pipe when the next one issues, and the extra registers cost more than the shorter chain returns)
    constexpr int ND = 1;
    double vk = 0.0, vkn = 0.0, dot0[ND], dot1[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) dot0[i] = dot1[i] = 0.0;
    const int pa0 = (R0 + g) * LD + t, pa1 = (R1 + g) * LD + t, pb = g * LD + t;
    auto load_frag = [&](int s, int kk, double& x0, double& x1, double (&bS)[NS], double (&bV)[NV], double& vv) {
        const double* ps = Ps + s * WS_BM * LD + kk;
        const double* qs = Qs + s * WS_BN * LD + pb + kk;
        x0 = ps[pa0];
        x1 = ps[pa1];
        if (SCALE) {
            const double dv = Ds[s * WS_BK + t + kk];
            const double dk = (EPI == 1) ? -dv : dv;
            x0 *= dk;
            x1 *= dk;
        } else if (EPI == 1) {
            x0 = -x0;
            x1 = -x1;
        }
        if (RHS) vv = Vs[s * WS_BK + t + kk];
#pragma unroll
        for (int j = 0; j < NS; ++j) bS[j] = qs[j * 8 * LD];
#pragma unroll
        for (int i = 0; i < NV; ++i) bV[i] = qs[subv[i] * 8 * LD];
    };
    auto mma = [&](const double x0, const double x1, const double (&bS)[NS], const double (&bV)[NV], const double vv,
                   double& d0, double& d1) {
        if (RHS) {
            d0 = fma(x0, vv, d0);
            d1 = fma(x1, vv, d1);
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) dmma884(accS[j][0], accS[j][1], x1, bS[j]);
        dmma884(acc00[0], acc00[1], x0, bS[0]);
#pragma unroll
        for (int i = 0; i < NV; ++i) dmma884(accV[i][0], accV[i][1], lo[i] ? x0 : x1, bV[i]);
    };
    if (nk > 0) {
        mbar_wait(full + it % S, (it / S) & 1);
        load_frag(it % S, 0, a0, a1, bs, bv, vk);
    }
    for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % S;
        // The NEXT stage is probed here, at the top (the producer runs S - 1 stages ahead: it is complete), so that the
        // body is straight-line code: two fragment sets in ping-pong (no register copies), one feeds the MMAs while
        // the other is loaded and scaled; the last load of the body already reads the next stage.  After the last
        // stage of a tile that load re-reads the current one (harmless, nothing uses it).
        int sn = s;
        uint32_t pn = 0;
        bool ready = true;
        if (kt + 1 < nk) {
            const uint32_t it1 = it + 1;
            sn = it1 % S;
            pn = (it1 / S) & 1;
            ready = mbar_test(full + sn, pn);      // probed here, looked at only before the first load from that stage:
        }                                          // the probe's latency hides behind a k-step of MMAs
#pragma unroll
        for (int kk = 0; kk < WS_BK; kk += 8) {
            load_frag(s, kk + 4, a0n, a1n, bsn, bvn, vkn);
            mma(a0, a1, bs, bv, vk, dot0[(kk / 4) % ND], dot1[(kk / 4) % ND]);
            if (kk + 8 < WS_BK) {
                load_frag(s, kk + 8, a0, a1, bs, bv, vk);
            } else {
                if (!ready) mbar_wait(full + sn, pn);
                load_frag(sn, 0, a0, a1, bs, bv, vk);
            }
            mma(a0n, a1n, bsn, bvn, vkn, dot0[(kk / 4 + 1) % ND], dot1[(kk / 4 + 1) % ND]);
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
    for (int j = 0; j < NS; ++j) store(row0 + R1 + g, row0 + 8 * j + 2 * t, accS[j][0], accS[j][1]);
#pragma unroll
    for (int i = 0; i < NV; ++i) store(row0 + (lo[i] ? R0 : R1) + g, row0 + 8 * subv[i] + 2 * t, accV[i][0], accV[i][1]);
    store(row0 + R0 + g, row0 + 2 * t, acc00[0], acc00[1]);
    if (RHS) {
        // the four lanes of a row group hold the columns k = t (mod 4): fixed shuffle tree, then one lane writes
        double s0 = dot0[0], s1 = dot1[0];
#pragma unroll
        for (int i = 1; i < ND; ++i) { s0 += dot0[i]; s1 += dot1[i]; }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (t == 0) {
            const double* rb = a.rbvec + (size_t)z * a.strideR;
            double* out = a.rhs + (size_t)z * a.strideR;
            const int r0 = row0 + R0 + g, r1 = row0 + R1 + g;
            if (r0 < a.rowsP) out[r0] = -rb[r0] - s0;
            if (r1 < a.rowsP) out[r1] = -rb[r1] - s1;
        }
    }
}

// Producer side of the ring (shared by the 8- and the 16-consumer kernel): producer warp pw of NPW streams its rows of
// the P (even pw) or Q (odd pw) slabs of every tile this CTA owns, warp 0 also the matching entries of d and v.
template <bool SCALE, int WS_BK, bool RHS, int NPW>
__device__ __forceinline__ void ws_produce(const DmmaArgs& a, double* Ps, double* Qs, double* Ds, double* Vs,
                                           uint64_t* full, uint64_t* empty, int ntri, int total_tiles, int pw, int lane) {
    constexpr int LD = WsGeom<WS_BK>::LD, S = WsGeom<WS_BK>::STAGES;
    constexpr int HK = WS_BK / 2;                 // lanes that cover one row of a slab (16 bytes each)
    constexpr int RPP = 32 / HK;                  // rows per pass of a producer warp
    const int K = a.K;
    const int nk = (K + WS_BK - 1) / WS_BK;
    uint32_t it = 0;
    const bool isQ = (pw & 1) != 0;
    constexpr int HROWS = WS_BM / (NPW / 2);      // rows of a slab per producer warp
    const int rhalf = (pw >> 1) * HROWS;
    const bool aux = pw == 0;                   // this warp also brings d and v
    const int rsub = rhalf + lane / HK;         // row inside the slab: rsub + RPP * j
    const int kq = (lane % HK) * 2;             // this lane's column pair inside the slab
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / ntri;
        if (a.active && a.active[z] == 0) continue;
        int bi, bj;
        if (a.col0_only) { bi = tile - z * ntri; bj = 0; }
        else tri_decode(tile - z * ntri, bi, bj);
        const double* base = isQ ? a.Q + (size_t)z * a.strideQ : a.P + (size_t)z * a.strideP;
        const int64_t ld = isQ ? a.ldq : a.ldp;
        const int nrows = isQ ? a.rowsQ : a.rowsP;
        const int r0 = (isQ ? bj : bi) * WS_BM;
        const double* dv = (SCALE && aux) ? a.dvec + (size_t)z * a.strideD : nullptr;
        const double* vv = (RHS && aux) ? a.vvec + (size_t)z * a.strideV : nullptr;
        const bool rows_full = r0 + WS_BM <= nrows;
        const double* src0 = base + (size_t)(r0 + rsub) * ld + kq;
        const size_t rstep = (size_t)RPP * ld;
        for (int kt = 0; kt < nk; ++kt, ++it) {
            const int s = it % S;
            mbar_wait(empty + s, ((it / S) & 1) ^ 1);
            double* dst = (isQ ? Qs + s * WS_BN * LD : Ps + s * WS_BM * LD) + rsub * LD + kq;
            const int k0 = kt * WS_BK;
            if (rows_full && k0 + WS_BK <= K) {
                const double* src = src0 + k0;
#pragma unroll 8                 // (the producers run on 56 registers: no 32 precomputed addresses)
                for (int j = 0; j < HROWS / RPP; ++j) cp_async16_zfill(dst + j * RPP * LD, src + j * rstep, 16u);
                if (dv != nullptr && lane < HK) cp_async16_zfill(Ds + s * WS_BK + kq, dv + k0 + kq, 16u);
                if (RHS && vv != nullptr && lane >= HK && lane < 2 * HK) cp_async16_zfill(Vs + s * WS_BK + kq, vv + k0 + kq, 16u);
            } else {
                const int k = k0 + kq;
                const uint32_t kbytes = (k + 1 < K) ? 16u : ((k < K) ? 8u : 0u);
#pragma unroll 4
                for (int j = 0; j < HROWS / RPP; ++j) {
                    const int gr = r0 + rsub + RPP * j;
                    const uint32_t nb = (gr < nrows) ? kbytes : 0u;
                    const double* src = nb ? base + (size_t)gr * ld + k : base;
                    cp_async16_zfill(dst + j * RPP * LD, src, nb);
                }
                if (dv != nullptr && lane < HK) cp_async16_zfill(Ds + s * WS_BK + kq, kbytes ? dv + k : dv, kbytes);
                if (RHS && vv != nullptr && lane >= HK && lane < 2 * HK)
                    cp_async16_zfill(Vs + s * WS_BK + kq, kbytes ? vv + k : vv, kbytes);
            }
            cp_async_mbar_arrive_noinc(full + s);
        }
    }
}

template <int EPI, bool SCALE, int WS_BK, bool RHS = false>
__global__ void __launch_bounds__(WS_THREADS, 1) dmma_ws_kernel(const DmmaArgs a, int ntri, int total_tiles) {
    constexpr int LD = WsGeom<WS_BK>::LD, S = WsGeom<WS_BK>::STAGES;
    constexpr int HK = WS_BK / 2;                 // lanes that cover one row of a slab (16 bytes each)
    constexpr int RPP = 32 / HK;                  // rows per pass of a producer warp
    extern __shared__ __align__(128) unsigned char ws_raw[];
    double* Ps = reinterpret_cast<double*>(ws_raw);            // [S][128][LD]
    double* Qs = Ps + S * WS_BM * LD;                          // [S][128][LD]
    double* Ds = Qs + S * WS_BN * LD;                          // [S][BK]
    double* Vs = Ds + S * WS_BK;                               // [S][BK]  (RHS only)
    uint64_t* full = reinterpret_cast<uint64_t*>(Vs + S * WS_BK);
    uint64_t* empty = full + S;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 32 * WS_PRODUCER_WARPS);   // every producer lane arrives once (noinc)
            mbar_init(empty + s, WS_CONSUMER_WARPS);   // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int K = a.K;
    const int nk = (K + WS_BK - 1) / WS_BK;
    uint32_t it = 0;                                    // ring position, continues across tiles

    if (warp >= WS_CONSUMER_WARPS) {
        setmaxnreg_dec<WS_REGS_PRODUCER>();
        if (warp >= WS_CONSUMER_WARPS + WS_PRODUCER_WARPS) return;      // the warpgroup's spare warps
        // ------------------------------------------------------------------ producers
        ws_produce<SCALE, WS_BK, RHS, WS_PRODUCER_WARPS>(a, Ps, Qs, Ds, Vs, full, empty, ntri, total_tiles, warp - WS_CONSUMER_WARPS, lane);
    } else {
        // ------------------------------------------------------------------ consumers: 4 x 2 warps, 32 x 64 each
        setmaxnreg_inc<WS_REGS_CONSUMER>();
        constexpr int MI = 4, NI = 8;
        const int g = lane >> 2, t = lane & 3;
        const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 64;
        // the active flag of the NEXT tile's matrix is fetched while this tile is computed (a global-memory round trip
        // at every tile boundary otherwise: all eight consumers stall on it together and the pipe drains)
        int act_next = (a.active && (int)blockIdx.x < total_tiles) ? a.active[blockIdx.x / ntri] : 1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int z = tile / ntri;
            const int act = act_next;
            {
                const int tn = tile + gridDim.x;
                act_next = (a.active && tn < total_tiles) ? a.active[tn / ntri] : 1;
            }
            if (act == 0) continue;
            int bi, bj;
            if (a.col0_only) { bi = tile - z * ntri; bj = 0; }
            else tri_decode(tile - z * ntri, bi, bj);
            if (bi == bj) {
                const int r0d = bi * WS_BM;
                ws_diag_tile<EPI, SCALE, WS_BK, RHS>(a, Ps, Qs, Ds, Vs, full, empty, it, nk, z, r0d, lane, warp);
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
                    if (EPI == 1 && r < a.rowsP && c < a.rowsQ) {       // accumulate on top of C (see ws_diag_tile)
                        const double* cp = C + (size_t)r * a.ldc + c;
                        if (c + 1 < a.rowsQ) { const double2 v = *reinterpret_cast<const double2*>(cp); v0 = v.x; v1 = v.y; }
                        else v0 = cp[0];
                    }
                    acc[i][j][0] = v0;
                    acc[i][j][1] = v1;
                }
            }
            double af[MI], bf[NI], afn[MI], bfn[NI];
            auto load_frag = [&](int s, int kk, double (&xa)[MI], double (&xb)[NI]) {
                const double* ps = Ps + s * WS_BM * LD + (wm0 + g) * LD + t + kk;
                const double* qs = Qs + s * WS_BN * LD + (wn0 + g) * LD + t + kk;
#pragma unroll
                for (int i = 0; i < MI; ++i) xa[i] = ps[i * 8 * LD];
#pragma unroll
                for (int j = 0; j < NI; ++j) xb[j] = qs[j * 8 * LD];
                if (SCALE) {
                    const double dv = Ds[s * WS_BK + t + kk];
                    const double dk = (EPI == 1) ? -dv : dv;
#pragma unroll
                    for (int i = 0; i < MI; ++i) xa[i] *= dk;
                } else if (EPI == 1) {
#pragma unroll
                    for (int i = 0; i < MI; ++i) xa[i] = -xa[i];
                }
            };
            if (nk > 0) {
                mbar_wait(full + it % S, (it / S) & 1);
                load_frag(it % S, 0, af, bf);
            }
            auto mma = [&](const double (&xa)[MI], const double (&xb)[NI]) {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], xa[i], xb[j]);
            };
            for (int kt = 0; kt < nk; ++kt, ++it) {
                const int s = it % S;
                int sn = s;                                      // see ws_diag_tile
                uint32_t pn = 0;
                bool ready = true;
                if (kt + 1 < nk) {
                    const uint32_t it1 = it + 1;
                    sn = it1 % S;
                    pn = (it1 / S) & 1;
                    ready = mbar_test(full + sn, pn);
                }
#pragma unroll
                for (int kk = 0; kk < WS_BK; kk += 8) {
                    load_frag(s, kk + 4, afn, bfn);
                    mma(af, bf);
                    if (kk + 8 < WS_BK) {
                        load_frag(s, kk + 8, af, bf);
                    } else {
                        if (!ready) mbar_wait(full + sn, pn);
                        load_frag(sn, 0, af, bf);
                    }
                    mma(afn, bfn);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            // epilogue (the producer is already filling the ring for the next tile)
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
}

// true when the operands satisfy the 16-byte alignment the cp.async path and the vector epilogue need
inline bool ws_eligible(const DmmaArgs& a) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return al16(a.P) && al16(a.Q) && al16(a.C) && (a.ldp % 2 == 0) && (a.ldq % 2 == 0) && (a.ldc % 2 == 0) &&
           (a.strideP % 2 == 0) && (a.strideQ % 2 == 0) && (a.strideC % 2 == 0) &&
           (!a.dvec || (al16(a.dvec) && a.strideD % 2 == 0)) && a.lower_only && a.rowsP == a.rowsQ;
}

inline std::atomic<int>& ws_stage_width() {          // 16 (default) or 32: ipm_set_syrk_stage_width, A/B measurements
    static std::atomic<int> bk{16};
    return bk;
}

template <int EPI, bool SCALE, int BK, bool RHS = false>
inline int dmma_ws_launch_bk(const DmmaArgs& a, int batch, cudaStream_t st) {
    auto kern = dmma_ws_kernel<EPI, SCALE, BK, RHS>;
    constexpr size_t smem = WsGeom<BK>::smem;
    IPM_TRY(ensure_dyn_smem(kern, smem));
    if (a.rowsP <= 0 || batch <= 0 || a.K <= 0) return IPM_OK;
    const int T = ceil_div(a.rowsP, WS_BM);
    const int ntri = a.col0_only ? T : T * (T + 1) / 2;      // tiles per matrix
    const int64_t total = (int64_t)ntri * batch;
    if (total > 0x7fffffff) return IPM_ERR_SHAPE;
    const int cap = (a.max_ctas > 0 && a.max_ctas < kNumSMs) ? a.max_ctas : kNumSMs;
    const int grid = (int)std::min<int64_t>(total, cap);
    kern<<<grid, WS_THREADS, smem, st>>>(a, ntri, (int)total);
    count_launch();
    return launch_check();
}
// 16-consumer variant (dmma_ws16.cuh, included at the end of this file)
inline std::atomic<int>& ws_consumer_warps();
template <int EPI, bool SCALE, bool RHS>
inline int dmma_ws16_launch(const DmmaArgs& a, int batch, cudaStream_t st);

template <int EPI, bool SCALE>
inline int dmma_ws_launch(const DmmaArgs& a, int batch, cudaStream_t st) {
    if (ws_consumer_warps().load() == 16) return dmma_ws16_launch<EPI, SCALE, false>(a, batch, st);
    if (ws_stage_width().load() == 32) return dmma_ws_launch_bk<EPI, SCALE, 32>(a, batch, st);
    return dmma_ws_launch_bk<EPI, SCALE, 16>(a, batch, st);
}

// SYRK-shaped product with the best available kernel: the warp-specialised one when the alignment allows,
// the register-staged one (arbitrary leading dimensions) otherwise.
template <int EPI>
inline int dmma_syrk_auto(const DmmaArgs& a, int batch, cudaStream_t st) {
    if (a.vvec != nullptr) {             // right-hand side fused into the diagonal tiles: the caller checked ws_eligible
        if constexpr (EPI == 0) {
            if (!ws_eligible(a) || !a.dvec || !a.rbvec || !a.rhs || (a.strideV & 1) ||
                (reinterpret_cast<uintptr_t>(a.vvec) & 15))
                return IPM_ERR_ARG;
            if (ws_consumer_warps().load() == 16) return dmma_ws16_launch<0, true, true>(a, batch, st);
            return dmma_ws_launch_bk<0, true, 16, true>(a, batch, st);
        } else {
            return IPM_ERR_ARG;
        }
    }
    if (ws_eligible(a)) {
        if (a.dvec) return dmma_ws_launch<EPI, true>(a, batch, st);
        return dmma_ws_launch<EPI, false>(a, batch, st);
    }
    return dmma_nt_launch<128, 128, 4, 2, EPI>(a, batch, st);
}
#endif

}  // namespace ipm

#include "dmma_ws16.cuh"
