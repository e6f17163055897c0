// Shared definitions for the sm_100a Newton-step kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/ipm_b200.h"

namespace ipm {

// ---------------------------------------------------------------- launch accounting / errors
extern std::atomic<int64_t> g_launches;
extern thread_local std::string g_last_error;   // for calls without a handle

inline void count_launch(int64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define IPM_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ipm::g_last_error = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +   \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")";              \
            return IPM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define IPM_TRY(expr)                    \
    do {                                 \
        int _r = (expr);                 \
        if (_r != IPM_OK) return _r;     \
    } while (0)

inline int launch_check() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_error = std::string("kernel launch: ") + cudaGetErrorString(e);
        return IPM_ERR_CUDA;
    }
    return IPM_OK;
}

// Dynamic shared memory above 48 KB needs cudaFuncAttributeMaxDynamicSharedMemorySize on the kernel, per device.
// The registry is keyed on (kernel entry point, device ordinal): the kernels live in headers with internal linkage,
// so every translation unit of the library has its OWN instance of each - a function-local "configured" flag inside
// an inline launcher is shared by all of them (one merged variable) and leaves every instance but the first one
// unconfigured ("invalid argument" at the first launch above 48 KB; found by running the single-LP tests before the
// batched ones in one process).  Thread-safe; one map lookup per launch.
inline int ensure_dyn_smem_ptr(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> configured;
    int dev = 0;
    IPM_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = configured.find(key);
    if (it != configured.end() && it->second >= bytes) return IPM_OK;
    IPM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    configured[key] = bytes;
    return IPM_OK;
}
template <class K>
inline int ensure_dyn_smem(K kernel, size_t bytes) {
    return ensure_dyn_smem_ptr(reinterpret_cast<const void*>(kernel), (int)bytes);
}

static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr double kPivotBig = 1e128;
constexpr int kNumSMs = 148;          // B200

// ---------------------------------------------------------------- device scalar slots (one block of doubles per LP)
enum Scal {
    S_NRB2 = 0,     // |rb|^2
    S_NRC2,         // |rc|^2
    S_XS,           // x^T s
    S_NB,           // |b|
    S_NC,           // |c|
    S_AP_AFF,       // predictor step lengths
    S_AD_AFF,
    S_MU_AFF,
    S_MU,
    S_SIGMA,
    S_SIGMA_MU,     // sigma * mu
    S_AP,           // corrector step lengths
    S_AD,
    S_MAXDIAG,      // max_i M_ii
    S_OBJ,          // c^T x
    S_NRB,          // |rb|
    S_NRC,          // |rc|
    S_CONT,         // 1.0 = continue, 0.0 = stop
    S_NFIXED,       // pivots replaced in the last factorisation
    S_RAW_P,        // min({-x/dx : dx<0} U {1}) of the last direction
    S_RAW_D,        // min({-s/ds : ds<0} U {1}) of the last direction
    S_NREFINE,      // corrector refinements taken so far
    S_REFINE_FLAG,  // single LP: 1.0 = the current corrector takes the refinement step
    S_HANDOFF,      // batched solver: 1.0 = the LP was handed to the augmented-system kernel (kkt_dense.cuh)
    S_COUNT = 24
};

// ---------------------------------------------------------------- warp / block reductions (fixed order => deterministic)
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

enum RedOp { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2 };

template <int OP>
__device__ __forceinline__ double red_combine(double a, double b) {
    if (OP == RED_SUM) return a + b;
    if (OP == RED_MIN) return fmin(a, b);
    return fmax(a, b);
}
template <int OP>
__device__ __forceinline__ double red_identity() {
    if (OP == RED_SUM) return 0.0;
    if (OP == RED_MIN) return __longlong_as_double(0x7ff0000000000000LL);   // +inf
    return __longlong_as_double(0xfff0000000000000LL);                      // -inf
}
template <int OP>
__device__ __forceinline__ double warp_red(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_combine<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block reduction; result valid in thread 0.  `sh` must hold 32 doubles.  All threads must call.
template <int OP>
__device__ __forceinline__ double block_red(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_red<OP>(v);
    __syncthreads();                 // protect sh reuse
    if (lane == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = red_identity<OP>();
    if (w == 0) {
        r = (lane < nw) ? sh[lane] : red_identity<OP>();
        r = warp_red<OP>(r);
    }
    return r;
}

// Grid-wide reduction of up to NV values per block with the "last block finalises" pattern.
// Every block writes its partials, takes a ticket; the last one combines partials[0..gridDim.x) in index
// order (deterministic) and returns true in ALL its threads with the totals in out[] (thread 0 holds them).
// partials: gridDim.x * NV doubles; counter: one unsigned, zero before the launch, reset to zero on exit.
template <int NV, int OP0, int OP1 = RED_SUM, int OP2 = RED_SUM, int OP3 = RED_SUM>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* partials, unsigned* counter, double* sh,
                                            double (&out)[NV]) {
    __shared__ bool is_last;
    double r[NV];
    r[0] = block_red<OP0>(v[0], sh);
    if (NV > 1) r[1 % NV] = block_red<OP1>(v[1 % NV], sh);
    if (NV > 2) r[2 % NV] = block_red<OP2>(v[2 % NV], sh);
    if (NV > 3) r[3 % NV] = block_red<OP3>(v[3 % NV], sh);
    if (gridDim.x == 1) {
        // one block: the combine below would fold r with identities only (r + 0, min(r, inf)) - the same value without
        // the fence, the atomic and the round trip through `partials` (small LPs pay them five times per iteration)
#pragma unroll
        for (int i = 0; i < NV; ++i) out[i] = r[i];
        return true;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) partials[(size_t)blockIdx.x * NV + i] = r[i];
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // final combine by warp 0 in a fixed order: lane l takes blocks l, l+32, ... then a shuffle tree
    double acc[NV];
    acc[0] = red_identity<OP0>();
    if (NV > 1) acc[1 % NV] = red_identity<OP1>();
    if (NV > 2) acc[2 % NV] = red_identity<OP2>();
    if (NV > 3) acc[3 % NV] = red_identity<OP3>();
    if (threadIdx.x < 32) {
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) {
            const volatile double* p = partials + (size_t)b * NV;
            acc[0] = red_combine<OP0>(acc[0], p[0]);
            if (NV > 1) acc[1 % NV] = red_combine<OP1>(acc[1 % NV], p[1 % NV]);
            if (NV > 2) acc[2 % NV] = red_combine<OP2>(acc[2 % NV], p[2 % NV]);
            if (NV > 3) acc[3 % NV] = red_combine<OP3>(acc[3 % NV], p[3 % NV]);
        }
        acc[0] = warp_red<OP0>(acc[0]);
        if (NV > 1) acc[1 % NV] = warp_red<OP1>(acc[1 % NV]);
        if (NV > 2) acc[2 % NV] = warp_red<OP2>(acc[2 % NV]);
        if (NV > 3) acc[3 % NV] = warp_red<OP3>(acc[3 % NV]);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) out[i] = acc[i];
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}

// FP64 tensor-core op: D(8x8) += A(8x4, row) * B(4x8, col).  Lowers to DMMA.8x8x4 on sm_100a.
// lane = 4*g + t : A holds A[g][t], B holds B[t][g], C holds C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
#endif  // __CUDACC__

}  // namespace ipm
