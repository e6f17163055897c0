// Micro-benchmark: DMMA.8x8x4 issue rate with a REALISTIC operand stream (register-tiled 32x64 block: 4 A fragments
// x 8 B fragments -> 32 accumulators, every MMA reads a different (A, B, C) register triple) against the constant-operand
// loop of tools/dmma_peak.cu that defines the ceiling.  No memory traffic in either.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_operand_probe dmma_operand_probe.cu && ./dmma_operand_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MI x NI register tile; ROT > 0 rotates the fragments every trip like a k-loop that loads new ones
template <int MI, int NI, int ROT>
__global__ void k_tile(double* out, int iters) {
    double c[MI][NI][2], a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
        for (int j = 0; j < NI; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) b[j] = 1.0 - (threadIdx.x + j) * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
        if (ROT) {
            const double t = a[0];
#pragma unroll
            for (int i = 0; i + 1 < MI; ++i) a[i] = a[i + 1];
            a[MI - 1] = t;
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) out[0] = s;
}

// the same 4 x 8 tile fed from shared memory like the SYRK consumers: per k-step 12 LDS.64 (conflict-free layout,
// row stride 20 doubles), LDSMODE 1: fragments loaded one step ahead (ping-pong), 2: plus one DMUL per A fragment
template <int LDSMODE>
__global__ void k_tile_lds(double* out, int iters) {
    __shared__ double sm[2][128 * 20];
    for (int i = threadIdx.x; i < 2 * 128 * 20; i += blockDim.x) (&sm[0][0])[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const double* ps = &sm[0][0] + ((warp & 3) * 32 + g) * 20 + t;
    const double* qs = &sm[1][0] + ((warp & 1) * 64 + g) * 20 + t;
    double c[4][8][2], a[4], b[8], an[4], bn[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    auto load = [&](int kk, double (&xa)[4], double (&xb)[8]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xa[i] = ps[i * 8 * 20 + kk];
#pragma unroll
        for (int j = 0; j < 8; ++j) xb[j] = qs[j * 8 * 20 + kk];
        if (LDSMODE == 2) {
            const double dk = qs[kk + 3 * 20];
#pragma unroll
            for (int i = 0; i < 4; ++i) xa[i] *= dk;
        }
    };
    auto mma = [&](const double (&xa)[4], const double (&xb)[8]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma(c[i][j][0], c[i][j][1], xa[i], xb[j]);
    };
    load(0, a, b);
    for (int it = 0; it < iters; it += 4) {
        load(4, an, bn); mma(a, b);
        load(8, a, b); mma(an, bn);
        load(12, an, bn); mma(a, b);
        load(0, a, b); mma(an, bn);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) out[0] = s;
}
template <int LDSMODE>
static void run_lds(double* out, int nsm, int warps, const char* what) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8000;
    k_tile_lds<LDSMODE><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e0);
    k_tile_lds<LDSMODE><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 256 * 32 * (double)iters * warps * nsm;
    printf("%-44s warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", what, warps, ms, flop / ms * 1e-9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

template <int MI, int NI, int ROT>
static void run(double* out, int nsm, int warps, const char* what) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8000;
    k_tile<MI, NI, ROT><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e0);
    k_tile<MI, NI, ROT><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 256 * MI * NI * (double)iters * warps * nsm;
    printf("%-44s warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", what, warps, ms, flop / ms * 1e-9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    double* out; cudaMalloc(&out, 8);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    for (int warps = 4; warps <= 8; warps *= 2) {
        run<1, 16, 0>(out, nsm, warps, "1 x 16 tile (one A fragment, 16 B)");
        run<4, 8, 0>(out, nsm, warps, "4 x 8 tile, fixed fragments");
        run<4, 8, 1>(out, nsm, warps, "4 x 8 tile, A fragments rotated per trip");
        run<2, 8, 0>(out, nsm, warps, "2 x 8 tile");
        run<4, 4, 0>(out, nsm, warps, "4 x 4 tile");
        run<8, 4, 0>(out, nsm, warps, "8 x 4 tile");
        run_lds<1>(out, nsm, warps, "4 x 8 tile, fragments from shared memory");
        run_lds<2>(out, nsm, warps, "4 x 8 tile, shared memory + scaling (DMUL)");
    }
    printf("SMs=%d err=%s\n", nsm, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
