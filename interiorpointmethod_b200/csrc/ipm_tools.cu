// Measuring sticks that ship with the library (no product path depends on them).
#include "common.cuh"

using namespace ipm;

namespace {
// 16 independent accumulators per warp: DMMA.8x8x4 issue-rate loop, no memory traffic.
__global__ void k_dmma_issue(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
}  // namespace

extern "C" double ipm_measure_dmma_peak(int device_ordinal) {
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return -1.0;
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device_ordinal);
    double* out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int warps = 16, iters = 20000;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, 0);
        k_dmma_issue<<<nsm, warps * 32>>>(out, iters);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 256 * 16.0 * iters * warps * nsm / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    count_launch(4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}
