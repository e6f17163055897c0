// Micro-benchmark: issue-rate ceilings of the FP64 pipes on one B200 (no memory traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
// Prints DMMA.8x8x4 and DFMA throughput in TFLOP/s for several resident-warp counts; the DMMA figure is the
// practical FP64 tensor peak the SYRK/Cholesky roofline fractions are quoted against (MEASURED_PEAKS.json
// carries no FP64 number).
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

int main() {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            k_dmma<16><<<nsm, warps * 32>>>(out, iters);
            cudaEventRecord(e0);
            k_dmma<16><<<nsm, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flop = 2.0 * 256 * 16.0 * iters * warps * nsm;
            if (rep) printf("DMMA  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
        }
    }
    for (int warps = 4; warps <= 32; warps *= 2) {
        k_dfma<16><<<nsm, warps * 32>>>(out, iters);
        cudaEventRecord(e0);
        k_dfma<16><<<nsm, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 32 * 16.0 * iters * warps * nsm;
        printf("DFMA  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
    }
    printf("SMs=%d err=%s\n", nsm, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
