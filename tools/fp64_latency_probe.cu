// Latencies (cycles, one warp alone on an SM) of the operations on the pivot-to-pivot chain of the 32x32 diagonal
// block factorisation in kb_chol: dependent DFMA, DMUL, rsqrt(double), shuffle of a double, and the whole pivot step.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double seed) {
    const int lane = threadIdx.x & 31;
    double x = seed + lane * 1e-3, y = 1.0000001;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // dependent DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = x * y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // dependent rsqrt
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = rsqrt(x) + 2.0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) * 4;     // per 256, includes one DADD each
    // dependent shuffle of a double
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __shfl_sync(0xffffffffu, x, (i * 7) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // the pivot step: shuffle, compare, rsqrt, multiply, own diagonal update
    double mydiag = fabs(x) + 1000.0 + lane, a = 0.5 + lane * 1e-3;
    t0 = clock64();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        double p = __shfl_sync(0xffffffffu, mydiag, j);
        if (!(p > 1e-30)) p = 1e128;
        const double inv = rsqrt(p);
        const double lij = (lane == j) ? p * inv : a * inv;
        mydiag = fma(-lij, lij, mydiag);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) * 8;     // per 256 pivots
    // 32 independent DFMAs back to back (issue rate)
    double r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = x + i;
    t0 = clock64();
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = fma(r[i], y, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    double s = x + mydiag;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += r[i];
    out[threadIdx.x] = s;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 64); cudaMalloc(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 32>>>(out, cyc, 1.5);
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* nm[] = {"dependent DFMA", "dependent DMUL", "dependent rsqrt(double)+DADD", "dependent shuffle (64-bit)", "pivot step (shfl, cmp, rsqrt, mul, fma)", "independent DFMA (issue)"};
    for (int i = 0; i < 6; ++i) printf("%-42s %7.1f cycles each\n", nm[i], h[i] / 256.0);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
