// Micro-benchmark: do DMMA.8x8x4 (FP64 tensor) and DFMA (FP64 vector) share an execution pipe on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix_probe fp64_mix_probe.cu && ./fp64_mix_probe
// Each CTA has 16 warps; warps [0, nd) run an issue-rate DMMA loop, the rest a DFMA loop of the SAME duration when
// run alone (iteration counts calibrated from the solo rates).  If the pipes are distinct the mixed launch takes
// as long as the longer solo part and the combined rate approaches the SUM of the two peaks; if DMMA is executed
// on the DFMA units the combined rate stays at the single-pipe ceiling.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_mix(double* out, int nd_warps, int it_dmma, int it_dfma) {
    const int warp = threadIdx.x >> 5;
    double s = 0.0;
    if (warp < nd_warps) {
        double c[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
        const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
        for (int it = 0; it < it_dmma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    } else {
        double c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = i;
        const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
        for (int it = 0; it < it_dfma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i];
    }
    if (s == 123.456) out[0] = s;
}

// the same question inside ONE warp: every loop trip issues 16 DMMA and `r` x 16 DFMA on independent registers
template <int R>
__global__ void k_interleave(double* out, int iters) {
    double c[16][2], f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = c[i][1] = 0.0; f[i] = i; }
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int r = 0; r < R; ++r) f[(i + r) & 15] = fma(f[(i + r) & 15], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) out[0] = s;
}

static float run_mix(double* out, int nsm, int nd, int itd, int itf) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<<<nsm, 512>>>(out, nd, itd, itf);
    cudaEventRecord(e0);
    k_mix<<<nsm, 512>>>(out, nd, itd, itf);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

template <int R>
static void run_il(double* out, int nsm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, warps = 16;
    k_interleave<R><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e0);
    k_interleave<R><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fd = 2.0 * 256 * 16.0 * iters * warps * nsm, ff = 2.0 * 32 * 16.0 * R * iters * warps * nsm;
    printf("interleaved in one warp, %d DFMA per DMMA: %.3f ms  DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", R, ms,
           fd / ms * 1e-9, ff / ms * 1e-9, (fd + ff) / ms * 1e-9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    double* out; cudaMalloc(&out, 8);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    // one DMMA = 512 flop per warp instruction, one DFMA warp instruction = 64 flop: equal flops per warp = 8 DFMA trips
    const int itd = 10000;
    for (int nd = 0; nd <= 16; nd += 4) {
        const int itf = itd * 8;
        const float ms = run_mix(out, nsm, nd, itd, itf);
        const double fd = 2.0 * 256 * 16.0 * itd * nd * nsm, ff = 2.0 * 32 * 16.0 * itf * (16 - nd) * nsm;
        printf("mixed CTA: %2d DMMA warps + %2d DFMA warps: %.3f ms  DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", nd, 16 - nd,
               ms, fd / ms * 1e-9, ff / ms * 1e-9, (fd + ff) / ms * 1e-9);
    }
    run_il<1>(out, nsm);
    run_il<2>(out, nsm);
    run_il<4>(out, nsm);
    run_il<8>(out, nsm);
    printf("SMs=%d err=%s\n", nsm, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
