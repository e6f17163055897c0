// Where do the ~400 cycles per column of the 32x32 diagonal-block factorisation of kb_chol go?  One warp alone on an
// SM, lane i keeps row i in registers (as in chol_batched.cuh); variants isolate the pivot chain, the trailing
// multiply-adds and the trip of the column through shared memory.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LD = 34;
template <int V>
__global__ void probe(const double* M, double* out, long long* cyc) {
    __shared__ double D[32 * LD];
    __shared__ __align__(16) double colb[64];
    const int lane = threadIdx.x;
    for (int c = 0; c < 32; ++c) D[lane * LD + c] = M[lane * 32 + c];
    __syncwarp();
    long long total = 0;
    double keep = 0.0;
    for (int rep = 0; rep < 8; ++rep) {
        double arow[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) arow[c] = D[lane * LD + c];
        double mydiag = D[lane * LD + lane];
        const double thresh = 1e-30;
        const long long t0 = clock64();
        if (V == 0 || V == 1 || V == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                double p = __shfl_sync(0xffffffffu, mydiag, j);
                const bool bad = !(p > thresh);
                if (bad) p = 1e128;
                const double inv = rsqrt(p);
                const double lij = (lane == j) ? p * inv : arow[j] * inv;
                mydiag = fma(-lij, lij, mydiag);
                arow[j] = lij;
                if (V == 1) continue;                                  // chain only
                double* cb = colb + (j & 1) * 32;
                cb[lane] = lij;
                __syncwarp();
                if (V == 2) { keep += cb[(lane + 1) & 31]; continue; }  // chain + shared-memory trip, no trailing update
                const double2* cb2 = reinterpret_cast<const double2*>(cb);
#pragma unroll
                for (int k2 = (j + 1) >> 1; k2 < 16; ++k2) {
                    const double2 v = cb2[k2];
                    if (2 * k2 > j) arow[2 * k2] = fma(-lij, v.x, arow[2 * k2]);
                    arow[2 * k2 + 1] = fma(-lij, v.y, arow[2 * k2 + 1]);
                }
            }
        } else if (V == 3) {
            // blocked: columns in groups of 8; inside a group the trailing update touches the group only, the
            // columns to the right are updated once per group (8 columns at a time, all multiply-adds independent)
            __shared__ __align__(16) double grp[8 * 32];
#pragma unroll
            for (int jb = 0; jb < 32; jb += 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = jb + q;
                    double p = __shfl_sync(0xffffffffu, mydiag, j);
                    const bool bad = !(p > thresh);
                    if (bad) p = 1e128;
                    const double inv = rsqrt(p);
                    const double lij = (lane == j) ? p * inv : arow[j] * inv;
                    mydiag = fma(-lij, lij, mydiag);
                    arow[j] = lij;
                    grp[q * 32 + lane] = lij;
                    __syncwarp();
#pragma unroll
                    for (int k = j + 1; k < jb + 8; ++k) arow[k] = fma(-lij, grp[q * 32 + k], arow[k]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double lij = arow[jb + q];
                    const double2* g2 = reinterpret_cast<const double2*>(grp + q * 32);
#pragma unroll
                    for (int k2 = (jb + 8) >> 1; k2 < 16; ++k2) {
                        const double2 v = g2[k2];
                        arow[2 * k2] = fma(-lij, v.x, arow[2 * k2]);
                        arow[2 * k2 + 1] = fma(-lij, v.y, arow[2 * k2 + 1]);
                    }
                }
                __syncwarp();
            }
        }
        const long long t1 = clock64();
        total += t1 - t0;
#pragma unroll
        for (int c = 0; c < 32; ++c) keep += arow[c];
        keep += mydiag;
    }
    out[lane] = keep;
    if (lane == 0) cyc[V] = total / 8;
}
int main() {
    double h[32 * 32];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) h[i * 32 + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1 + abs(i - j));
    double *M, *out; long long* cyc;
    cudaMalloc(&M, sizeof(h)); cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
    cudaMemcpy(M, h, sizeof(h), cudaMemcpyHostToDevice);
    probe<0><<<1, 32>>>(M, out, cyc); probe<1><<<1, 32>>>(M, out, cyc); probe<2><<<1, 32>>>(M, out, cyc); probe<3><<<1, 32>>>(M, out, cyc);
    long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
    const char* nm[] = {"as in kb_chol (column through shared memory, trailing multiply-adds)", "pivot chain only", "pivot chain + shared-memory trip", "columns in groups of 8, right part updated once per group"};
    for (int v = 0; v < 4; ++v) printf("%-75s %6lld cycles per block = %5.1f per column\n", nm[v], c[v], c[v] / 32.0);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
