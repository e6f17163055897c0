// Per-phase cycle counts of the fused Cholesky (lp 0), build: nvcc -DKBC_PROFILE ... (debug aid, not shipped)
#include <cstdio>
#include <vector>
#include "../interiorpointmethod_b200/csrc/chol_batched.cuh"
namespace ipm { std::atomic<int64_t> g_launches{0}; thread_local std::string g_last_error; }
using namespace ipm;
int main(int argc, char**) {
    const int B = 2048, m = 256; const int64_t ldm = 256;
    std::vector<double> h((size_t)m * ldm);
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) h[i * ldm + j] = (i == j ? 300.0 : 0.0) + 1.0 / (1 + abs(i - j));
    double* M; cudaMalloc(&M, (size_t)B * m * ldm * 8);
    for (int b = 0; b < B; ++b) cudaMemcpy(M + (size_t)b * m * ldm, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    double* scal; cudaMalloc(&scal, B * S_COUNT * 8);
    potrf_batched_fused(M, ldm, m * ldm, m, B, scal, S_COUNT, 1e-30, nullptr, 0);
    cudaDeviceSynchronize();
    for (int b = 0; b < B; ++b) cudaMemcpy(M + (size_t)b * m * ldm, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    long long z[32] = {0}; cudaMemcpyToSymbol(kbc_prof, z, sizeof(z));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (argc > 1) {   // one CTA per SM: pad the dynamic shared memory
        CholBatchedArgs a; a.M = M; a.ldm = ldm; a.strideM = m * ldm; a.scal = scal; a.strideScal = S_COUNT; a.tau = 1e-30; a.m = m; a.active = nullptr;
        ensure_dyn_smem(kb_chol<KBC_NT>, 120 * 1024);
        kb_chol<KBC_NT><<<B, KBC_NT, 120 * 1024>>>(a);
    } else
    potrf_batched_fused(M, ldm, m * ldm, m, B, scal, S_COUNT, 1e-30, nullptr, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpyFromSymbol(z, kbc_prof, sizeof(z));
    const char* un[] = {"early update", "-", "wait S1", "trsm+S3", "store", "late update", "wait S4", "dump+init+S5"};
    const char* fn[] = {"diag", "-", "wait S1", "trsm+S3", "store+S4+S5"};
    long long tu = 0, tf = 0;
    for (int i = 0; i < 8; ++i) tu += z[i];
    for (int i = 0; i < 5; ++i) tf += z[16 + i];
    printf("kb_chol B=%d m=%d: %.3f ms; err=%s\n", B, m, ms, cudaGetErrorString(cudaGetLastError()));
    printf("update warp (lp 0), total %lld cycles:\n", tu);
    for (int i = 0; i < 8; ++i) printf("  %-14s %8lld  %5.1f%%\n", un[i], z[i], 100.0 * z[i] / tu);
    printf("factor warp (lp 0), total %lld cycles:\n", tf);
    for (int i = 0; i < 5; ++i) printf("  %-14s %8lld  %5.1f%%\n", fn[i], z[16 + i], 100.0 * z[16 + i] / tf);
    return 0;
}
