// Device-side problem ingestion: the compressed-sparse arrays of A arrive exactly as the
// reference's loader holds them (scipy csc_matrix out of loadmat, sparse_interior.py:139-167, 211-216 - or CSR),
// and everything derived from the sparsity STRUCTURE is built on the GPU and cached per structure:
//   * the other orientation (CSC -> CSR or CSR -> CSC) with a stable, deterministic transposition,
//   * the symbolic pattern of M = A diag(d) A^T (lower triangle): entry list, term lists ordered by the shared
//     column index - bit for bit the pattern spgemm_symbolic (sparse.cuh, host) produces, so the numeric SpGEMM
//     and every parity result are unchanged.
// A second load of a problem with the same structure (same LP again, or new values on the same pattern) finds
// the pattern in the process-wide cache and only uploads val, b, c.
#pragma once
#include <chrono>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "sparse.cuh"

namespace ipm {

// Structure of one sparse A on one device, shared by every handle that loaded the same structure.
struct DevPattern {
    int dev = 0, m = 0, n = 0;
    bool from_csc = false;           // orientation the caller supplied (part of the cache key)
    int64_t nnz = 0, ldm = 0, nent = 0, nterms = 0;
    bool device_built = false;
    double build_ms = 0.0;
    int32_t* islab = nullptr;        // rowptr | colind | t_rowptr | t_colind | perm
    int32_t *rowptr = nullptr, *colind = nullptr, *t_rowptr = nullptr, *t_colind = nullptr, *perm = nullptr;
    int64_t* pslab = nullptr;        // out_idx | prod_ptr | pa | pb
    int64_t *out_idx = nullptr, *prod_ptr = nullptr;
    int32_t *pa = nullptr, *pb = nullptr;
    // cache key: the caller's structure arrays
    uint64_t hash = 0;
    std::vector<int32_t> key_ptr, key_idx;
    size_t bytes = 0;
    ~DevPattern() {
        if (islab || pslab) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(dev);
            if (islab) cudaFree(islab);
            if (pslab) cudaFree(pslab);
            cudaSetDevice(cur);
        }
    }
};

constexpr int ING_NT = 256;
constexpr int ING_SCAN_NT = 1024;
constexpr int ING_PLACE_CAP = 2048;        // keys of one destination segment staged in shared memory
constexpr size_t ING_SYM_MAX_SMEM = 224 * 1024;   // marker array of the symbolic pass: m int32 <= 224 KB

#ifdef __CUDACC__
// cnt[idx[p]] += 1 (integer atomics: the result does not depend on the order)
static __global__ void k_ing_count(int64_t nnz, const int32_t* __restrict__ idx, int32_t* __restrict__ cnt) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[idx[p]], 1);
}

// Exclusive scan by ONE CTA of ING_SCAN_NT threads: out[i] = sum_{k<i} in[k] for i = 0..N (out[N] = total).
template <class TI, class TO>
static __global__ void __launch_bounds__(ING_SCAN_NT) k_ing_scan(const TI* __restrict__ in, TO* __restrict__ out,
                                                                 int64_t N) {
    __shared__ TO wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TO carry = 0;
    for (int64_t base = 0; base < N; base += ING_SCAN_NT) {
        const int64_t i = base + tid;
        const TO v = (i < N) ? (TO)in[i] : (TO)0;
        TO x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const TO y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            TO w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const TO y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const TO prefix = carry + (warp > 0 ? wsum[warp - 1] : (TO)0);
        if (i < N) out[i] = prefix + x - v;
        carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) out[N] = carry;
}

// Transposition, step 1: every entry p of source segment r is dropped into destination segment t = src_idx[p]
// at an arbitrary free slot, tagged with (r, p).  One warp per source segment.
static __global__ void k_ing_scatter(int nsrc, const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src_idx,
                                     const int32_t* __restrict__ dst_ptr, int32_t* __restrict__ cursor,
                                     int32_t* __restrict__ tkey, int32_t* __restrict__ tpos) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w0; r < nsrc; r += nw) {
        const int p1 = src_ptr[r + 1];
        for (int p = src_ptr[r] + lane; p < p1; p += 32) {
            const int t = src_idx[p];
            const int q = dst_ptr[t] + atomicAdd(&cursor[t], 1);
            tkey[q] = (int32_t)r;
            tpos[q] = p;
        }
    }
}

// Transposition, step 2: inside every destination segment the tags are ranked by r (unique inside a segment:
// the source indices are strictly ascending), which makes the result independent of the order in which the
// atomics of step 1 were served.  One CTA per destination segment.
static __global__ void __launch_bounds__(128) k_ing_place(int ndst, const int32_t* __restrict__ dst_ptr,
                                                          const int32_t* __restrict__ tkey,
                                                          const int32_t* __restrict__ tpos,
                                                          int32_t* __restrict__ dst_idx, int32_t* __restrict__ perm) {
    __shared__ int32_t keys[ING_PLACE_CAP];
    for (int t = blockIdx.x; t < ndst; t += gridDim.x) {
        const int q0 = dst_ptr[t], len = dst_ptr[t + 1] - q0;
        const bool staged = len <= ING_PLACE_CAP;
        if (staged)
            for (int e = threadIdx.x; e < len; e += blockDim.x) keys[e] = tkey[q0 + e];
        __syncthreads();
        for (int e = threadIdx.x; e < len; e += blockDim.x) {
            const int32_t k = staged ? keys[e] : tkey[q0 + e];
            int rank = 0;
            if (staged) {
                for (int f = 0; f < len; ++f) rank += (keys[f] < k) ? 1 : 0;
            } else {
                for (int f = 0; f < len; ++f) rank += (tkey[q0 + f] < k) ? 1 : 0;
            }
            dst_idx[q0 + rank] = k;
            perm[q0 + rank] = tpos[q0 + e];
        }
        __syncthreads();
    }
}

// dst[q] = src[perm[q]]: values of the derived orientation
static __global__ void k_ing_gather(int64_t nnz, const double* __restrict__ src, const int32_t* __restrict__ perm,
                                    double* __restrict__ dst) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x)
        dst[q] = src[perm[q]];
}

// Symbolic product, rows of M.  One CTA per row i (grid-stride).  cnt[j] (shared, j <= i) = number of columns k
// rows i and j of A share = number of terms of M_ij, found by walking the columns of row i (CSC view; the rows
// inside a column ascend, so the walk stops at the first j > i).
//   WRITE = false: rowent[i] = number of structural non-zeros M_ij (j <= i), rowterm[i] = their terms in total.
//   WRITE = true : with the exclusive scans entbase/termbase of those counts, emits for every entry in (i, j)
//                  order its position in the dense M, the start of its term list and (i, j) for the term pass.
template <bool WRITE>
static __global__ void __launch_bounds__(ING_NT) k_sym_rows(int m, const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colind,
                                                            const int32_t* __restrict__ colptr,
                                                            const int32_t* __restrict__ rowind, int32_t* rowent,
                                                            int64_t* rowterm, const int64_t* __restrict__ entbase,
                                                            const int64_t* __restrict__ termbase, int64_t ldm,
                                                            int64_t* __restrict__ out_idx,
                                                            int64_t* __restrict__ prod_ptr,
                                                            int32_t* __restrict__ ent_i, int32_t* __restrict__ ent_j) {
    extern __shared__ int32_t cnt[];
    __shared__ long long wsum[ING_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = blockIdx.x; i < m; i += gridDim.x) {
        for (int j = tid; j <= i; j += ING_NT) cnt[j] = 0;
        __syncthreads();
        const int p1 = rowptr[i + 1];
        for (int p = rowptr[i] + tid; p < p1; p += ING_NT) {
            const int k = colind[p];
            const int q1 = colptr[k + 1];
            for (int q = colptr[k]; q < q1; ++q) {
                const int j = rowind[q];
                if (j > i) break;
                atomicAdd(&cnt[j], 1);
            }
        }
        __syncthreads();
        // packed (entries << 40 | terms): one 64-bit scan/reduction serves both counts
        long long run = 0;      // WRITE: entries and terms emitted so far in this row
        long long tot = 0;      // !WRITE: this thread's share
        for (int base = 0; base <= i; base += ING_NT) {
            const int j = base + tid;
            const int c = (j <= i) ? cnt[j] : 0;
            const long long v = (c > 0) ? ((1LL << 40) | (long long)c) : 0LL;
            if (!WRITE) {
                tot += v;
            } else {
                long long x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                if (lane == 31) wsum[warp] = x;
                __syncthreads();
                long long prefix = run, total = 0;
#pragma unroll
                for (int w = 0; w < ING_NT / 32; ++w) {
                    const long long s = wsum[w];
                    if (w < warp) prefix += s;
                    total += s;
                }
                if (c > 0) {
                    const long long ex = prefix + x - v;
                    const int64_t e = entbase[i] + (ex >> 40);
                    out_idx[e] = (int64_t)i * ldm + j;
                    prod_ptr[e] = termbase[i] + (ex & ((1LL << 40) - 1));
                    ent_i[e] = i;
                    ent_j[e] = j;
                }
                run += total;
                __syncthreads();
            }
        }
        if (!WRITE) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) wsum[warp] = tot;
            __syncthreads();
            if (tid == 0) {
                long long s = 0;
                for (int w = 0; w < ING_NT / 32; ++w) s += wsum[w];
                rowent[i] = (int32_t)(s >> 40);
                rowterm[i] = (int64_t)(s & ((1LL << 40) - 1));
            }
        }
        __syncthreads();
    }
}

// Symbolic product, term lists: entry (i, j) = the merge of the sorted rows i and j of A; every shared column k
// contributes (position of a_ik, position of a_jk), k ascending - the summation order of the numeric phase.
static __global__ void k_sym_terms(int64_t nent, const int32_t* __restrict__ ent_i, const int32_t* __restrict__ ent_j,
                                   const int64_t* __restrict__ prod_ptr, const int32_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ colind, int32_t* __restrict__ pa,
                                   int32_t* __restrict__ pb) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nent; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = ent_i[e], j = ent_j[e];
        int a = rowptr[i], b = rowptr[j];
        const int a1 = rowptr[i + 1], b1 = rowptr[j + 1];
        int64_t t = prod_ptr[e];
        const int64_t t1 = prod_ptr[e + 1];
        while (a < a1 && b < b1 && t < t1) {
            const int ka = colind[a], kb = colind[b];
            if (ka == kb) {
                pa[t] = a;
                pb[t] = b;
                ++t; ++a; ++b;
            } else if (ka < kb) {
                ++a;
            } else {
                ++b;
            }
        }
    }
}

static __global__ void k_ing_set_last(int64_t* prod_ptr, int64_t nent, int64_t nterms) { prod_ptr[nent] = nterms; }
#endif  // __CUDACC__

// ---------------------------------------------------------------- host side: cache + build
struct PatternCache {
    std::mutex mu;
    std::list<std::shared_ptr<DevPattern>> items;    // most recently used first
    int64_t hits = 0, misses = 0;
    size_t max_items = 64;
    size_t max_bytes = (size_t)8 << 30;
};
inline PatternCache& pattern_cache() {
    static PatternCache c;
    return c;
}
// bit 0: symbolic phase on the host (legacy path, kept for cross-checks and for m beyond the shared-memory marker)
// bit 1: bypass the cache
inline std::atomic<int>& ingest_mode() {
    static std::atomic<int> m{0};
    return m;
}

inline uint64_t ing_hash(uint64_t h, const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001b3ULL;
        h ^= h >> 29;
    }
    for (; i < bytes; ++i) h = (h ^ p[i]) * 0x100000001b3ULL;
    return h;
}

inline void pattern_cache_clear() {
    PatternCache& C = pattern_cache();
    std::lock_guard<std::mutex> g(C.mu);
    C.items.clear();
}

#ifdef __CUDACC__
// Builds the structure for (ptr, idx) = the caller's orientation (from_csc: columns of A, else rows).
// On return pat is complete and the stream has been synchronised.
inline int build_pattern(DevPattern& pat, const int32_t* h_ptr, const int32_t* h_idx, cudaStream_t st) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int m = pat.m, n = pat.n;
    const int64_t nnz = pat.nnz;
    const int nsrc = pat.from_csc ? n : m, ndst = pat.from_csc ? m : n;
    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    auto al = [](size_t v) { return (v + 3) & ~(size_t)3; };      // keep every array 16-byte aligned
    const size_t isz = al(m + 1) + al(nz) + al(n + 1) + al(nz) + al(nz);
    IPM_CUDA_OK(cudaMalloc(&pat.islab, isz * sizeof(int32_t)));
    pat.rowptr = pat.islab;
    pat.colind = pat.rowptr + al(m + 1);
    pat.t_rowptr = pat.colind + al(nz);
    pat.t_colind = pat.t_rowptr + al(n + 1);
    pat.perm = pat.t_colind + al(nz);
    int32_t* src_ptr = pat.from_csc ? pat.t_rowptr : pat.rowptr;
    int32_t* src_idx = pat.from_csc ? pat.t_colind : pat.colind;
    int32_t* dst_ptr = pat.from_csc ? pat.rowptr : pat.t_rowptr;
    int32_t* dst_idx = pat.from_csc ? pat.colind : pat.t_colind;
    IPM_CUDA_OK(cudaMemcpyAsync(src_ptr, h_ptr, (size_t)(nsrc + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    IPM_CUDA_OK(cudaMemcpyAsync(src_idx, h_idx, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));

    // scratch: cnt/cursor [ndst] | tkey [nnz] | tpos [nnz] | rowent [m]  (int32), then int64: rowterm [m] |
    // entbase [m+1] | termbase [m+1]
    const size_t s32 = al(ndst) + al(nz) + al(nz) + al(m);
    const size_t s64 = (size_t)m + 2 * (size_t)(m + 1);
    int32_t* scratch = nullptr;
    IPM_CUDA_OK(cudaMalloc(&scratch, s32 * sizeof(int32_t) + s64 * sizeof(int64_t)));
    struct Guard { int32_t* p; ~Guard() { if (p) cudaFree(p); } } guard{scratch};
    int32_t* cnt = scratch;
    int32_t* tkey = cnt + al(ndst);
    int32_t* tpos = tkey + al(nz);
    int32_t* rowent = tpos + al(nz);
    int64_t* rowterm = reinterpret_cast<int64_t*>(scratch + s32);
    int64_t* entbase = rowterm + m;
    int64_t* termbase = entbase + (m + 1);

    const int gnnz = std::max(1, std::min<int>(ceil_div(nnz, ING_NT), 8 * kNumSMs));
    // ---- the other orientation
    IPM_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)ndst * sizeof(int32_t), st));
    k_ing_count<<<gnnz, ING_NT, 0, st>>>(nnz, src_idx, cnt);
    k_ing_scan<int32_t, int32_t><<<1, ING_SCAN_NT, 0, st>>>(cnt, dst_ptr, ndst);
    IPM_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)ndst * sizeof(int32_t), st));
    k_ing_scatter<<<std::max(1, std::min<int>(ceil_div((int64_t)nsrc * 32, ING_NT), 8 * kNumSMs)), ING_NT, 0, st>>>(
        nsrc, src_ptr, src_idx, dst_ptr, cnt, tkey, tpos);
    k_ing_place<<<std::max(1, std::min(ndst, 16 * kNumSMs)), 128, 0, st>>>(ndst, dst_ptr, tkey, tpos, dst_idx, pat.perm);
    count_launch(4);
    IPM_TRY(launch_check());

    const bool host_symbolic = (ingest_mode().load() & 1) || (size_t)m * sizeof(int32_t) > ING_SYM_MAX_SMEM;
    if (host_symbolic) {
        // legacy: pattern on the host from the CSR arrays (downloaded when the caller gave CSC)
        std::vector<int32_t> rp, ci;
        const int32_t *hrp = h_ptr, *hci = h_idx;
        if (pat.from_csc) {
            rp.resize(m + 1);
            ci.resize(nz);
            IPM_CUDA_OK(cudaMemcpyAsync(rp.data(), pat.rowptr, (size_t)(m + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            IPM_CUDA_OK(cudaMemcpyAsync(ci.data(), pat.colind, (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            IPM_CUDA_OK(cudaStreamSynchronize(st));
            hrp = rp.data();
            hci = ci.data();
        }
        SpgemmPattern hp;
        spgemm_symbolic(m, n, hrp, hci, pat.ldm, hp);
        pat.nent = (int64_t)hp.out_idx.size();
        pat.nterms = (int64_t)hp.pa.size();
        const size_t ne = (size_t)std::max<int64_t>(pat.nent, 1), nt = (size_t)std::max<int64_t>(pat.nterms, 1);
        const size_t psz = ne + (ne + 1) + (nt + 1) / 2 + (nt + 1) / 2;
        IPM_CUDA_OK(cudaMalloc(&pat.pslab, psz * sizeof(int64_t)));
        pat.out_idx = pat.pslab;
        pat.prod_ptr = pat.out_idx + ne;
        pat.pa = reinterpret_cast<int32_t*>(pat.prod_ptr + ne + 1);
        pat.pb = pat.pa + 2 * ((nt + 1) / 2);
        IPM_CUDA_OK(cudaMemcpyAsync(pat.out_idx, hp.out_idx.data(), hp.out_idx.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        IPM_CUDA_OK(cudaMemcpyAsync(pat.prod_ptr, hp.prod_ptr.data(), hp.prod_ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        IPM_CUDA_OK(cudaMemcpyAsync(pat.pa, hp.pa.data(), hp.pa.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        IPM_CUDA_OK(cudaMemcpyAsync(pat.pb, hp.pb.data(), hp.pb.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        IPM_CUDA_OK(cudaStreamSynchronize(st));
        pat.bytes = isz * sizeof(int32_t) + psz * sizeof(int64_t);
        pat.device_built = false;
    } else {
        // ---- symbolic product on the device
        IPM_TRY(ensure_dyn_smem(k_sym_rows<false>, ING_SYM_MAX_SMEM));
        IPM_TRY(ensure_dyn_smem(k_sym_rows<true>, ING_SYM_MAX_SMEM));
        const size_t smem = (size_t)round_up((int64_t)m * sizeof(int32_t), 16);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)200 * 1024) / std::max<size_t>(smem, 1)));
        const int grows = std::max(1, std::min(m, per_sm * kNumSMs));
        k_sym_rows<false><<<grows, ING_NT, smem, st>>>(m, pat.rowptr, pat.colind, pat.t_rowptr, pat.t_colind, rowent,
                                                       rowterm, nullptr, nullptr, pat.ldm, nullptr, nullptr, nullptr, nullptr);
        k_ing_scan<int32_t, int64_t><<<1, ING_SCAN_NT, 0, st>>>(rowent, entbase, m);
        k_ing_scan<int64_t, int64_t><<<1, ING_SCAN_NT, 0, st>>>(rowterm, termbase, m);
        count_launch(3);
        IPM_TRY(launch_check());
        int64_t totals[2] = {0, 0};
        IPM_CUDA_OK(cudaMemcpyAsync(&totals[0], entbase + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        IPM_CUDA_OK(cudaMemcpyAsync(&totals[1], termbase + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        IPM_CUDA_OK(cudaStreamSynchronize(st));
        pat.nent = totals[0];
        pat.nterms = totals[1];
        const size_t ne = (size_t)std::max<int64_t>(pat.nent, 1), nt = (size_t)std::max<int64_t>(pat.nterms, 1);
        // out_idx [ne] | prod_ptr [ne+1] | pa [nt] | pb [nt] | ent_i [ne] | ent_j [ne]   (int32 parts rounded to int64)
        const size_t psz = ne + (ne + 1) + (nt + 1) / 2 + (nt + 1) / 2 + (ne + 1) / 2 + (ne + 1) / 2;
        IPM_CUDA_OK(cudaMalloc(&pat.pslab, psz * sizeof(int64_t)));
        pat.out_idx = pat.pslab;
        pat.prod_ptr = pat.out_idx + ne;
        pat.pa = reinterpret_cast<int32_t*>(pat.prod_ptr + ne + 1);
        pat.pb = pat.pa + 2 * ((nt + 1) / 2);
        int32_t* ent_i = pat.pb + 2 * ((nt + 1) / 2);
        int32_t* ent_j = ent_i + 2 * ((ne + 1) / 2);
        k_sym_rows<true><<<grows, ING_NT, smem, st>>>(m, pat.rowptr, pat.colind, pat.t_rowptr, pat.t_colind, nullptr,
                                                      nullptr, entbase, termbase, pat.ldm, pat.out_idx, pat.prod_ptr,
                                                      ent_i, ent_j);
        k_ing_set_last<<<1, 1, 0, st>>>(pat.prod_ptr, pat.nent, pat.nterms);
        k_sym_terms<<<std::max(1, std::min<int>(ceil_div(pat.nent, 128), 16 * kNumSMs)), 128, 0, st>>>(
            pat.nent, ent_i, ent_j, pat.prod_ptr, pat.rowptr, pat.colind, pat.pa, pat.pb);
        count_launch(3);
        IPM_TRY(launch_check());
        IPM_CUDA_OK(cudaStreamSynchronize(st));
        pat.bytes = isz * sizeof(int32_t) + psz * sizeof(int64_t);
        pat.device_built = true;
    }
    pat.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return IPM_OK;
}

// Looks the structure up in the cache or builds it.  `hit` reports which.
inline int acquire_pattern(int dev, int m, int n, int64_t nnz, int64_t ldm, bool from_csc, const int32_t* h_ptr,
                           const int32_t* h_idx, cudaStream_t st, std::shared_ptr<DevPattern>& out, bool& hit) {
    PatternCache& C = pattern_cache();
    const int nsrc = from_csc ? n : m;
    const bool use_cache = !(ingest_mode().load() & 2);
    uint64_t h = 0xcbf29ce484222325ULL;
    const int64_t head[6] = {dev, m, n, nnz, ldm, from_csc ? 1 : 0};
    h = ing_hash(h, head, sizeof(head));
    h = ing_hash(h, h_ptr, (size_t)(nsrc + 1) * sizeof(int32_t));
    h = ing_hash(h, h_idx, (size_t)nnz * sizeof(int32_t));
    hit = false;
    if (use_cache) {
        std::lock_guard<std::mutex> g(C.mu);
        for (auto it = C.items.begin(); it != C.items.end(); ++it) {
            DevPattern& p = **it;
            if (p.hash == h && p.dev == dev && p.m == m && p.n == n && p.nnz == nnz && p.from_csc == from_csc &&
                std::memcmp(p.key_ptr.data(), h_ptr, (size_t)(nsrc + 1) * sizeof(int32_t)) == 0 &&
                (nnz == 0 || std::memcmp(p.key_idx.data(), h_idx, (size_t)nnz * sizeof(int32_t)) == 0)) {
                out = *it;
                C.items.splice(C.items.begin(), C.items, it);
                C.hits++;
                hit = true;
                return IPM_OK;
            }
        }
        C.misses++;
    }
    auto pat = std::make_shared<DevPattern>();
    pat->dev = dev; pat->m = m; pat->n = n; pat->nnz = nnz; pat->ldm = ldm; pat->from_csc = from_csc; pat->hash = h;
    IPM_TRY(build_pattern(*pat, h_ptr, h_idx, st));
    if (use_cache) {
        pat->key_ptr.assign(h_ptr, h_ptr + nsrc + 1);
        pat->key_idx.assign(h_idx, h_idx + nnz);
        std::lock_guard<std::mutex> g(C.mu);
        C.items.push_front(pat);
        size_t total = 0;
        for (auto& p : C.items) total += p->bytes;
        while (C.items.size() > 1 && (C.items.size() > C.max_items || total > C.max_bytes)) {
            total -= C.items.back()->bytes;
            C.items.pop_back();
        }
    }
    out = pat;
    return IPM_OK;
}
#endif  // __CUDACC__

}  // namespace ipm
