// Shared device/host helpers of libb200ret (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200ret.h"

namespace b2r {

void set_error(const char *fmt, ...);


#define B2R_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            b2r::set_error(__VA_ARGS__);    \
            return B2R_ERR_ARG;             \
        }                                   \
    } while (0)

#define B2R_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            b2r::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return B2R_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// every kernel launch of the library goes through this macro; the counter backs b2r_launch_count()
// (relaxed atomic: searches may be enqueued from several host threads)
extern unsigned long long g_launches;
#define B2R_LAUNCH_CHECK()                                                \
    do {                                                                  \
        __atomic_fetch_add(&b2r::g_launches, 1ull, __ATOMIC_RELAXED);     \
        B2R_CUDA(cudaGetLastError());                                     \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Candidate keys.  Ranking rule: f32 score descending (-0.0 == +0.0, NaN last), then global doc
// index ascending.  key = ordered_u32(score) << 32 | (0xFFFFFFFF - index); larger key ranks first;
// 0 is the "no candidate" sentinel (every real key is >= 1 because index < 2^32 - 1).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ord_f32(float f) {
    uint32_t u = __float_as_uint(f);
    if (f != f) return 0u;
    if (f == 0.0f) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
__device__ __forceinline__ uint64_t make_key(uint32_t o, uint32_t gid) {
    return ((uint64_t)o << 32) | (uint64_t)(0xFFFFFFFFu - gid);
}

__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// In-shared-memory bitonic sort, descending, of P (power of two) u64 keys by the whole CTA.
// Ends with a __syncthreads().
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t *arr, int P) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (P >> 1); i += THREADS) {
                int pos = 2 * i - (i & (stride - 1));
                uint64_t a = arr[pos], b = arr[pos + stride];
                bool desc = ((pos & size) == 0);
                if (desc ? (a < b) : (a > b)) {
                    arr[pos] = b;
                    arr[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
}

// internal entry points shared between translation units
struct TopkOpts {
    // chunk_shift > 0: a score row is a compacted list of 2^chunk_shift-element chunks taken every
    // chunk_stride documents: element j has index id_base + (j >> shift) * chunk_stride + (j & mask)
    int chunk_shift = 0;
    uint32_t chunk_stride = 0;
    // gate != nullptr: only rows with gate[row] > gate_cap are processed, the others are left untouched
    const int32_t *gate = nullptr;
    int32_t gate_cap = 0;
};
int topk_keys_rows(const uint64_t *keys_in, int64_t n_rows, int64_t n, int64_t row_stride, int64_t piece_len,
                   int64_t piece_stride, int32_t k, uint64_t *keys_out, void *ws, size_t ws_bytes,
                   cudaStream_t st, const TopkOpts &opts = TopkOpts());
int topk_scores_rows(const float *scores, int64_t n_rows, int64_t n, int64_t row_stride, int32_t k,
                     int64_t doc_id_base, uint64_t *keys_out, void *ws, size_t ws_bytes, cudaStream_t st,
                     const TopkOpts &opts = TopkOpts());
size_t topk_ws_bytes(int64_t n_rows, int64_t n, int32_t k);       // for topk_scores_rows
size_t topk_keys_ws_bytes(int64_t n_rows, int64_t n, int32_t k);  // for topk_keys_rows
// thr_out[row] = key with the score of the k-th largest of the n_groups f32 group maxima of the row and doc
// bits 0 (every document with at least that score beats it); 0 when fewer than k maxima are finite.
// lower = true: the maxima are f32 approximations within 2^-22 of the exact scores -- lower the threshold by
// 2^-20 relative so that it stays a valid bound for the exact scores (INT8 scan).
// positive_floor = true: a threshold score <= 0 is replaced by "strictly positive scores only" (the caller must
// then gate on short lists: topk_of_lists(min_cnt = k)).
// zero_a / zero_b (optional): int32[n_rows] counters the kernel resets to 0 for its row; zero_scalar (optional): one
// int32 reset by the first row's CTA (the number of marked rows, see topk_of_lists).
int kth_of_maxima(const float *maxima, int64_t n_rows, int64_t n_groups, int64_t row_stride, int32_t k, bool lower,
                  bool positive_floor, uint64_t *thr_out, cudaStream_t st, int32_t *zero_a = nullptr,
                  int32_t *zero_b = nullptr, int32_t *zero_scalar = nullptr);
// keys_out[row, 0..k) = the k best of the first min(cnt[row], cap) keys of lists[row, 0..cap) (0-padded); cap <= 4096.
// Rows with cnt[row] < min_cnt get cnt[row] = cap + 1, i.e. they are marked like overflowed rows for the fallback gate.
// idx_out / val_out (optional): the decoded form of the ranked keys (-1 / -inf for "no candidate"); keys_out may be null.
// marked (optional): int32[1 + n_rows]; every marked row (overflowed or short) appends its index: marked[1 + marked[0]++].
int topk_of_lists(const uint64_t *lists, int64_t n_rows, int cap, int32_t *cnt, int32_t k, int32_t min_cnt,
                  uint64_t *keys_out, cudaStream_t st, int64_t *idx_out = nullptr, float *val_out = nullptr,
                  int32_t *marked = nullptr);
// approximate pre-filter of the BM25 search path (score_approx.cu): drop-in replacements of the MAXIMA launch, the
// FUSED launch and topk_of_lists of the f64 fused path; the exhaustive fallback gate (cnt > cap) is shared
bool approx_usable(const b2r_index *ix, int k);
int approx_maxima(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                  const float *idf, int q0, int nq, int tile_step, int n_sample, float *maxima, int64_t maxima_stride,
                  cudaStream_t st);
int approx_fused(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                 const float *idf, int q0, int nq, const uint64_t *thr, uint64_t *cand, int32_t *cand_cnt, int cap,
                 cudaStream_t st);
int approx_select(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                  const float *idf, int q0, int nq, const uint64_t *thr, const uint64_t *cand, int32_t *cand_cnt,
                  int cap, int k, uint64_t *keys_out, int64_t *idx_out, float *val_out, int32_t *marked,
                  cudaStream_t st);
int decode_keys(const uint64_t *keys, int64_t n, int64_t *idx_out, float *val_out, const float *scores,
                int64_t row_stride, int32_t k, int64_t doc_id_base, cudaStream_t st);

}  // namespace b2r
