// Hybrid sparse -> dense rerank on candidates only.
//
// The reference names a "hybrid" retriever (rag_system/configs/ms_marco_paper_results.yaml:108-124:
// sparse bm25_custom + dense encoder, sparse_weight / dense_weight) but does not implement one; BASELINE
// configuration 5 describes the step as "top-100 rerank of BM25 candidates" with the INT8 vectors.  This
// file is that step: for every (query, candidate) pair the INT8 similarity of quantized_dot_product_batch
// (rag_system/core/retriever_registry.py:90-117: exact integer dot, f32((f64(dot) * f64(qs)) * f64(ds)))
// is evaluated on the candidate's vector only -- a gather of k_in x dim bytes per query instead of a scan of
// the corpus -- and combined with the sparse score as
//     hybrid = f32(f64(sparse_weight) * f64(sparse) + f64(dense_weight) * f64(dense))
// (both products and the sum rounded to nearest in f64, one rounding to f32), then the k_out best per query
// are selected under the library's ranking rule (score descending, document index ascending).
#include "common.cuh"

namespace b2r {

constexpr int RR_THREADS = 256;  // 8 warps, one candidate per warp and step

// grid = (ceil(k_in / 64), n_q): a CTA stages its query's vector in shared memory and its warps walk 64
// candidates; a warp reads the candidate's row with 16-byte loads (coalesced: consecutive lanes, consecutive
// 16 B), 4 x dp4a per load, and reduces the partial dots with shuffles.
__global__ void __launch_bounds__(RR_THREADS)
rerank_kernel(const int64_t *__restrict__ cand_idx, const float *__restrict__ cand_sparse, int k_in,
              const int8_t *__restrict__ q8, const float *__restrict__ q_scale, const int8_t *__restrict__ d8,
              const float *__restrict__ d_scale, int64_t n_docs, int dim, int64_t doc_id_base, double w_sparse,
              double w_dense, float *__restrict__ dense_out, uint64_t *__restrict__ keys) {
    extern __shared__ __align__(16) int8_t q_s[];  // [dim rounded up to 16]
    const int q = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dim16 = (dim + 15) & ~15;
    for (int i = threadIdx.x; i < dim16; i += RR_THREADS) q_s[i] = i < dim ? q8[(int64_t)q * dim + i] : (int8_t)0;
    __syncthreads();
    const double qs = (double)q_scale[q];
    const bool vec = (dim & 15) == 0 && (reinterpret_cast<uintptr_t>(d8) & 15) == 0;
    const int c_end = min(k_in, (int)(blockIdx.x + 1) * 64);
    for (int c = blockIdx.x * 64 + warp; c < c_end; c += RR_THREADS / 32) {
        const int64_t slot = (int64_t)q * k_in + c;
        const int64_t gid = cand_idx[slot];
        const int64_t d = gid - doc_id_base;
        if (gid < 0 || d < 0 || d >= n_docs) {  // "no candidate" (warp-uniform)
            if (lane == 0) {
                keys[slot] = 0;
                if (dense_out) dense_out[slot] = __int_as_float(0xff800000);
            }
            continue;
        }
        const int8_t *row = d8 + d * (int64_t)dim;
        int dot = 0;
        if (vec) {
            for (int o = lane * 16; o < dim; o += 512) {
                const int4 a = __ldg(reinterpret_cast<const int4 *>(row + o));
                const int4 b = *reinterpret_cast<const int4 *>(q_s + o);
                dot = __dp4a(a.x, b.x, dot);
                dot = __dp4a(a.y, b.y, dot);
                dot = __dp4a(a.z, b.z, dot);
                dot = __dp4a(a.w, b.w, dot);
            }
        } else {
            for (int o = lane; o < dim; o += 32) dot += (int)row[o] * (int)q_s[o];
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, s);
        if (lane == 0) {
            const float dense = __double2float_rn(__dmul_rn(__dmul_rn((double)dot, qs), (double)d_scale[d]));
            float score = dense;
            if (cand_sparse != nullptr)
                score = __double2float_rn(__dadd_rn(__dmul_rn(w_sparse, (double)cand_sparse[slot]),
                                                    __dmul_rn(w_dense, (double)dense)));
            if (dense_out) dense_out[slot] = dense;
            keys[slot] = make_key(ord_f32(score), (uint32_t)gid);
        }
    }
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_int8_rerank_workspace(int32_t n_q, int32_t k_in, int32_t k_out, size_t *bytes) {
    B2R_CHECK_ARG(bytes && n_q >= 0 && k_in >= 1 && k_out >= 1 && k_out <= B2R_TOPK_MAX_FAST,
                  "b2r_int8_rerank_workspace: bad arguments");
    const size_t nq = n_q > 0 ? n_q : 1;
    *bytes = align_up(nq * k_in * 8, 256) + align_up(nq * k_out * 8, 256) + topk_keys_ws_bytes(nq, k_in, k_out) + 512;
    return B2R_OK;
}

extern "C" int b2r_int8_rerank(const int64_t *cand_idx, const float *cand_sparse, int32_t n_q, int32_t k_in,
                               const int8_t *q8, const float *q_scale, const int8_t *d8, const float *d_scale,
                               int64_t n_docs, int32_t dim, int64_t doc_id_base, double sparse_weight,
                               double dense_weight, int32_t k_out, float *dense_out, int64_t *idx_out, float *val_out,
                               void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(cand_idx && q8 && q_scale && d8 && d_scale && n_q >= 0 && k_in >= 1 && n_docs >= 1 && dim >= 1,
                  "b2r_int8_rerank: bad arguments");
    B2R_CHECK_ARG(k_out >= 1 && k_out <= B2R_TOPK_MAX_FAST && k_out <= k_in, "b2r_int8_rerank: k_out=%d outside [1, min(k_in, %d)]",
                  k_out, B2R_TOPK_MAX_FAST);
    B2R_CHECK_ARG((int64_t)dim * 127 * 127 < 0x7FFFFFFFll && dim <= 32768, "b2r_int8_rerank: dim too large");
    B2R_CHECK_ARG(doc_id_base >= 0 && doc_id_base + n_docs < 0xFFFFFFFFll, "b2r_int8_rerank: doc index range");
    if (n_q == 0) return B2R_OK;
    size_t need = 0;
    b2r_int8_rerank_workspace(n_q, k_in, k_out, &need);
    if (workspace_bytes < need) {
        set_error("b2r_int8_rerank: workspace too small (%zu < %zu)", workspace_bytes, need);
        return B2R_ERR_WORKSPACE;
    }
    char *wp = static_cast<char *>(workspace);
    uint64_t *keys = reinterpret_cast<uint64_t *>(wp);
    wp += align_up((size_t)n_q * k_in * 8, 256);
    uint64_t *best = reinterpret_cast<uint64_t *>(wp);
    wp += align_up((size_t)n_q * k_out * 8, 256);
    const size_t left = workspace_bytes - (size_t)(wp - static_cast<char *>(workspace));
    dim3 grid((unsigned)((k_in + 63) / 64), (unsigned)n_q);
    const size_t smem = (size_t)((dim + 15) & ~15);
    rerank_kernel<<<grid, RR_THREADS, smem, st>>>(cand_idx, cand_sparse, k_in, q8, q_scale, d8, d_scale, n_docs, dim,
                                                  doc_id_base, sparse_weight, dense_weight, dense_out,
                                                  keys);
    B2R_LAUNCH_CHECK();
    int rc = topk_keys_rows(keys, n_q, k_in, k_in, k_in, 0, k_out, best, wp, left, st);
    if (rc) return rc;
    return decode_keys(best, (int64_t)n_q * k_out, idx_out, val_out, nullptr, 0, k_out, 0, st);
}
