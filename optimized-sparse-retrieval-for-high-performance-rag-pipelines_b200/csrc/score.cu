// Batched term-at-a-time scoring over the term-major index
// (reference: simd_bm25_score, rag_system/core/retrieval.py:41-76, and simd_tfidf_score,
//  rag_system/pipeline/evaluate_rag_pipeline.py:95-121 -- both doc-major O(nnz) scans per query).
//
// One CTA = one (query, doc tile), one warp = one sub-tile.  The f64 accumulators live in shared
// memory; a warp applies the query's terms in ascending term id to its own sub-tile, which reproduces
// the reference's summation order (CSR rows are sorted by term id) with no atomics and no CTA barrier:
// a document occurs at most once in a term's posting list.  The grid is (queries, tiles) with the
// query index fastest, so CTAs resident at the same time work on the same doc tile and share its
// posting blocks through L2; HBM sees each posting once per batch.
#include "common.cuh"

namespace b2r {

// posting streams are read once per CTA: keep them out of L1 (the accumulators own the SM's L1/smem pipe)
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

constexpr int SC_THREADS = 32 * B2R_SUBTILES;  // one warp per sub-tile of the CTA's doc tile

// acc[rel] += contribution of one posting (reference order of operations, no FMA contraction)
template <int KIND>
__device__ __forceinline__ void apply_posting(double *acc_w, uint32_t rel, double u_or_w, float w_idf, float w_q,
                                              double w_idf64, double w_q64) {
    if (KIND == B2R_KIND_BM25) {
        acc_w[rel] = __dadd_rn(acc_w[rel], __dmul_rn(__dmul_rn(w_idf64, u_or_w), w_q64));
    } else {
        // reference (fastmath) evaluates (tf * qtf) * idf in f32, then widens: see oracle/np_oracle.py
        float c = __fmul_rn(__fmul_rn((float)u_or_w, w_q), w_idf);
        acc_w[rel] = __dadd_rn(acc_w[rel], (double)c);
    }
}

template <int KIND>
__device__ __forceinline__ double load_val(const void *post_val, uint32_t p) {
    if (KIND == B2R_KIND_BM25) return ld_stream_f64(static_cast<const double *>(post_val) + p);
    return (double)ld_stream_f32(static_cast<const float *>(post_val) + p);
}

// One CTA = one (query, doc tile); one WARP = one sub-tile of tile_docs/8 docs whose f64 accumulators
// it alone touches.  A warp applies the query's terms in ascending term id to its own sub-tile, so the
// only synchronisation is __syncwarp: no CTA barrier, no atomics, no load imbalance between warps
// (a dense term's postings are split by sub-tile through dense_ptr; a sparse term's small block is
// scanned by every warp, each keeping the postings that fall in its range).
template <int KIND>
__global__ void __launch_bounds__(SC_THREADS)
score_tiles_kernel(const uint32_t *__restrict__ post_doc, const void *__restrict__ post_val,
                   const uint32_t *__restrict__ blk_ptr, const int32_t *__restrict__ dense_id,
                   const uint32_t *__restrict__ dense_ptr, int n_tiles, int tile_docs,
                   const int32_t *__restrict__ q_ptr, const int32_t *__restrict__ q_terms,
                   const float *__restrict__ q_weights, const float *__restrict__ idf, int q0,
                   float *__restrict__ scores, int64_t scores_stride) {
    extern __shared__ double acc[];  // [tile_docs]
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = q0 + blockIdx.x;
    const int tile = blockIdx.y;
    const int sub = tile_docs / B2R_SUBTILES;
    const uint32_t my_doc0 = (uint32_t)tile * (uint32_t)tile_docs + (uint32_t)w * (uint32_t)sub;
    double *acc_w = acc + w * sub;
    const size_t dense_row = (size_t)n_tiles * B2R_SUBTILES + 1;
    const size_t my_sub = (size_t)tile * B2R_SUBTILES + w;

    for (int i = lane * 2; i < sub; i += 64) *reinterpret_cast<double2 *>(acc_w + i) = make_double2(0.0, 0.0);
    __syncwarp();

    const int qs = q_ptr[q], qe = q_ptr[q + 1];
    for (int j0 = qs; j0 < qe; j0 += 32) {
        const int nt = min(32, qe - j0);
        // lane j stages term j0+j: posting range of this warp (dense) or of the whole tile block (sparse)
        uint32_t my_beg = 0, my_end = 0;
        float my_idf = 0.f, my_qw = 0.f;
        int my_dense = 0;
        if (lane < nt) {
            const int t = q_terms[j0 + lane];
            my_qw = q_weights[j0 + lane];
            my_idf = idf[t];
            const int32_t did = dense_id[t];
            if (did >= 0) {
                const uint32_t *row = dense_ptr + (size_t)did * dense_row + my_sub;
                my_beg = row[0];
                my_end = row[1];
                my_dense = 1;
            } else {
                const size_t e = (size_t)t * n_tiles + tile;
                my_beg = blk_ptr[e];
                my_end = blk_ptr[e + 1];
            }
        }
        for (int j = 0; j < nt; ++j) {
            const uint32_t beg = __shfl_sync(full, my_beg, j), end = __shfl_sync(full, my_end, j);
            if (beg == end) continue;  // warp-uniform
            const int dense = __shfl_sync(full, my_dense, j);
            const float w_idf = __shfl_sync(full, my_idf, j), w_q = __shfl_sync(full, my_qw, j);
            const double w_idf64 = (double)w_idf, w_q64 = (double)w_q;
            if (dense) {
                uint32_t p = beg + lane;
                for (; p + 96 < end; p += 128) {  // 4 independent postings in flight per lane
                    uint32_t d0 = ld_stream_u32(post_doc + p), d1 = ld_stream_u32(post_doc + p + 32);
                    uint32_t d2 = ld_stream_u32(post_doc + p + 64), d3 = ld_stream_u32(post_doc + p + 96);
                    double u0 = load_val<KIND>(post_val, p), u1 = load_val<KIND>(post_val, p + 32);
                    double u2 = load_val<KIND>(post_val, p + 64), u3 = load_val<KIND>(post_val, p + 96);
                    apply_posting<KIND>(acc_w, d0 - my_doc0, u0, w_idf, w_q, w_idf64, w_q64);
                    apply_posting<KIND>(acc_w, d1 - my_doc0, u1, w_idf, w_q, w_idf64, w_q64);
                    apply_posting<KIND>(acc_w, d2 - my_doc0, u2, w_idf, w_q, w_idf64, w_q64);
                    apply_posting<KIND>(acc_w, d3 - my_doc0, u3, w_idf, w_q, w_idf64, w_q64);
                }
                for (; p < end; p += 32) {
                    uint32_t d = ld_stream_u32(post_doc + p);
                    double u = load_val<KIND>(post_val, p);
                    apply_posting<KIND>(acc_w, d - my_doc0, u, w_idf, w_q, w_idf64, w_q64);
                }
            } else {
                for (uint32_t p = beg + lane; p < end; p += 32) {
                    const uint32_t rel = __ldg(post_doc + p) - my_doc0;  // blocks are shared by the 8 warps: keep in L1
                    if (rel < (uint32_t)sub) {
                        double u = load_val<KIND>(post_val, p);
                        apply_posting<KIND>(acc_w, rel, u, w_idf, w_q, w_idf64, w_q64);
                    }
                }
            }
            __syncwarp();
        }
    }

    float *out = scores + (int64_t)blockIdx.x * scores_stride + my_doc0;
    for (int i = lane * 2; i < sub; i += 64) {
        double2 a = *reinterpret_cast<const double2 *>(acc_w + i);
        *reinterpret_cast<float2 *>(out + i) = make_float2(__double2float_rn(a.x), __double2float_rn(a.y));
    }
}

static int launch_score(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                        const float *idf, int q0, int nq, float *scores, int64_t stride, cudaStream_t st) {
    if (nq == 0) return B2R_OK;
    const size_t smem = (size_t)ix->tile_docs * sizeof(double);
    dim3 grid((unsigned)nq, (unsigned)ix->n_tiles);
    if (ix->kind == B2R_KIND_BM25) {
        B2R_CUDA(cudaFuncSetAttribute(score_tiles_kernel<B2R_KIND_BM25>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        score_tiles_kernel<B2R_KIND_BM25><<<grid, SC_THREADS, smem, st>>>(
            ix->post_doc, ix->post_val, ix->blk_ptr, ix->dense_id, ix->dense_ptr, ix->n_tiles, ix->tile_docs, q_ptr,
            q_terms, q_weights, idf, q0, scores, stride);
    } else {
        B2R_CUDA(cudaFuncSetAttribute(score_tiles_kernel<B2R_KIND_IMPACT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        score_tiles_kernel<B2R_KIND_IMPACT><<<grid, SC_THREADS, smem, st>>>(
            ix->post_doc, ix->post_val, ix->blk_ptr, ix->dense_id, ix->dense_ptr, ix->n_tiles, ix->tile_docs, q_ptr,
            q_terms, q_weights, idf, q0, scores, stride);
    }
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

static int check_index(const b2r_index *ix) {
    B2R_CHECK_ARG(ix && ix->post_doc && ix->post_val && ix->blk_ptr && ix->dense_id && ix->dense_ptr,
                  "search: index not built");
    B2R_CHECK_ARG(ix->tile_docs >= 256 && (ix->tile_docs & (ix->tile_docs - 1)) == 0, "search: bad tile_docs");
    B2R_CHECK_ARG(ix->n_tiles == (ix->n_docs + ix->tile_docs - 1) / ix->tile_docs && ix->n_tiles <= 65535,
                  "search: bad n_tiles");
    B2R_CHECK_ARG(ix->doc_id_base >= 0 && ix->doc_id_base + ix->n_docs < 0xFFFFFFFFll,
                  "search: global doc index exceeds 2^32-2");
    return B2R_OK;
}

static int64_t padded_docs(const b2r_index *ix) { return (int64_t)ix->n_tiles * ix->tile_docs; }

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_search_workspace(const b2r_index *ix, int32_t n_queries, int32_t k, size_t *min_bytes,
                                    size_t *full_bytes) {
    int rc = check_index(ix);
    if (rc) return rc;
    B2R_CHECK_ARG(n_queries >= 0 && k >= 0 && k <= B2R_TOPK_MAX_FAST, "b2r_search_workspace: bad n_queries/k");
    const size_t row = (size_t)padded_docs(ix) * 4;
    const int kk = k > 0 ? k : 1;
    const size_t keys = align_up((size_t)(n_queries > 0 ? n_queries : 1) * kk * 8, 256);
    if (min_bytes) *min_bytes = align_up(row, 256) + topk_ws_bytes(1, ix->n_docs, kk) + keys + 512;
    if (full_bytes) {
        int64_t nq = n_queries > 0 ? n_queries : 1;
        *full_bytes = align_up(row * (size_t)nq, 256) + topk_ws_bytes(nq, ix->n_docs, kk) + keys + 512;
    }
    return B2R_OK;
}

extern "C" int b2r_search_batch(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms,
                                const float *q_weights, const float *idf, int32_t n_queries, int32_t k,
                                float *scores_out, int64_t scores_stride, uint64_t *keys_out, int64_t *idx_out,
                                float *val_out, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = check_index(ix);
    if (rc) return rc;
    B2R_CHECK_ARG(n_queries >= 0 && q_ptr && idf, "b2r_search_batch: null query buffers");
    B2R_CHECK_ARG(k >= 0 && k <= B2R_TOPK_MAX_FAST, "b2r_search_batch: k=%d outside [0,%d]", k, B2R_TOPK_MAX_FAST);
    B2R_CHECK_ARG(k > 0 || scores_out, "b2r_search_batch: nothing to compute (k == 0 and no scores_out)");
    B2R_CHECK_ARG(k > 0 || (!keys_out && !idx_out && !val_out), "b2r_search_batch: k == 0 with top-k outputs");
    if (n_queries == 0) return B2R_OK;
    const int64_t pad = padded_docs(ix);
    if (scores_out)
        B2R_CHECK_ARG(scores_stride >= pad && (scores_stride & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(scores_out) & 15) == 0,
                      "b2r_search_batch: scores_out needs stride >= %lld, stride %% 4 == 0, 16-byte alignment",
                      (long long)pad);

    char *wp = static_cast<char *>(workspace);
    size_t left = workspace_bytes;
    auto carve = [&](size_t bytes) -> void * {
        bytes = align_up(bytes, 256);
        if (bytes > left) return nullptr;
        void *p = wp;
        wp += bytes;
        left -= bytes;
        return p;
    };

    uint64_t *keys = keys_out;
    if (k > 0 && !keys) {
        keys = static_cast<uint64_t *>(carve((size_t)n_queries * k * 8));
        if (!keys) {
            set_error("b2r_search_batch: workspace too small for the key buffer");
            return B2R_ERR_WORKSPACE;
        }
    }

    if (scores_out) {
        rc = launch_score(ix, q_ptr, q_terms, q_weights, idf, 0, n_queries, scores_out, scores_stride, st);
        if (rc) return rc;
        if (k > 0) {
            rc = topk_scores_rows(scores_out, n_queries, ix->n_docs, scores_stride, k, ix->doc_id_base, keys, wp, left,
                                  st);
            if (rc) return rc;
        }
    } else {
        // score in query chunks sized by the workspace
        const size_t row_bytes = (size_t)pad * 4;
        int64_t qc = n_queries;
        while (qc > 1 && align_up(row_bytes * (size_t)qc, 256) + topk_ws_bytes(qc, ix->n_docs, k) > left)
            qc = (qc + 1) / 2;
        if (align_up(row_bytes * (size_t)qc, 256) + topk_ws_bytes(qc, ix->n_docs, k) > left) {
            set_error("b2r_search_batch: workspace too small (%zu bytes left, one query needs %zu)", left,
                      align_up(row_bytes, 256) + topk_ws_bytes(1, ix->n_docs, k));
            return B2R_ERR_WORKSPACE;
        }
        float *chunk = static_cast<float *>(carve(row_bytes * (size_t)qc));
        for (int64_t q0 = 0; q0 < n_queries; q0 += qc) {
            int nq = (int)((n_queries - q0) < qc ? (n_queries - q0) : qc);
            rc = launch_score(ix, q_ptr, q_terms, q_weights, idf, (int)q0, nq, chunk, pad, st);
            if (rc) return rc;
            rc = topk_scores_rows(chunk, nq, ix->n_docs, pad, k, ix->doc_id_base, keys + q0 * k, wp, left, st);
            if (rc) return rc;
        }
    }
    if (k > 0) return decode_keys(keys, (int64_t)n_queries * k, idx_out, val_out, nullptr, 0, k, 0, st);
    return B2R_OK;
}

extern "C" size_t b2r_search_host_extra_bytes(int32_t n_queries, int64_t n_query_terms, int32_t k) {
    size_t q = (size_t)(n_queries > 0 ? n_queries : 1);
    size_t t = (size_t)(n_query_terms > 0 ? n_query_terms : 1);
    size_t kk = (size_t)(k > 0 ? k : 1);
    return align_up((q + 1) * 4, 256) + 2 * align_up(t * 4, 256) + align_up(q * kk * 8, 256) * 2 +
           align_up(q * kk * 4, 256) + 256;
}

extern "C" int b2r_search_batch_host(const b2r_index *ix, const int32_t *q_ptr_h, const int32_t *q_terms_h,
                                     const float *q_weights_h, const float *idf, int32_t n_queries, int32_t k,
                                     uint64_t *keys_out_h, int64_t *idx_out_h, float *val_out_h, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(q_ptr_h && n_queries >= 0 && k >= 1, "b2r_search_batch_host: bad arguments");
    if (n_queries == 0) return B2R_OK;
    const int64_t n_terms = q_ptr_h[n_queries];
    B2R_CHECK_ARG(n_terms >= 0 && (n_terms == 0 || (q_terms_h && q_weights_h)), "b2r_search_batch_host: null terms");
    char *wp = static_cast<char *>(workspace);
    size_t left = workspace_bytes;
    auto carve = [&](size_t bytes) -> void * {
        bytes = align_up(bytes, 256);
        if (bytes > left) return nullptr;
        void *p = wp;
        wp += bytes;
        left -= bytes;
        return p;
    };
    int32_t *d_ptr = static_cast<int32_t *>(carve((size_t)(n_queries + 1) * 4));
    int32_t *d_terms = static_cast<int32_t *>(carve((size_t)(n_terms > 0 ? n_terms : 1) * 4));
    float *d_w = static_cast<float *>(carve((size_t)(n_terms > 0 ? n_terms : 1) * 4));
    uint64_t *d_keys = static_cast<uint64_t *>(carve((size_t)n_queries * k * 8));
    int64_t *d_idx = static_cast<int64_t *>(carve((size_t)n_queries * k * 8));
    float *d_val = static_cast<float *>(carve((size_t)n_queries * k * 4));
    if (!d_ptr || !d_terms || !d_w || !d_keys || !d_idx || !d_val) {
        set_error("b2r_search_batch_host: workspace too small for the staging buffers");
        return B2R_ERR_WORKSPACE;
    }
    B2R_CUDA(cudaMemcpyAsync(d_ptr, q_ptr_h, (size_t)(n_queries + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_terms > 0) {
        B2R_CUDA(cudaMemcpyAsync(d_terms, q_terms_h, (size_t)n_terms * 4, cudaMemcpyHostToDevice, st));
        B2R_CUDA(cudaMemcpyAsync(d_w, q_weights_h, (size_t)n_terms * 4, cudaMemcpyHostToDevice, st));
    }
    int rc = b2r_search_batch(ix, d_ptr, d_terms, d_w, idf, n_queries, k, nullptr, 0, d_keys, d_idx, d_val, wp, left,
                              stream);
    if (rc) return rc;
    if (keys_out_h)
        B2R_CUDA(cudaMemcpyAsync(keys_out_h, d_keys, (size_t)n_queries * k * 8, cudaMemcpyDeviceToHost, st));
    if (idx_out_h) B2R_CUDA(cudaMemcpyAsync(idx_out_h, d_idx, (size_t)n_queries * k * 8, cudaMemcpyDeviceToHost, st));
    if (val_out_h) B2R_CUDA(cudaMemcpyAsync(val_out_h, d_val, (size_t)n_queries * k * 4, cudaMemcpyDeviceToHost, st));
    B2R_CUDA(cudaStreamSynchronize(st));
    return B2R_OK;
}
