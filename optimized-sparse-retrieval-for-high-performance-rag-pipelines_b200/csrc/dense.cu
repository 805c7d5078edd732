// fp32 dense similarity + top-k (reference: RetrievalService.search_by_vector,
// rag_system/core/retrieval.py:402-436: np.dot(embedding_index, query_vector) through host BLAS, then
// fast_topk_selection and the min_score cut-off).  SURVEY.md section 8 row a8: not in any BASELINE
// configuration, HBM-bound gemv.  The embedding matrix f32[N, D] is resident in HBM; one warp owns a row at
// a time, keeps partial sums for up to 8 queries (staged in shared memory) and reduces them with shuffles:
// every row is read once per chunk of 8 queries, with 16-byte loads.
//
// Numerics: f32 products and f32 pairwise sums (32 lane-partials + a shuffle tree); BLAS does not specify its
// summation order either, so parity with the reference is by tolerance: |score - exact| <= 1e-5 * sum|a_i b_i|
// (tests/test_gpu_parity.py), the selection on the computed scores is exact under the library's ranking rule.
#include "common.cuh"

namespace b2r {

constexpr int DN_THREADS = 256;
constexpr int DN_QC = 8;  // queries per pass over the matrix

__global__ void __launch_bounds__(DN_THREADS)
f32_dot_kernel(const float *__restrict__ emb, int64_t n_rows, int dim, const float *__restrict__ queries, int q0,
               int nq, float *__restrict__ scores, int64_t scores_stride) {
    extern __shared__ __align__(16) float q_s[];  // [nq][dim4]
    const int dim4 = (dim + 3) & ~3;
    for (int i = threadIdx.x; i < nq * dim4; i += DN_THREADS) {
        const int q = i / dim4, d = i - q * dim4;
        q_s[i] = d < dim ? queries[(int64_t)(q0 + q) * dim + d] : 0.0f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * DN_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * DN_THREADS) >> 5;
    const bool vec = (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(emb) & 15) == 0;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float *row = emb + r * dim;
        float acc[DN_QC];
#pragma unroll
        for (int q = 0; q < DN_QC; ++q) acc[q] = 0.0f;
        if (vec) {
            for (int d = lane * 4; d < dim; d += 128) {
                const float4 a = ldg_stream_f4(row + d);
#pragma unroll
                for (int q = 0; q < DN_QC; ++q) {
                    if (q < nq) {
                        const float4 b = *reinterpret_cast<const float4 *>(q_s + q * dim4 + d);
                        acc[q] = __fadd_rn(acc[q], __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                                                             __fadd_rn(__fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w))));
                    }
                }
            }
        } else {
            for (int d = lane; d < dim; d += 32) {
                const float a = __ldg(row + d);
#pragma unroll
                for (int q = 0; q < DN_QC; ++q)
                    if (q < nq) acc[q] = __fadd_rn(acc[q], __fmul_rn(a, q_s[q * dim4 + d]));
            }
        }
#pragma unroll
        for (int q = 0; q < DN_QC; ++q) {
            if (q < nq) {
                float v = acc[q];
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, s));
                if (lane == 0) scores[(int64_t)(q0 + q) * scores_stride + r] = v;
            }
        }
    }
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_f32_dot_topk_workspace(int32_t n_q, int64_t n_rows, int32_t k, size_t *bytes) {
    B2R_CHECK_ARG(bytes && n_q >= 0 && n_rows >= 1 && k >= 0 && k <= B2R_TOPK_MAX_FAST,
                  "b2r_f32_dot_topk_workspace: bad arguments");
    const size_t nq = n_q > 0 ? n_q : 1;
    const size_t stride = align_up((size_t)n_rows, 4);
    *bytes = align_up(nq * stride * 4, 256) + align_up(nq * (k > 0 ? k : 1) * 8, 256) +
             topk_ws_bytes(nq, n_rows, k > 0 ? k : 1) + 512;
    return B2R_OK;
}

extern "C" int b2r_f32_dot_topk(const float *emb, int64_t n_rows, int32_t dim, const float *queries, int32_t n_q,
                                int32_t k, int64_t doc_id_base, float *scores_out, int64_t scores_stride,
                                int64_t *idx_out, float *val_out, void *workspace, size_t workspace_bytes,
                                void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(emb && queries && n_rows >= 1 && dim >= 1 && dim <= 8192 && n_q >= 0, "b2r_f32_dot_topk: bad arguments");
    B2R_CHECK_ARG(k >= 0 && k <= B2R_TOPK_MAX_FAST && (k > 0 || scores_out), "b2r_f32_dot_topk: k=%d outside [0,%d]", k,
                  B2R_TOPK_MAX_FAST);
    B2R_CHECK_ARG(doc_id_base >= 0 && doc_id_base + n_rows < 0xFFFFFFFFll, "b2r_f32_dot_topk: doc index range");
    if (n_q == 0) return B2R_OK;
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
    float *scores = scores_out;
    int64_t stride = scores_stride;
    if (!scores) {
        stride = (int64_t)align_up((size_t)n_rows, 4);
        scores = static_cast<float *>(carve((size_t)n_q * stride * 4));
    } else {
        B2R_CHECK_ARG(stride >= n_rows, "b2r_f32_dot_topk: scores_stride < n_rows");
    }
    uint64_t *keys = k > 0 ? static_cast<uint64_t *>(carve((size_t)n_q * k * 8)) : nullptr;
    if (!scores || (k > 0 && !keys)) {
        set_error("b2r_f32_dot_topk: workspace too small");
        return B2R_ERR_WORKSPACE;
    }
    const int dim4 = (dim + 3) & ~3;
    int64_t blocks = (n_rows * 32 + DN_THREADS - 1) / DN_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    // queries per pass: 8, fewer when their vectors would not fit 192 KB of shared memory (very wide embeddings)
    int qc = (int)(196608 / ((size_t)dim4 * 4));
    qc = qc < 1 ? 1 : (qc > DN_QC ? DN_QC : qc);
    for (int q0 = 0; q0 < n_q; q0 += qc) {
        const int nq = n_q - q0 < qc ? n_q - q0 : qc;
        const size_t smem = (size_t)nq * dim4 * 4;
        B2R_CUDA(cudaFuncSetAttribute(f32_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        f32_dot_kernel<<<(unsigned)blocks, DN_THREADS, smem, st>>>(emb, n_rows, dim, queries, q0, nq, scores, stride);
        B2R_LAUNCH_CHECK();
    }
    if (k == 0) return B2R_OK;
    int rc = topk_scores_rows(scores, n_q, n_rows, stride, k, doc_id_base, keys, wp, left, st);
    if (rc) return rc;
    return decode_keys(keys, (int64_t)n_q * k, idx_out, val_out, nullptr, 0, k, 0, st);
}
