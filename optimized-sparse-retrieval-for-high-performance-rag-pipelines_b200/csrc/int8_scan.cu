// INT8 dense similarity scan (reference: quantized_dot_product_batch,
// rag_system/core/retriever_registry.py:90-117, and the scan + argpartition of
// QuantizedEmbeddingRetriever.search, :495-515).
//
// Round-1 kernel: shared-memory tiled dp4a (IDP.4A) contraction -- exact int32 dot products --
// followed by the reference's f64 scale chain  f32((f64(dot) * f64(qs)) * f64(ds)).
// b2r_int8_scan_topk walks the corpus in document chunks: dots of one chunk go to a workspace
// tile [n_q, chunk], the streaming top-k (topk.cu) reduces it to k keys per query, and the
// per-chunk winners are merged at the end, so [n_q, n_docs] is never materialised.
#include "common.cuh"

namespace b2r {

constexpr int I8_THREADS = 256;
constexpr int I8_DT = 128;          // docs per CTA tile   (lane l owns docs l, l+32, l+64, l+96)
constexpr int I8_QT = 32;           // queries per CTA tile (warp w owns queries 4w..4w+3)
constexpr int I8_KC = 128;          // bytes of the embedding dimension staged per step
constexpr int I8_ROW = I8_KC + 16;  // padded smem row (144 B = 36 words: conflict-free LDS.128 across rows)

__device__ __forceinline__ void i8_stage_rows(int8_t *smem, const int8_t *__restrict__ g, int64_t row0, int64_t n_rows,
                                              int rows_in_tile, int dim, int k0, bool vec) {
    // copies rows [row0, row0+rows_in_tile) x bytes [k0, k0+KC) into smem (zero filled outside)
    const int tid = threadIdx.x;
    if (vec) {
        for (int v = tid; v < rows_in_tile * (I8_KC / 16); v += I8_THREADS) {
            int r = v / (I8_KC / 16), c = (v % (I8_KC / 16)) * 16;
            int4 val = make_int4(0, 0, 0, 0);
            if (row0 + r < n_rows && k0 + c < dim)
                val = __ldg(reinterpret_cast<const int4 *>(g + (row0 + r) * (int64_t)dim + k0 + c));
            *reinterpret_cast<int4 *>(smem + r * I8_ROW + c) = val;
        }
    } else {
        for (int v = tid; v < rows_in_tile * I8_KC; v += I8_THREADS) {
            int r = v / I8_KC, c = v % I8_KC;
            int8_t val = 0;
            if (row0 + r < n_rows && k0 + c < dim) val = g[(row0 + r) * (int64_t)dim + k0 + c];
            smem[r * I8_ROW + c] = val;
        }
    }
}

__global__ void __launch_bounds__(I8_THREADS)
int8_dot_kernel(const int8_t *__restrict__ q8, int n_q, const int8_t *__restrict__ d8, int64_t n_docs, int dim,
                const float *__restrict__ q_scale, const float *__restrict__ d_scale, float *__restrict__ out,
                int64_t out_stride, bool vec) {
    __shared__ __align__(16) int8_t s_d[I8_DT * I8_ROW];
    __shared__ __align__(16) int8_t s_q[I8_QT * I8_ROW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t doc0 = (int64_t)blockIdx.x * I8_DT;
    const int q0 = blockIdx.y * I8_QT;
    int acc[4][4];  // [doc j][query i]
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0;

    for (int k0 = 0; k0 < dim; k0 += I8_KC) {
        __syncthreads();
        i8_stage_rows(s_d, d8, doc0, n_docs, I8_DT, dim, k0, vec);
        i8_stage_rows(s_q, q8, q0, n_q, I8_QT, dim, k0, vec);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < I8_KC; c += 16) {
            int4 dv[4], qv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) dv[j] = *reinterpret_cast<const int4 *>(s_d + (lane + 32 * j) * I8_ROW + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const int4 *>(s_q + (warp * 4 + i) * I8_ROW + c);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[j][i] = __dp4a(dv[j].x, qv[i].x, acc[j][i]);
                    acc[j][i] = __dp4a(dv[j].y, qv[i].y, acc[j][i]);
                    acc[j][i] = __dp4a(dv[j].z, qv[i].z, acc[j][i]);
                    acc[j][i] = __dp4a(dv[j].w, qv[i].w, acc[j][i]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + warp * 4 + i;
        if (q >= n_q) continue;
        const double qs = (double)q_scale[q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t d = doc0 + lane + 32 * j;
            if (d < n_docs) {
                double v = __dmul_rn(__dmul_rn((double)acc[j][i], qs), (double)d_scale[d]);
                out[(int64_t)q * out_stride + d] = __double2float_rn(v);
            }
        }
    }
}

static int launch_int8_dot(const int8_t *q8, int n_q, const int8_t *d8, int64_t n_docs, int dim, const float *qs,
                           const float *ds, float *out, int64_t stride, cudaStream_t st) {
    if (n_q == 0 || n_docs == 0) return B2R_OK;
    bool vec = (dim % 16 == 0) && ((reinterpret_cast<uintptr_t>(q8) & 15) == 0) &&
               ((reinterpret_cast<uintptr_t>(d8) & 15) == 0);
    int64_t gx = (n_docs + I8_DT - 1) / I8_DT;
    int gy = (n_q + I8_QT - 1) / I8_QT;
    B2R_CHECK_ARG(gx < 0x7FFFFFFFll && gy <= 65535, "int8 scan: grid too large");
    dim3 grid((unsigned)gx, (unsigned)gy);
    int8_dot_kernel<<<grid, I8_THREADS, 0, st>>>(q8, n_q, d8, n_docs, dim, qs, ds, out, stride, vec);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

static int64_t i8_chunk_docs(int32_t n_q, int64_t n_docs) {
    // keep the per-chunk score tile around 1 GiB
    int64_t c = ((int64_t)1 << 28) / (n_q > 0 ? n_q : 1);
    c = (c / 4096) * 4096;
    if (c < 4096) c = 4096;
    if (c > n_docs) c = (n_docs + 3) / 4 * 4;
    return c;
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_int8_dot_batch(const int8_t *q8, int32_t n_q, const int8_t *d8, int64_t n_docs, int32_t dim,
                                  const float *q_scale, const float *d_scale, float *out, void *stream) {
    B2R_CHECK_ARG(q8 && d8 && q_scale && d_scale && out && dim >= 1 && n_q >= 0 && n_docs >= 0,
                  "b2r_int8_dot_batch: bad arguments");
    B2R_CHECK_ARG((int64_t)dim * 127 * 127 < 0x7FFFFFFFll, "b2r_int8_dot_batch: dim too large for int32 dots");
    return launch_int8_dot(q8, n_q, d8, n_docs, dim, q_scale, d_scale, out, n_docs, static_cast<cudaStream_t>(stream));
}

extern "C" int b2r_int8_scan_workspace(int32_t n_q, int64_t n_docs, int32_t dim, int32_t k, size_t *bytes) {
    (void)dim;
    B2R_CHECK_ARG(bytes && n_q >= 0 && n_docs >= 1 && k >= 1 && k <= B2R_TOPK_MAX_FAST,
                  "b2r_int8_scan_workspace: bad arguments");
    int64_t nq = n_q > 0 ? n_q : 1;
    int64_t chunk = i8_chunk_docs(n_q, n_docs);
    int64_t n_chunks = (n_docs + chunk - 1) / chunk;
    *bytes = align_up((size_t)nq * chunk * 4, 256) + topk_ws_bytes(nq, chunk, k) +
             align_up((size_t)n_chunks * nq * k * 8, 256) + topk_keys_ws_bytes(nq, n_chunks * k, k) +
             align_up((size_t)nq * k * 8, 256) + 1024;
    return B2R_OK;
}

extern "C" int b2r_int8_scan_topk(const int8_t *q8, int32_t n_q, const int8_t *d8, int64_t n_docs, int32_t dim,
                                  const float *q_scale, const float *d_scale, int32_t k, int64_t doc_id_base,
                                  uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(q8 && d8 && q_scale && d_scale && dim >= 1 && n_q >= 0 && n_docs >= 1,
                  "b2r_int8_scan_topk: bad arguments");
    B2R_CHECK_ARG(k >= 1 && k <= B2R_TOPK_MAX_FAST, "b2r_int8_scan_topk: k=%d outside [1,%d]", k, B2R_TOPK_MAX_FAST);
    B2R_CHECK_ARG((int64_t)dim * 127 * 127 < 0x7FFFFFFFll, "b2r_int8_scan_topk: dim too large for int32 dots");
    B2R_CHECK_ARG(doc_id_base >= 0 && doc_id_base + n_docs < 0xFFFFFFFFll, "b2r_int8_scan_topk: doc index range");
    if (n_q == 0) return B2R_OK;
    size_t need = 0;
    b2r_int8_scan_workspace(n_q, n_docs, dim, k, &need);
    if (workspace_bytes < need) {
        set_error("b2r_int8_scan_topk: workspace too small (%zu < %zu)", workspace_bytes, need);
        return B2R_ERR_WORKSPACE;
    }
    char *wp = static_cast<char *>(workspace);
    auto carve = [&](size_t bytes) -> void * {
        void *p = wp;
        wp += align_up(bytes, 256);
        return p;
    };
    const int64_t chunk = i8_chunk_docs(n_q, n_docs);
    const int64_t n_chunks = (n_docs + chunk - 1) / chunk;
    float *tile = static_cast<float *>(carve((size_t)n_q * chunk * 4));
    size_t tk_ws = topk_ws_bytes(n_q, chunk, k);
    void *tk = carve(tk_ws);
    uint64_t *part = static_cast<uint64_t *>(carve((size_t)n_chunks * n_q * k * 8));
    size_t mg_ws = topk_keys_ws_bytes(n_q, n_chunks * k, k);
    void *mg = carve(mg_ws);
    uint64_t *keys = keys_out ? keys_out : static_cast<uint64_t *>(carve((size_t)n_q * k * 8));
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t d0 = c * chunk;
        const int64_t nd = (n_docs - d0) < chunk ? (n_docs - d0) : chunk;
        int rc = launch_int8_dot(q8, n_q, d8 + d0 * dim, nd, dim, q_scale, d_scale + d0, tile, chunk, st);
        if (rc) return rc;
        rc = topk_scores_rows(tile, n_q, nd, chunk, k, doc_id_base + d0,
                              n_chunks == 1 ? keys : part + c * (int64_t)n_q * k, tk, tk_ws, st);
        if (rc) return rc;
    }
    if (n_chunks > 1) {
        int rc = topk_keys_rows(part, n_q, n_chunks * k, k, k, (int64_t)n_q * k, k, keys, mg, mg_ws, st);
        if (rc) return rc;
    }
    return decode_keys(keys, (int64_t)n_q * k, idx_out, val_out, nullptr, 0, k, 0, st);
}
