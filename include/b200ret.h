/*
 * b200ret.h -- C ABI of libb200ret.so: the B200 (sm_100a) retrieval scoring hot path.
 *
 * Every entry point replaces one piece of the reference's Python/Numba hot path (paths relative
 * to the reference root); INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 *   b2r_index_build        <- the CSR half of RetrievalService.build_bm25_index
 *                             (rag_system/core/retrieval.py:176-190): takes the doc-major CSR the
 *                             reference builds and lays it out term-major in HBM.
 *   b2r_search_batch       <- simd_bm25_score + fast_topk_selection as called per query by
 *                             RetrievalService._score_bm25_query (retrieval.py:256-273), batched;
 *                             with an impact-kind index it is simd_tfidf_score + top-k
 *                             (rag_system/pipeline/evaluate_rag_pipeline.py:95-121, :391-399).
 *   b2r_search_batch_host  <- same, host buffers in / host buffers out (the plugin-facing call).
 *   b2r_topk               <- fast_topk_selection (retrieval.py:79-92; retriever_registry.py:75-87;
 *                             evaluate_rag_pipeline.py:124-159).
 *   b2r_merge_candidates   <- no reference counterpart (the reference is single-process); merges the
 *                             per-shard top-k lists of a doc-sharded corpus.
 *   b2r_int8_dot_batch     <- quantized_dot_product_batch (retriever_registry.py:90-117).
 *   b2r_int8_scan_topk     <- QuantizedEmbeddingRetriever.search's scan + argpartition
 *                             (retriever_registry.py:495-515) without materialising [Q, N].
 *
 * Conventions
 *   - Plain C types only; every pointer is a DEVICE pointer unless the name ends in _h.
 *   - The caller owns all memory (PyTorch tensors in the Python host); the library never allocates
 *     device memory.  Workspace sizes come from the *_workspace functions.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*) and, except for the *_host
 *     call, nothing synchronises.
 *   - Return value: 0 on success, negative b2r_status otherwise; b2r_last_error() gives the
 *     thread-local message.  There is no CPU fallback anywhere in this library.
 *   - Ranking rule everywhere: f32 score descending, then document index ascending.  A candidate
 *     travels as a 64-bit key  (ordered_u32(score) << 32) | (0xFFFFFFFF - global_doc_index);
 *     larger key = better rank; key 0 = "no candidate".  Global doc indices must be < 2^32 - 1.
 */
#ifndef B200RET_H
#define B200RET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_VERSION 100
#define B2R_TOPK_MAX_FAST 1024 /* larger k takes the full-sort path of b2r_topk */

typedef enum b2r_status {
    B2R_OK = 0,
    B2R_ERR_ARG = -1,
    B2R_ERR_CUDA = -2,
    B2R_ERR_WORKSPACE = -3,
    B2R_ERR_UNSUPPORTED = -4,
    B2R_ERR_DATA = -5
} b2r_status;

typedef enum b2r_kind {
    B2R_KIND_BM25 = 0,  /* post_val = f64 saturation factor (tf*(k1+1))/(tf+k1*(1-b+b*dl/avgdl)) */
    B2R_KIND_IMPACT = 1 /* post_val = f32 weight; score = sum f64(f32(f32(w*qw)*idf)) */
} b2r_kind;

/* Term-major index over one shard of documents.  A plain descriptor: the buffers belong to the
 * caller.  Postings of term t that fall in document tile T (tile_docs consecutive local docs) are
 * post_doc/post_val[blk_ptr[t*n_tiles+T] .. blk_ptr[t*n_tiles+T+1]), ordered by doc index
 * (non-dense terms) or, for dense terms, by sub-tile and inside a sub-tile segment round-robin over the 16
 * shared-memory accumulator slots (doc mod 16), which keeps the scorer's f64 read-modify-write conflict-free.
 * Terms that average >= B2R_DENSE_MIN_PER_TILE postings per tile ("dense" terms) additionally get
 * a row of sub-tile offsets (B2R_SUBTILES sub-tiles of tile_docs/B2R_SUBTILES docs per tile):
 * postings of dense term t in sub-tile S are [dense_ptr[r*(n_tiles*8+1)+S], dense_ptr[...+S+1])
 * with r = dense_id[t]; the scorer gives every warp one sub-tile and never needs a CTA barrier. */
#define B2R_SUBTILES 8
#define B2R_DENSE_MIN_PER_TILE 64
typedef struct b2r_index {
    int64_t n_docs;      /* documents in this shard */
    int64_t doc_id_base; /* global index of local document 0 */
    int64_t nnz;         /* postings in this shard (< 2^32) */
    int32_t n_vocab;
    int32_t tile_docs;   /* power of two, 256..16384 */
    int32_t n_tiles;     /* ceil(n_docs / tile_docs) */
    int32_t kind;        /* b2r_kind */
    uint32_t *post_doc;  /* [nnz] local doc index */
    void *post_val;      /* [nnz] f64 (BM25) or f32 (IMPACT) */
    uint32_t *blk_ptr;   /* [n_vocab * n_tiles + 1] */
    int32_t *dense_id;   /* [n_vocab] row of dense_ptr, or -1 */
    uint32_t *dense_ptr; /* [n_dense_max * (n_tiles * B2R_SUBTILES + 1)] */
    int32_t n_dense_max; /* rows allocated in dense_ptr (from b2r_index_sizes_for) */
    int32_t reserved0;
    uint32_t *post_pk;   /* optional (BM25, tile_docs == 4096): b2r_index_pack_bytes(nnz) bytes filled by b2r_index_pack;
                            NULL = the search path scores in f64 only */
} b2r_index;

typedef struct b2r_index_sizes {
    size_t post_doc_bytes;
    size_t post_val_bytes;
    size_t blk_ptr_bytes;
    size_t scratch_bytes; /* build-time scratch */
    size_t dense_id_bytes;
    size_t dense_ptr_bytes;
    int64_t n_dense_max;
} b2r_index_sizes;

int b2r_version(void);
const char *b2r_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long b2r_launch_count(void);

/* Byte sizes of the buffers of an index with these dimensions (no device access). */
int b2r_index_sizes_for(int64_t nnz, int64_t n_docs, int32_t n_vocab, int32_t tile_docs, int32_t kind,
                        b2r_index_sizes *out);

/* Fill ix->post_doc / post_val / blk_ptr (caller-allocated, sizes from b2r_index_sizes_for) from a
 * doc-major CSR resident on the device: tf f32[nnz], indices i32[nnz] (any order inside a row),
 * indptr i64[n_docs+1], doc_len f32[n_docs] (ignored for IMPACT).  k1, b, avgdl as the reference
 * holds them (retrieval.py:116-117,190).  Enqueues only; *status_flag (device int32, inside scratch)
 * is checked by b2r_index_build_status after the caller synchronises. */
int b2r_index_build(const b2r_index *ix, const float *tf, const int32_t *indices, const int64_t *indptr,
                    const float *doc_len, double k1, double b, double avgdl, void *scratch,
                    size_t scratch_bytes, void *stream);
/* Synchronises the stream and reports malformed input (term id out of range) found by the build. */
int b2r_index_build_status(const void *scratch, void *stream);

/* Packed copy of a BM25 index for the approximate pre-filter of the search path (csrc/score_approx.cu): 4 bytes per
 * posting -- the f32 bits of post_val rounded to 11 explicit mantissa bits | the 12-bit document offset inside its
 * tile -- plus a 256-byte trailer (largest magnitude, "not finite" flag).  b2r_search_batch then scores every posting
 * in f32 from this copy, keeps the documents that can still be in the top-k under a proven per-query error bound and
 * rescores only those with the f64 chain of post_val: results are bit-identical to the f64-only path, which stays in
 * use for dense score output, impact indexes, other tile sizes and as the exhaustive fallback.  The copy is derived
 * data: it is not part of the index file; call b2r_index_pack after b2r_index_build or after loading a file.
 * b2r_index_pack_status synchronises and returns B2R_ERR_UNSUPPORTED when a value is not finite (the caller then
 * clears ix->post_pk); *u_max_out (optional) receives the largest packed magnitude. */
size_t b2r_index_pack_bytes(int64_t nnz);
int b2r_index_pack(const b2r_index *ix, void *stream);
int b2r_index_pack_status(const b2r_index *ix, void *stream, float *u_max_out);
/* Test / profiling hook: 0 = b2r_search_batch ignores post_pk (f64 scoring of every posting, as before). */
void b2r_set_approx_prefilter(int enabled);

/* ---- On-disk form of a b2r_index (the reference only caches the doc-major CSR as
 * an .npz, evaluate_rag_pipeline.py:280-312).  One file = this header (4096 bytes) followed by the six device
 * buffers verbatim, each starting on a 4096-byte boundary, so a shard can be read (or cuFile-DMAed) straight
 * into HBM with no re-layout.  Little-endian.  The buffers hold shard-local document indices: doc_id_base may
 * be changed at load time. */
#define B2R_FILE_MAGIC "B2RIDX01"
#define B2R_FILE_VERSION 1u
#define B2R_FILE_ALIGN 4096u
enum { B2R_SEC_POST_DOC = 0, B2R_SEC_POST_VAL, B2R_SEC_BLK_PTR, B2R_SEC_DENSE_ID, B2R_SEC_DENSE_PTR, B2R_SEC_IDF,
       B2R_SEC_COUNT };
typedef struct b2r_file_section {
    uint64_t offset;   /* from the start of the file, multiple of B2R_FILE_ALIGN */
    uint64_t bytes;    /* the device buffer size (b2r_index_sizes_for); idf: 4 * n_vocab */
    uint64_t checksum; /* b2r_checksum64 of the section */
} b2r_file_section;
typedef struct b2r_index_file_header {
    char magic[8];
    uint32_t version;
    uint32_t header_bytes; /* B2R_FILE_ALIGN */
    int64_t n_docs, doc_id_base, nnz;
    int32_t n_vocab, tile_docs, n_tiles, kind, n_dense_max, subtiles /* B2R_SUBTILES */;
    double k1, b, avgdl;   /* baked into post_val of a BM25 index */
    b2r_file_section sections[B2R_SEC_COUNT];
} b2r_index_file_header;
/* Host-only helpers (no device access).  b2r_checksum64: order-dependent 64-bit checksum of a byte range
 * (any length).  b2r_index_file_layout: fills magic/version/sizes/offsets of *hdr from its dimension fields
 * (n_docs, nnz, n_vocab, tile_docs, kind, n_dense_max; checksums are left 0) and returns the total file size in
 * *file_bytes.  b2r_index_file_check: validates a header read from a file of file_bytes bytes (magic, version,
 * dimensions, section sizes and bounds) -- B2R_ERR_DATA with a message when it is not a loadable index. */
uint64_t b2r_checksum64(const void *data, size_t bytes);
int b2r_index_file_layout(b2r_index_file_header *hdr, uint64_t *file_bytes);
int b2r_index_file_check(const b2r_index_file_header *hdr, uint64_t file_bytes);

/* Workspace for b2r_search_batch: *min_bytes lets it run one query at a time, *full_bytes lets it
 * score all n_queries in one pass; anything in between is used as given. */
int b2r_search_workspace(const b2r_index *ix, int32_t n_queries, int32_t k, size_t *min_bytes,
                         size_t *full_bytes);

/* Batched scoring + selection.  Queries are CSR-style: terms of query q are
 * q_terms[q_ptr[q]..q_ptr[q+1]) (ascending, unique, weights > 0 -- the positive entries of the
 * reference's dense query_tf vector), idf is f32[n_vocab].
 *   scores_out : optional f32[n_queries, scores_stride] dense scores (stride >= n_tiles*tile_docs,
 *                multiple of 4); entries [n_docs, stride) of a row are padding.
 *   keys_out   : optional u64[n_queries, k] ranked candidate keys (needs k >= 1).
 *   idx_out/val_out : optional i64 / f32 [n_queries, k]; -1 / -inf where fewer than k docs exist.
 */
int b2r_search_batch(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                     const float *idf, int32_t n_queries, int32_t k, float *scores_out, int64_t scores_stride,
                     uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace,
                     size_t workspace_bytes, void *stream);

/* Test / profiling hook: 0 makes b2r_search_batch use the plain "score everything, then select" path
 * instead of the fused-selection path (both are exact). */
void b2r_set_fused_selection(int enabled);
/* Test hook: candidate-list capacity of the fused path (0 = the plan's own choice).  A tiny capacity makes every
 * list overflow, which routes every query through the exhaustive fallback (results are identical). */
void b2r_set_fused_cap(int cap);
/* Test / profiling hook: 0 = later index builds keep dense segments doc-ascending (no bank schedule). */
void b2r_set_bank_schedule(int enabled);

/* Profiling hooks used by bench.py: bracket the fused scoring launch of b2r_search_batch with CUDA
 * events on its stream; b2r_profile_fused_ms returns the duration of the most recent one.
 * b2r_fused_plan reports whether (and how) the fused path applies: *n_sample_tiles == 0 means the
 * plain path; otherwise every tile_step-th tile forms the threshold sample. */
int b2r_set_profiling(int enabled);
int b2r_profile_fused_ms(float *ms, int32_t *reserved);
int b2r_fused_plan(const b2r_index *ix, int32_t k, int32_t *n_sample_tiles, int32_t *tile_step, int32_t *cap);

/* Same with HOST query buffers and HOST outputs (pinned memory recommended): copies the queries in,
 * runs b2r_search_batch, copies idx/val/keys out and synchronises the stream.  `workspace` must hold
 * b2r_search_host_extra_bytes(n_queries, n_query_terms, k) more bytes than the device call needs. */
size_t b2r_search_host_extra_bytes(int32_t n_queries, int64_t n_query_terms, int32_t k);
int b2r_search_batch_host(const b2r_index *ix, const int32_t *q_ptr_h, const int32_t *q_terms_h,
                          const float *q_weights_h, const float *idf, int32_t n_queries, int32_t k,
                          uint64_t *keys_out_h, int64_t *idx_out_h, float *val_out_h, void *workspace,
                          size_t workspace_bytes, void *stream);

/* Row-wise top-k of f32 scores [n_rows, n] (row stride in elements).  Global index of element j of a
 * row is doc_id_base + j.  val_out holds scores[idx] bit-for-bit.  k <= n required (the caller
 * clamps, as the reference does with `k >= n`). */
int b2r_topk_workspace(int64_t n_rows, int64_t n, int32_t k, size_t *bytes);
int b2r_topk(const float *scores, int64_t n_rows, int64_t n, int64_t row_stride, int32_t k, int64_t doc_id_base,
             uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace, size_t workspace_bytes,
             void *stream);

/* Merge the ranked candidate keys of n_parts shards, gathered as u64[n_parts, n_queries, k], into
 * the global top-k per query.  workspace >= n_queries*k*8 bytes when keys_out is NULL. */
int b2r_merge_candidates(const uint64_t *gathered, int32_t n_parts, int32_t n_queries, int32_t k,
                         uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace,
                         size_t workspace_bytes, void *stream);

/* The same merge with the exchange built in, over peer memory (NVLink): every rank passes its ranked keys
 * local_keys u64[n_queries, k] and the DEVICE array peer_bufs_dev[n_parts] of all ranks' receive buffers (memory every
 * rank of the box can write, e.g. PyTorch symmetric memory; b2r_exchange_bytes(...) bytes each, zero-filled once and
 * then owned by this call sequence).  One launch per rank: push my keys into every rank's buffer, wait chunk by chunk
 * for the other ranks' keys, merge.  All ranks must make the same sequence of calls (same n_queries, k) on the same
 * buffers; no host synchronisation and no NCCL call is involved, so the step can be captured in a CUDA graph.
 * b2r_exchange_status synchronises and reports a timed-out wait (a rank that never called). */
size_t b2r_exchange_bytes(int32_t n_parts, int32_t n_queries, int32_t k);
int b2r_exchange_merge(const uint64_t *local_keys, void *const *peer_bufs_dev, int32_t rank, int32_t n_parts,
                       int32_t n_queries, int32_t k, uint64_t *keys_out, int64_t *idx_out, float *val_out, void *stream);
int b2r_exchange_status(const void *recv_buf, void *stream);

/* Decode ranked keys into (global doc index, f32 score). */
int b2r_decode_keys(const uint64_t *keys, int64_t n, int64_t *idx_out, float *val_out, void *stream);

/* INT8 dense similarity, out f32[n_q, n_docs] = f32((f64(dot) * f64(qscale[q])) * f64(dscale[d])). */
int b2r_int8_dot_batch(const int8_t *q8, int32_t n_q, const int8_t *d8, int64_t n_docs, int32_t dim,
                       const float *q_scale, const float *d_scale, float *out, void *stream);
/* Test / profiling hook: 0 forces the dp4a kernel even for shapes the tcgen05 kernel supports. */
void b2r_set_int8_mma(int enabled);
/* Test / profiling hook: largest thread-block cluster (1, 2, 4 or 8 query-tile CTAs sharing every document
 * chunk by TMA multicast) the tcgen05 kernel may use; 1 = no clusters. */
void b2r_set_int8_cluster(int max_cluster);
/* Test / profiling hook: 1 = the fused scan of batches of more than 128 queries runs on CTA pairs
 * (tcgen05.mma cta_group::2, 256 documents x 256 queries per pair tile); 0 = single-CTA 128 x 128 tiles. */
void b2r_set_int8_pair(int enabled);
/* Test / profiling hook: 0 = b2r_int8_scan_topk uses the plain chunked "dense tile + select" path;
 * otherwise (default) the fused-selection path: sampled threshold, f32 pre-filter + exact f64 check in the
 * MMA epilogue, candidate lists, device-gated exact fallback. */
void b2r_set_int8_fused(int enabled);
/* Exhaustive INT8 scan with fused per-query top-k (never materialises [n_q, n_docs]). */
int b2r_int8_scan_workspace(int32_t n_q, int64_t n_docs, int32_t dim, int32_t k, size_t *bytes);
int b2r_int8_scan_topk(const int8_t *q8, int32_t n_q, const int8_t *d8, int64_t n_docs, int32_t dim,
                       const float *q_scale, const float *d_scale, int32_t k, int64_t doc_id_base,
                       uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace,
                       size_t workspace_bytes, void *stream);

/* Hybrid sparse -> dense rerank on candidates only (the reference names a "hybrid"
 * retriever in configs/ms_marco_paper_results.yaml:108-124 but has no implementation).  cand_idx i64[n_q, k_in]
 * holds global document indices (e.g. idx_out of b2r_search_batch; < 0 or outside this shard = no candidate),
 * cand_sparse f32[n_q, k_in] their sparse scores or NULL.  For every pair the INT8 similarity
 * f32((f64(dot) * f64(q_scale[q])) * f64(d_scale[doc])) is evaluated on the candidate's vector
 * (retriever_registry.py:90-117 restricted to the candidates), then
 *   score = cand_sparse ? f32(f64(sparse_weight) * f64(sparse) + f64(dense_weight) * f64(dense)) : dense
 * and the k_out best per query are returned (score descending, document index ascending; -1 / -inf padding).
 * dense_out: optional f32[n_q, k_in] dense similarities (-inf for "no candidate"). */
int b2r_int8_rerank_workspace(int32_t n_q, int32_t k_in, int32_t k_out, size_t *bytes);
int b2r_int8_rerank(const int64_t *cand_idx, const float *cand_sparse, int32_t n_q, int32_t k_in, const int8_t *q8,
                    const float *q_scale, const int8_t *d8, const float *d_scale, int64_t n_docs, int32_t dim,
                    int64_t doc_id_base, double sparse_weight, double dense_weight, int32_t k_out, float *dense_out,
                    int64_t *idx_out, float *val_out, void *workspace, size_t workspace_bytes, void *stream);

/* fp32 dense similarity + top-k: scores[q, r] = <emb[r, :], queries[q, :]> over a row-major f32[n_rows, dim]
 * matrix resident on the device, then the k best rows per query (reference: RetrievalService.search_by_vector,
 * rag_system/core/retrieval.py:402-436 -- np.dot through host BLAS + fast_topk_selection).  f32 products and
 * f32 pairwise sums: parity with BLAS is by tolerance (1e-5 * sum |a_i b_i|), the selection is exact on the
 * computed scores.  scores_out: optional f32[n_q, scores_stride] (then k may be 0); idx/val as everywhere. */
int b2r_f32_dot_topk_workspace(int32_t n_q, int64_t n_rows, int32_t k, size_t *bytes);
int b2r_f32_dot_topk(const float *emb, int64_t n_rows, int32_t dim, const float *queries, int32_t n_q, int32_t k,
                     int64_t doc_id_base, float *scores_out, int64_t scores_stride, int64_t *idx_out, float *val_out,
                     void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RET_H */
