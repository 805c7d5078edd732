// Approximate pre-filter of the BM25 search path: 4-byte postings, f32 accumulators, exact f64 rescoring of the
// handful of documents that can still be in the top-k.
// (reference: simd_bm25_score + fast_topk_selection as called by RetrievalService._score_bm25_query,
//  rag_system/core/retrieval.py:41-92, :256-273 -- the RESULT is bit-identical to the f64 chain of the reference; only
//  the order in which documents are ruled out changes.)
//
// The exact scorer (score.cu) spends its time on the shared-memory pipe: an f64 read-modify-write per posting, fed by
// 12 B of posting data.  Selection, however, needs the exact score of very few documents.  So the search path scores
// every posting ONCE in f32 from a packed copy of the index
//     post_pk[p] = (f32 bits of u, rounded to 11 explicit mantissa bits: the upper 20 bits) |
//                  (byte offset of the document's accumulator inside its warp's 1024-document region: bits 2..11) |
//                  (the owning warp: bits 0..1)
// (4 B per posting, one LDG.32 and two LOPs to decode; u = the BM25 saturation factor of post_val), with a proven bound
// on |approx - exact| per document (approx_bound_warp below): relative to the score, delta = 2^-12 + (n_q + 16) 2^-23,
// plus 2 delta N for the negative contributions a document may have received (N = u_max * sum of |negative weights|)
// (value rounding 2^-12 relative, f32 weight rounding and n_q fused multiply-adds 2^-24 each, f32 rounding of the exact
// score 2^-24; any summation order).  Pipeline per batch, same launches as the exact fused path:
//   1. MAXIMA epilogue on the sample tiles -> T = k-th largest group maximum of the APPROXIMATE scores (kth_of_maxima);
//      k documents have approx >= T, hence exact >= g(T) (the lower end of the error interval), hence the exact k-th
//      best E_k >= g(T).
//   2. FUSED epilogue on all tiles: candidates = touched documents whose interval reaches g(T), i.e. approx >=
//      h^-1(g(T)) (every member of the exact top-k has exact >= E_k).  If that threshold is not positive (or there is
//      no threshold: few matches) the query runs in "positive mode": candidates = touched documents whose interval
//      reaches above 0, a superset of the documents with a positive exact score; the result is accepted only if the
//      exact k-th best is > 0, otherwise the query takes the exhaustive exact fallback (score.cu) like an overflowed
//      list.
//   3. select + rescore (one CTA per query): A_k = k-th largest approximate score among the candidates; the survivors
//      approx >= h^-1(g(A_k)) (k + a few documents) are rescored EXACTLY by one warp each -- the reference's f64 chain
//      over the query's terms in ascending term id, postings looked up in the f64 index -- keyed, sorted, and the k best
//      are written out.
// "Touched": accumulators are cleared to -0.0f, and a sum that starts at -0.0 stays -0.0 only if every contribution was
// <= 0 and below 2^-150 in magnitude, i.e. exact <= 0: such a document (like an untouched one, exact == 0) is never in
// the top-k when E_k > 0, which both modes guarantee.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace b2r {

constexpr int AP_TILE = 4096;                // the packed format holds 12 doc bits: tile_docs must be 4096
#ifndef AP_WARPS_N
#define AP_WARPS_N 4
#endif
#ifndef AP_MIN_CTAS
#define AP_MIN_CTAS 12
#endif
#ifndef AP_UNROLL
#define AP_UNROLL 4
#endif
constexpr int AP_WARPS = AP_WARPS_N;         // one warp = 1024 documents = two sub-tiles of the index
constexpr int AP_THREADS = 32 * AP_WARPS;
constexpr int AP_SUB = AP_TILE / AP_WARPS;
constexpr int AP_GROUPS_PER_TILE = 256;      // MAXIMA: group maxima per tile (must equal score.cu's SC_GROUPS_PER_TILE)
#ifndef AP_FORMAT
#define AP_FORMAT 2
#endif
// packed posting formats (derived data, rebuilt by b2r_index_pack):
//   2: value in the upper 20 bits | byte offset of the accumulator inside its warp's 1024-document region (bits 2..11)
//      | owning warp (bits 0..1): one LOP gives the address, the value keeps 11 explicit mantissa bits
//   0: value in the upper 20 bits | document offset inside the tile (12 bits): LOP + shift for the address
//   1: value in the upper 18 bits | byte offset inside the tile (bits 2..13): one LOP, but 9 mantissa bits -- the 4x
//      wider error bound overflows the candidate lists of low-scoring queries (3 of config 2's 1024), measured slower
#if AP_FORMAT == 1
constexpr uint32_t AP_VAL_MASK = 0xFFFFC000u, AP_VAL_HALF = 0x2000u;
constexpr double AP_VAL_EPS = 0x1p-10, AP_UMAX_SLACK = 1.001953125;   // value rounding; |u| <= |packed| / (1 - 2^-10)
#else
constexpr uint32_t AP_VAL_MASK = 0xFFFFF000u, AP_VAL_HALF = 0x800u;
constexpr double AP_VAL_EPS = 0x1p-12, AP_UMAX_SLACK = 1.0009765625;
#endif
#ifndef AP_BALLOT
#define AP_BALLOT 1     // 1: the term loop visits only the terms with postings in the tile (one vote, no shuffles for the rest)
#endif
constexpr uint32_t AP_NEGZERO = 0x80000000u;
constexpr uint64_t AP_FLOOR_KEY = (0x80000000ull << 32) | 0xFFFFFFFFull;  // kth_of_maxima's "strictly positive only"
constexpr int AP_SURV_MAX = 1024;            // survivors rescored per query; more (mass ties) -> exhaustive fallback
constexpr int AP_ST_TERMS = 64;              // rescoring: queries of up to this many terms go one warp per (survivor, term)

static bool g_approx_enabled = true;

// page of -0.0f the accumulators are cleared from (bulk copy, like score.cu's zero page); filled by b2r_index_pack
__device__ __align__(128) uint32_t g_negzero_page[AP_SUB];

struct PackMeta {          // trailer of the post_pk buffer (device)
    uint32_t max_bits;     // f32 bits of max |packed value|
    uint32_t bad;          // a value was not finite: the approximate path must not be used
    uint32_t neg;          // a value is negative: the bound cannot separate positive from negative contributions
};

__host__ __device__ static inline size_t pack_meta_offset(int64_t nnz) { return ((size_t)nnz * 4 + 255) / 256 * 256; }

__global__ void __launch_bounds__(256)
pack_postings_kernel(const uint32_t *__restrict__ post_doc, const double *__restrict__ post_val, int64_t nnz,
                     uint32_t *__restrict__ pk, PackMeta *__restrict__ meta) {
    uint32_t mx = 0, bad = 0, neg = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const float v = __double2float_rn(post_val[p]);
        uint32_t r = 0;
        if (!(fabsf(v) <= 3.0e38f)) {   // inf, NaN, or so large that rounding could reach inf
            bad = 1;
        } else {
            r = (__float_as_uint(v) + AP_VAL_HALF) & AP_VAL_MASK;   // round to nearest (magnitude)
        }
        const uint32_t d = post_doc[p] & (uint32_t)(AP_TILE - 1);
#if AP_FORMAT == 2
        pk[p] = r | ((d & (uint32_t)(AP_SUB - 1)) << 2) | (d / (uint32_t)AP_SUB);
#elif AP_FORMAT == 1
        pk[p] = r | (d << 2);
#else
        pk[p] = r | d;
#endif
        mx = max(mx, r & 0x7fffffffu);
        neg |= (r & 0x7fffffffu) != 0 && (r >> 31);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        neg |= __shfl_xor_sync(0xffffffffu, neg, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mx) atomicMax(&meta->max_bits, mx);
        if (bad) atomicOr(&meta->bad, 1u);
        if (neg) atomicOr(&meta->neg, 1u);
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < AP_SUB; i += blockDim.x) g_negzero_page[i] = AP_NEGZERO;
}

// ---- per-query error bound and filter threshold (identical in the scoring and the selection kernel) --------------
// For a document with exact contributions c_t (sum S, positive part P, negative part N_d) the approximate sum s obeys
//     |s - S| <= delta * sum_t |c_t| = delta * (P + N_d) <= delta * (|S| + 2 N),   N = u_max * sum_{w_t < 0} |w_t|
// (delta: value rounding + f32 weight rounding + n fused multiply-adds; N bounds the negative part of ANY document: the
// packed values are >= 0 -- otherwise N is taken over all terms).  With dp = delta / (1 - delta) and c2 = 2 dp N:
//     g(s) = s - dp |s| - c2  <=  S  <=  h(s) = s + dp |s| + c2,       g and h increasing.
// The bound is RELATIVE to the score: a query of eight rare terms (sum |w| u_max ~ 200) with its k-th best score at 9
// loses 0.02 of threshold, not 0.4, so the candidate lists stay as short as those of the f64 path.
struct ApproxBound {
    double dp, c2;
    float thr_lo;   // candidates: touched && approx >= thr_lo
    int pos_mode;   // 1: accept only if the exact k-th best is > 0
    int ok;         // 0: the bound is not finite (absurd weights): the query takes the exact fallback
};

// all 32 lanes of a warp; every lane returns the same values
__device__ __forceinline__ ApproxBound approx_bound_warp(int qs, int qe, const int32_t *__restrict__ q_terms,
                                                         const float *__restrict__ q_weights,
                                                         const float *__restrict__ idf, const PackMeta *meta,
                                                         uint64_t thr_key, int lane) {
    double s = 0.0, sn = 0.0;
    for (int j = qs + lane; j < qe; j += 32) {
        const double wq = __dmul_rn((double)__ldg(idf + __ldg(q_terms + j)), (double)__ldg(q_weights + j));
        s = __dadd_rn(s, fabs(wq));
        if (!(wq >= 0.0)) sn = __dadd_rn(sn, fabs(wq));   // (a NaN weight lands here too and makes the bound NaN)
    }
    for (int o = 16; o > 0; o >>= 1) {
        s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        sn = __dadd_rn(sn, __shfl_xor_sync(0xffffffffu, sn, o));
    }
    if (meta->neg) sn = s;
    const double u_max = (double)__uint_as_float(meta->max_bits) * AP_UMAX_SLACK;
    const double B = __dmul_rn(s, u_max), N = __dmul_rn(sn, u_max) * 1.0000001;
    const double n = (double)(qe - qs);
    const double delta = AP_VAL_EPS + (n + 16.0) * 0x1p-23;
    ApproxBound r;
    r.dp = delta / (1.0 - delta) * 1.000001;
    // absolute slack: packed values below the normal f32 range (u < 2^-126) carry an absolute, not a relative, rounding
    // error of at most 2^-137; results of the fused multiply-adds in the denormal range one of 2^-150 each
    r.c2 = 2.0 * r.dp * N + s * 0x1p-135 + (n + 1.0) * 0x1p-140;
    r.ok = (r.c2 == r.c2) && B < 1.0e30 && delta < 0.25 && meta->bad == 0;
    const uint32_t hi = (uint32_t)(thr_key >> 32);
    const bool floor_mode = thr_key == AP_FLOOR_KEY || hi <= 0x80000000u;   // no threshold, or not a positive one
    const double ta = floor_mode ? 0.0 : (double)unord_f32(hi);
    // k sample documents have approx >= ta, hence exact >= g(ta): every member of the top-k has exact >= L (the
    // factor covers the f32 rounding of the ranked score) and therefore h(approx) >= L
    const double L = (ta - r.dp * ta - r.c2) * (1.0 - 0x1p-22);
    r.pos_mode = floor_mode || !(L > r.c2);
    const double lo = r.pos_mode ? -(r.c2 / (1.0 - r.dp)) * 1.000001 - 0x1p-130
                                 : (L - r.c2) / (1.0 + r.dp) * (1.0 - 0x1p-20);
    r.thr_lo = __double2float_rd(lo);
    if (!r.ok) r.thr_lo = __int_as_float(0x7f800000);   // +inf: nothing qualifies; the selection kernel marks the query
    return r;
}

__device__ __forceinline__ uint32_t ap_ld_stream(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ap_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ap_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tAP_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra AP_DONE;\n\tbra AP_WAIT;\n\tAP_DONE:\n\t}\n"
        :
        : "r"(bar), "r"(parity)
        : "memory");
}

// acc[doc in tile] = fma(w, value, acc) -- one LDS.32 / FFMA / STS.32 per posting.  FIRST: the warp's accumulators
// were just cleared to -0.0, so the read is skipped (fma(w, v, -0.0) is evaluated all the same).
// byte offset of a posting's accumulator from `base`, and the warp that owns it
#if AP_FORMAT == 2     // base = the owning warp's region
__device__ __forceinline__ uint32_t ap_offset(uint32_t pk) { return pk & 0xFFCu; }
__device__ __forceinline__ uint32_t ap_owner(uint32_t pk) { return pk & 3u; }
#elif AP_FORMAT == 1   // base = the tile
__device__ __forceinline__ uint32_t ap_offset(uint32_t pk) { return pk & 0x3FFCu; }
__device__ __forceinline__ uint32_t ap_owner(uint32_t pk) { return (pk & 0x3FFCu) / (4u * AP_SUB); }
#else
__device__ __forceinline__ uint32_t ap_offset(uint32_t pk) { return (pk & 0xFFFu) << 2; }
__device__ __forceinline__ uint32_t ap_owner(uint32_t pk) { return (pk & 0xFFFu) / (uint32_t)AP_SUB; }
#endif

// acc[doc] = fma(w, value, acc[doc]) -- one LDS.32 / FFMA / STS.32 per posting.  FIRST: the warp's accumulators
// were just cleared to -0.0, so the read is skipped (fma(w, v, -0.0) is evaluated all the same).
template <bool FIRST>
__device__ __forceinline__ void ap_apply(uint32_t base_s, uint32_t pk, float w) {   // base_s: shared-window address
    const uint32_t slot = base_s + ap_offset(pk);
    const float v = __uint_as_float(pk & AP_VAL_MASK);
    float a = __uint_as_float(AP_NEGZERO);
    if (!FIRST) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(slot) : "memory");
    const float r = __fmaf_rn(w, v, a);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot), "f"(r) : "memory");
}

template <bool FIRST>
__device__ __forceinline__ void ap_apply_term(int dense, uint32_t beg, uint32_t end, int lane, int w,
                                              const uint32_t *__restrict__ post_pk, uint32_t acc, float wt) {
    if (dense) {   // [beg, end) are exactly this warp's postings (its two sub-tiles are adjacent in the index)
        uint32_t p = beg + lane;
#if AP_UNROLL == 8
        for (; p + 224 < end; p += 256) {
            uint32_t a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ap_ld_stream(post_pk + p + 32 * i);
#pragma unroll
            for (int i = 0; i < 8; ++i) ap_apply<FIRST>(acc, a[i], wt);
        }
#endif
        for (; p + 96 < end; p += 128) {
            const uint32_t a0 = ap_ld_stream(post_pk + p), a1 = ap_ld_stream(post_pk + p + 32);
            const uint32_t a2 = ap_ld_stream(post_pk + p + 64), a3 = ap_ld_stream(post_pk + p + 96);
            ap_apply<FIRST>(acc, a0, wt);
            ap_apply<FIRST>(acc, a1, wt);
            ap_apply<FIRST>(acc, a2, wt);
            ap_apply<FIRST>(acc, a3, wt);
        }
        for (; p < end; p += 32) ap_apply<FIRST>(acc, ap_ld_stream(post_pk + p), wt);
    } else {       // the tile's small block: every warp scans it and keeps the postings of its own 1024 documents
        for (uint32_t p = beg + lane; p < end; p += 32) {
            const uint32_t a = __ldg(post_pk + p);
            if (ap_owner(a) == (uint32_t)w) ap_apply<FIRST>(acc, a, wt);
        }
    }
}

enum { AP_OUT_FUSED = 0, AP_OUT_MAXIMA = 1 };

struct ApproxOut {
    float *maxima;            // MAXIMA: [queries, maxima_stride], column = y * 256 + group
    int64_t maxima_stride;
    const uint64_t *thr_keys; // FUSED: [queries] from kth_of_maxima(positive_floor = true)
    uint64_t *cand;           // [queries, cap] keys of the APPROXIMATE scores
    int32_t *cand_cnt;
    int32_t cap;
    uint32_t n_docs, doc_id_base;
};

// grid (queries, tile groups), query index fastest (co-resident CTAs share a tile's postings through L2), one CTA =
// one query x several doc tiles (stride gridDim.y), one warp = 1024 documents whose accumulators only it touches.
template <int OUT>
__global__ void __launch_bounds__(AP_THREADS, AP_MIN_CTAS)
score_approx_kernel(const uint32_t *__restrict__ post_pk, const uint32_t *__restrict__ blk_ptr,
                    const int32_t *__restrict__ dense_id, const uint32_t *__restrict__ dense_ptr, int n_tiles,
                    const int32_t *__restrict__ q_ptr, const int32_t *__restrict__ q_terms,
                    const float *__restrict__ q_weights, const float *__restrict__ idf, const PackMeta *meta, int q0,
                    int tile_step, int n_y, ApproxOut o) {
    __shared__ __align__(128) float acc[AP_TILE];
    __shared__ __align__(8) uint64_t zbar[AP_WARPS];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ql = blockIdx.x, q = q0 + ql;
    float *acc_w = acc + w * AP_SUB;
#if AP_FORMAT == 2
    static_assert(AP_WARPS == 4, "packed format 2 stores the owning warp in two bits");
    const uint32_t acc_s = ap_smem_u32(acc_w);   // offsets are relative to the owning warp's region
#else
    const uint32_t acc_s = ap_smem_u32(acc);
#endif
    const size_t dense_row = (size_t)n_tiles * B2R_SUBTILES + 1;
    constexpr int SUBS_PER_WARP = B2R_SUBTILES / AP_WARPS;
    const uint32_t zbar_a = ap_smem_u32(&zbar[w]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(zbar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t zphase = 0;
    const int qs = q_ptr[q], qe = q_ptr[q + 1];

    float thr_lo = 0.0f;
    int pos_mode = 0;
    if (OUT == AP_OUT_FUSED) {
        const ApproxBound ab = approx_bound_warp(qs, qe, q_terms, q_weights, idf, meta, o.thr_keys[ql], lane);
        thr_lo = ab.thr_lo;
        pos_mode = ab.pos_mode;
    }

    // a query of <= 32 terms is staged once: lane j keeps term j's f32 weight and the base of its offset row; per
    // tile only the two offsets of the warp's posting range are fetched, one tile ahead
    const bool staged = qe - qs <= 32;
    const uint32_t *my_row = nullptr;
    float my_wt = 0.0f;
    int my_dense = 0;
    uint32_t nxt_beg = 0, nxt_end = 0;
    if (staged && lane < qe - qs) {
        const int t = q_terms[qs + lane];
        my_wt = __double2float_rn(__dmul_rn((double)idf[t], (double)q_weights[qs + lane]));
        const int32_t did = dense_id[t];
        my_dense = did >= 0;
        my_row = my_dense ? dense_ptr + (size_t)did * dense_row + w * SUBS_PER_WARP : blk_ptr + (size_t)t * n_tiles;
        if ((int)blockIdx.y < n_y) {
            const size_t i0 = (size_t)((int)blockIdx.y * tile_step) * (my_dense ? B2R_SUBTILES : 1);
            nxt_beg = my_row[i0];
            nxt_end = my_row[i0 + (my_dense ? SUBS_PER_WARP : 1)];
        }
    }
    for (int y = blockIdx.y; y < n_y; y += gridDim.y) {
        if (lane == 0) {   // clear my 1024 accumulators to -0.0 with one bulk copy; overlaps the staging below
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(zbar_a),
                         "r"((uint32_t)(AP_SUB * 4))
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    ap_smem_u32(acc_w)),
                "l"(g_negzero_page), "r"((uint32_t)(AP_SUB * 4)), "r"(zbar_a)
                : "memory");
        }
        const int tile = y * tile_step;
        uint32_t my_beg = nxt_beg, my_end = nxt_end;
        if (staged && my_row != nullptr && y + (int)gridDim.y < n_y) {
            const size_t i1 = (size_t)((y + (int)gridDim.y) * tile_step) * (my_dense ? B2R_SUBTILES : 1);
            nxt_beg = my_row[i1];
            nxt_end = my_row[i1 + (my_dense ? SUBS_PER_WARP : 1)];
        }
        bool cleared = false, first = true;
        for (int j0 = qs; j0 < qe; j0 += 32) {
            const int nt = min(32, qe - j0);
            if (!staged) {   // long query: lane j stages term j0 + j for this tile
                my_beg = my_end = 0;
                my_dense = 0;
                my_wt = 0.0f;
                if (lane < nt) {
                    const int t = q_terms[j0 + lane];
                    my_wt = __double2float_rn(__dmul_rn((double)idf[t], (double)q_weights[j0 + lane]));
                    const int32_t did = dense_id[t];
                    if (did >= 0) {
                        const uint32_t *row = dense_ptr + (size_t)did * dense_row + (size_t)tile * B2R_SUBTILES +
                                              w * SUBS_PER_WARP;
                        my_beg = row[0];
                        my_end = row[SUBS_PER_WARP];
                        my_dense = 1;
                    } else {
                        const size_t e = (size_t)t * n_tiles + tile;
                        my_beg = blk_ptr[e];
                        my_end = blk_ptr[e + 1];
                    }
                }
            }
            if (!cleared) {
                ap_mbar_wait(zbar_a, zphase);
                cleared = true;
            }
#if AP_BALLOT
            // (summation order is free here, but ascending term id keeps the densest term first: its postings then
            //  take the store-only FIRST path)
            for (unsigned live = __ballot_sync(full, lane < nt && my_beg != my_end); live; live &= live - 1) {
                const int j = __ffs(live) - 1;
                const uint32_t beg = __shfl_sync(full, my_beg, j), end = __shfl_sync(full, my_end, j);
#else
            for (int j = 0; j < nt; ++j) {
                const uint32_t beg = __shfl_sync(full, my_beg, j), end = __shfl_sync(full, my_end, j);
                if (beg == end) continue;   // warp-uniform
#endif
                const int dense = __shfl_sync(full, my_dense, j);
                const float wt = __shfl_sync(full, my_wt, j);
                if (first) ap_apply_term<true>(dense, beg, end, lane, w, post_pk, acc_s, wt);
                else ap_apply_term<false>(dense, beg, end, lane, w, post_pk, acc_s, wt);
                first = false;
                __syncwarp();
            }
        }
        if (!cleared) ap_mbar_wait(zbar_a, zphase);
        zphase ^= 1;

        const uint32_t doc0 = (uint32_t)tile * (uint32_t)AP_TILE + (uint32_t)w * (uint32_t)AP_SUB;
        if (OUT == AP_OUT_MAXIMA) {
            // two group maxima per lane (16 documents each); -0.0 is folded into +0.0, NaNs never win
            constexpr int GPL = AP_SUB / 512;   // groups per lane: 16 documents each, 256 groups per tile
            float m[GPL];
#pragma unroll
            for (int g = 0; g < GPL; ++g) {
                float mm = __int_as_float(0xff800000);
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int i = (g * 4 + it) * 128 + lane * 4;
                    const float4 a = *reinterpret_cast<const float4 *>(acc_w + i);
                    const uint32_t doc = doc0 + i;
                    if (doc + 3 < o.n_docs) {
                        mm = fmaxf(mm, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
                    } else {
                        if (doc < o.n_docs) mm = fmaxf(mm, a.x);
                        if (doc + 1 < o.n_docs) mm = fmaxf(mm, a.y);
                        if (doc + 2 < o.n_docs) mm = fmaxf(mm, a.z);
                    }
                }
                m[g] = __fadd_rn(mm, 0.0f);
            }
            float *out = o.maxima + (int64_t)ql * o.maxima_stride + (int64_t)y * AP_GROUPS_PER_TILE + w * (32 * GPL) +
                         lane;
#pragma unroll
            for (int g = 0; g < GPL; ++g) out[32 * g] = m[g];
        } else {
#pragma unroll 2
            for (int it = 0; it < AP_SUB / 128; ++it) {
                const int i = it * 128 + lane * 4;
                const float4 a = *reinterpret_cast<const float4 *>(acc_w + i);
                const float mx = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
                bool look = mx >= thr_lo;
                if (pos_mode && look) {   // thr_lo <= 0 admits cleared accumulators: look only at quads with a touched one
                    const uint32_t b0 = __float_as_uint(a.x), b1 = __float_as_uint(a.y), b2 = __float_as_uint(a.z),
                                   b3 = __float_as_uint(a.w);
                    look = (b0 | b1 | b2 | b3) != AP_NEGZERO || (b0 & b1 & b2 & b3) != AP_NEGZERO;
                }
                if (look) {
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t doc = doc0 + i + c;
                        if (av[c] >= thr_lo && __float_as_uint(av[c]) != AP_NEGZERO && doc < o.n_docs) {
                            const uint64_t key = make_key(ord_f32(av[c]), o.doc_id_base + doc);
                            const int slot = atomicAdd(o.cand_cnt + ql, 1);
                            if (slot < o.cap) o.cand[(int64_t)ql * o.cap + slot] = key;
                        }
                    }
                }
            }
        }
        // the next tile's bulk clear (async proxy) must not overtake this tile's accumulator reads (generic proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
}

// ---- selection + exact rescoring ------------------------------------------------------------------------------------
__device__ __forceinline__ void ap_select_bin(const uint32_t *hist, uint32_t kk, uint32_t *wsum, uint32_t *bin,
                                              uint32_t *kk_in_bin) {   // 256 threads; thread t owns bin 255 - t
    const int tid = threadIdx.x;
    const uint32_t v = hist[255 - tid];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += u;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int ww = 0; ww < (tid >> 5); ++ww) base += wsum[ww];
    inc += base;
    const uint32_t above = inc - v;
    if ((above < kk && inc >= kk) || (tid == 255 && inc < kk)) {
        *bin = 255u - (uint32_t)tid;
        *kk_in_bin = kk - above;
    }
    __syncthreads();
}

// exact score of local document d for query terms [qs, qe): the reference's f64 chain, terms ascending (one warp)
__device__ __forceinline__ double ap_exact_score(uint32_t d, int qs, int qe, const int32_t *__restrict__ q_terms,
                                                 const float *__restrict__ q_weights, const float *__restrict__ idf,
                                                 const uint32_t *__restrict__ post_doc,
                                                 const double *__restrict__ post_val,
                                                 const uint32_t *__restrict__ blk_ptr,
                                                 const int32_t *__restrict__ dense_id,
                                                 const uint32_t *__restrict__ dense_ptr, int n_tiles, int lane) {
    const unsigned full = 0xffffffffu;
    const uint32_t tile = d / AP_TILE, sub = (d % AP_TILE) / (AP_TILE / B2R_SUBTILES);
    const size_t dense_row = (size_t)n_tiles * B2R_SUBTILES + 1;
    double acc = 0.0;
    for (int j = qs; j < qe; ++j) {
        const int t = __ldg(q_terms + j);
        const int32_t did = __ldg(dense_id + t);
        uint32_t beg, end;
        if (did >= 0) {
            const uint32_t *row = dense_ptr + (size_t)did * dense_row + (size_t)tile * B2R_SUBTILES + sub;
            beg = __ldg(row);
            end = __ldg(row + 1);
        } else {
            const size_t e = (size_t)t * n_tiles + tile;
            beg = __ldg(blk_ptr + e);
            end = __ldg(blk_ptr + e + 1);
        }
        for (uint32_t p0 = beg; p0 < end; p0 += 32) {   // warp-uniform bounds
            const uint32_t p = p0 + lane;
            const unsigned hit = __ballot_sync(full, p < end && __ldg(post_doc + p) == d);
            if (hit) {
                const double u = __ldg(post_val + p0 + (__ffs(hit) - 1));
                acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)__ldg(idf + t), u), (double)__ldg(q_weights + j)));
                break;
            }
        }
    }
    return acc;
}

__global__ void __launch_bounds__(256, 8)   // 8 x 148 CTAs resident: a 1024-query batch is one wave
approx_select_kernel(const uint32_t *__restrict__ post_doc, const double *__restrict__ post_val,
                     const uint32_t *__restrict__ blk_ptr, const int32_t *__restrict__ dense_id,
                     const uint32_t *__restrict__ dense_ptr, int n_tiles, const int32_t *__restrict__ q_ptr,
                     const int32_t *__restrict__ q_terms, const float *__restrict__ q_weights,
                     const float *__restrict__ idf, const PackMeta *meta, int q0, const uint64_t *__restrict__ thr_keys,
                     const uint64_t *__restrict__ lists, int cap, int32_t *__restrict__ cnt, int k,
                     uint32_t doc_id_base, uint64_t *__restrict__ out, int64_t *__restrict__ idx_out,
                     float *__restrict__ val_out, int32_t *__restrict__ marked) {
    extern __shared__ uint64_t arr[];   // [cap]
    __shared__ uint64_t surv[AP_SURV_MAX];
    __shared__ uint32_t hist[256], wsum[8], s_bin, s_kk, s_n;
    __shared__ double s_dp, s_c2;
    __shared__ const uint32_t *st_row[AP_ST_TERMS];   // per query term: its offset row (dense: per sub-tile, else per tile)
    __shared__ double st_idf[AP_ST_TERMS], st_qw[AP_ST_TERMS];
    __shared__ int st_dense[AP_ST_TERMS];
    __shared__ int s_pos, s_ok;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int q = q0 + row;
    const int qs = q_ptr[q], qe = q_ptr[q + 1];
    const int c = cnt[row];
    if (qe - qs <= AP_ST_TERMS && tid >= 32 && tid - 32 < qe - qs) {   // stage the query's terms (warps 1..2)
        const int j = tid - 32;
        const int t = q_terms[qs + j];
        const int32_t did = dense_id[t];
        st_dense[j] = did >= 0;
        st_row[j] = did >= 0 ? dense_ptr + (size_t)did * ((size_t)n_tiles * B2R_SUBTILES + 1) : blk_ptr + (size_t)t * n_tiles;
        st_idf[j] = (double)idf[t];
        st_qw[j] = (double)q_weights[qs + j];
    }
    if (w == 0) {
        const ApproxBound ab = approx_bound_warp(qs, qe, q_terms, q_weights, idf, meta, thr_keys[row], lane);
        if (lane == 0) {
            s_dp = ab.dp;
            s_c2 = ab.c2;
            s_pos = ab.pos_mode;
            s_ok = ab.ok;
        }
    }
    __syncthreads();
    // overflowed (c > cap: already marked), short, or no usable bound: the exhaustive exact fallback takes the query
    auto mark = [&]() {   // thread 0: hand the query to the exhaustive fallback (cnt > cap is its marker, as for an overflow)
        cnt[row] = cap + 1;
        marked[1 + atomicAdd(marked, 1)] = row;
    };
    if (c > cap || c < k || !s_ok) {
        if (tid == 0) mark();
        return;
    }
    const uint64_t *src = lists + (int64_t)row * cap;
    for (int i = tid; i < c; i += 256) arr[i] = src[i];
    // k-th largest approximate SCORE, counted with multiplicity (radix select, 4 x 8 bits on the ordered encoding):
    // the survivors are defined by a score threshold, so the document bits of the keys play no part
    uint32_t prefix = 0, mask = 0, kk = (uint32_t)k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < c; i += 256) {
            const uint32_t o = (uint32_t)(arr[i] >> 32);
            if ((o & mask) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
        }
        __syncthreads();
        ap_select_bin(hist, kk, wsum, &s_bin, &s_kk);
        prefix |= s_bin << shift;
        kk = s_kk;
        mask |= 255u << shift;
    }
    // k candidates have approx >= A_k, hence exact >= G = g(A_k): a member of the top-k has exact >= G - slack (f32
    // rounding of the ranked score) and therefore h(approx) >= G - slack, i.e. approx >= h^-1(G - slack).  Compared in
    // the ordered encoding, which is monotone.
    const double a_k = (double)unord_f32(prefix);
    const double G = a_k - s_dp * fabs(a_k) - s_c2;
    const double yv = G - fabs(G) * 0x1p-22 - 0x1p-130;
    double x0 = yv >= s_c2 ? (yv - s_c2) / (1.0 + s_dp) : (yv - s_c2) / (1.0 - s_dp);
    x0 -= fabs(x0) * 0x1p-20;
    const uint32_t lo_ord = ord_f32(__double2float_rd(x0));
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < c; i += 256) {
        const uint64_t key = arr[i];
        if ((uint32_t)(key >> 32) >= lo_ord) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            if (slot < (uint32_t)AP_SURV_MAX) surv[slot] = key;
        }
    }
    __syncthreads();
    const int n = (int)s_n;
    if (n > AP_SURV_MAX || n < k) {   // (n < k cannot happen: the k best approximate keys survive)
        if (tid == 0) mark();
        return;
    }
    // exact rescoring.  Queries of <= AP_ST_TERMS terms: one warp per (survivor, term) pair looks the posting up and
    // writes the term's contribution (idf * u) * qtf -- +0.0 when the document does not hold the term, which leaves an
    // f64 sum that started at +0.0 unchanged bit for bit -- into the shared-memory area the candidate keys no longer
    // need; one thread per survivor then adds them in ascending term id, the reference's summation order.  The
    // look-ups are dependent loads: 8 warps x (n * n_terms) short chains instead of n long ones.
    const int nt = qe - qs;
    if (nt <= AP_ST_TERMS && nt <= cap) {
        double *contrib = reinterpret_cast<double *>(arr);   // [batch, nt], batch * nt <= cap
        const int batch = cap / (nt > 0 ? nt : 1);
        for (int b0 = 0; b0 < n; b0 += batch) {
            const int nb = min(batch, n - b0);
            __syncthreads();   // (the keys in arr / the previous batch's contributions are no longer read)
            for (int task = w; task < nb * nt; task += 8) {
                const int si = task / nt, j = task - si * nt;
                const uint32_t d = (0xFFFFFFFFu - (uint32_t)surv[b0 + si]) - doc_id_base;
                const uint32_t tile = d / AP_TILE, sub = (d % AP_TILE) / (AP_TILE / B2R_SUBTILES);
                const uint32_t *row = st_row[j] + (st_dense[j] ? (size_t)tile * B2R_SUBTILES + sub : (size_t)tile);
                const uint32_t beg = __ldg(row), end = __ldg(row + 1);
                uint32_t pos = 0xFFFFFFFFu;
                for (uint32_t p0 = beg; p0 < end && pos == 0xFFFFFFFFu; p0 += 128) {   // warp-uniform
                    uint32_t v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t p = p0 + 32 * i + lane;
                        v[i] = p < end ? __ldg(post_doc + p) : 0xFFFFFFFFu;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const unsigned hit = __ballot_sync(0xffffffffu, v[i] == d);
                        if (hit && pos == 0xFFFFFFFFu) pos = p0 + 32 * i + (__ffs(hit) - 1);
                    }
                }
                if (lane == 0)
                    contrib[task] = pos != 0xFFFFFFFFu
                                        ? __dmul_rn(__dmul_rn(st_idf[j], __ldg(post_val + pos)), st_qw[j])
                                        : 0.0;
            }
            __syncthreads();
            for (int i = tid; i < nb; i += 256) {
                double acc = 0.0;
                for (int j = 0; j < nt; ++j) acc = __dadd_rn(acc, contrib[i * nt + j]);
                const uint32_t gid = 0xFFFFFFFFu - (uint32_t)surv[b0 + i];
                surv[b0 + i] = make_key(ord_f32(__double2float_rn(acc)), gid);
            }
        }
    } else {   // long queries: one warp per survivor walks the terms
        for (int i = w; i < n; i += 8) {
            const uint32_t gid = 0xFFFFFFFFu - (uint32_t)surv[i];
            const uint32_t d = gid - doc_id_base;
            const double s = ap_exact_score(d, qs, qe, q_terms, q_weights, idf, post_doc, post_val, blk_ptr, dense_id,
                                            dense_ptr, n_tiles, lane);
            __syncwarp();
            if (lane == 0) surv[i] = make_key(ord_f32(__double2float_rn(s)), gid);
        }
    }
    int P = 32;
    while (P < n) P <<= 1;
    __syncthreads();
    for (int i = n + tid; i < P; i += 256) surv[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc<256>(surv, P);
    // positive mode: valid only if the exact k-th best is > 0 (then no untouched / non-positive document can be in
    // the top-k); otherwise the exhaustive fallback decides
    if (s_pos && (uint32_t)(surv[k - 1] >> 32) <= 0x80000000u) {
        if (tid == 0) mark();
        return;
    }
    for (int i = tid; i < k; i += 256) {
        const uint64_t key = surv[i];
        const int64_t at = (int64_t)row * k + i;
        if (out) out[at] = key;
        if (idx_out) idx_out[at] = key ? (int64_t)(0xFFFFFFFFu - (uint32_t)key) : -1;
        if (val_out) val_out[at] = key ? unord_f32((uint32_t)(key >> 32)) : __int_as_float(0xff800000);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
static const PackMeta *meta_of(const b2r_index *ix) {
    return reinterpret_cast<const PackMeta *>(reinterpret_cast<const char *>(ix->post_pk) + pack_meta_offset(ix->nnz));
}

bool approx_usable(const b2r_index *ix, int k) {
    return g_approx_enabled && ix->post_pk != nullptr && ix->kind == B2R_KIND_BM25 && ix->tile_docs == AP_TILE &&
           k >= 1 && k <= AP_SURV_MAX / 2;
}

// doc tiles walked by one CTA (B2R_AP_TILES_PER_CTA overrides it: tuning experiments only)
static const int g_ap_tiles_per_cta = [] {
    const char *e = getenv("B2R_AP_TILES_PER_CTA");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 64 ? v : 4;
}();

static int approx_grid_y(int nq, int n_y) {
    // fewer tiles per CTA while the grid would not fill the GPU a few times over (148 SMs x 12 CTAs)
    int per_cta = g_ap_tiles_per_cta;
    while (per_cta > 1 && (int64_t)nq * (n_y / per_cta) < 148 * AP_MIN_CTAS * 4) per_cta >>= 1;
    return (n_y + per_cta - 1) / per_cta;
}

int approx_maxima(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                  const float *idf, int q0, int nq, int tile_step, int n_sample, float *maxima, int64_t maxima_stride,
                  cudaStream_t st) {
    if (nq == 0 || n_sample == 0) return B2R_OK;
    ApproxOut o = {};
    o.maxima = maxima;
    o.maxima_stride = maxima_stride;
    o.n_docs = (uint32_t)ix->n_docs;
    dim3 grid((unsigned)nq, (unsigned)approx_grid_y(nq, n_sample));
    score_approx_kernel<AP_OUT_MAXIMA><<<grid, AP_THREADS, 0, st>>>(ix->post_pk, ix->blk_ptr, ix->dense_id,
                                                                    ix->dense_ptr, ix->n_tiles, q_ptr, q_terms,
                                                                    q_weights, idf, meta_of(ix), q0, tile_step,
                                                                    n_sample, o);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

int approx_fused(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                 const float *idf, int q0, int nq, const uint64_t *thr, uint64_t *cand, int32_t *cand_cnt, int cap,
                 cudaStream_t st) {
    if (nq == 0) return B2R_OK;
    ApproxOut o = {};
    o.thr_keys = thr;
    o.cand = cand;
    o.cand_cnt = cand_cnt;
    o.cap = cap;
    o.n_docs = (uint32_t)ix->n_docs;
    o.doc_id_base = (uint32_t)ix->doc_id_base;
    dim3 grid((unsigned)nq, (unsigned)approx_grid_y(nq, ix->n_tiles));
    score_approx_kernel<AP_OUT_FUSED><<<grid, AP_THREADS, 0, st>>>(ix->post_pk, ix->blk_ptr, ix->dense_id,
                                                                   ix->dense_ptr, ix->n_tiles, q_ptr, q_terms,
                                                                   q_weights, idf, meta_of(ix), q0, 1, ix->n_tiles, o);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

int approx_select(const b2r_index *ix, const int32_t *q_ptr, const int32_t *q_terms, const float *q_weights,
                  const float *idf, int q0, int nq, const uint64_t *thr, const uint64_t *cand, int32_t *cand_cnt,
                  int cap, int k, uint64_t *keys_out, int64_t *idx_out, float *val_out, int32_t *marked,
                  cudaStream_t st) {
    if (nq == 0) return B2R_OK;
    const size_t smem = (size_t)cap * 8;
    if (smem > 32 * 1024)
        B2R_CUDA(cudaFuncSetAttribute(approx_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    approx_select_kernel<<<(unsigned)nq, 256, smem, st>>>(
        ix->post_doc, static_cast<const double *>(ix->post_val), ix->blk_ptr, ix->dense_id, ix->dense_ptr, ix->n_tiles,
        q_ptr, q_terms, q_weights, idf, meta_of(ix), q0, thr, cand, cap, cand_cnt, k, (uint32_t)ix->doc_id_base,
        keys_out, idx_out, val_out, marked);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

}  // namespace b2r

using namespace b2r;

extern "C" void b2r_set_approx_prefilter(int enabled) { b2r::g_approx_enabled = enabled != 0; }

extern "C" size_t b2r_index_pack_bytes(int64_t nnz) { return pack_meta_offset(nnz > 0 ? nnz : 0) + 256; }

extern "C" int b2r_index_pack(const b2r_index *ix, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(ix && ix->post_doc && ix->post_val && ix->post_pk, "b2r_index_pack: index (or post_pk) missing");
    B2R_CHECK_ARG(ix->kind == B2R_KIND_BM25 && ix->tile_docs == AP_TILE,
                  "b2r_index_pack: the packed copy exists for BM25 indexes with 4096-document tiles only");
    PackMeta *meta = const_cast<PackMeta *>(meta_of(ix));
    B2R_CUDA(cudaMemsetAsync(meta, 0, 256, st));
    const int64_t want = (ix->nnz + 256 * 8 - 1) / (256 * 8);
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    pack_postings_kernel<<<blocks, 256, 0, st>>>(ix->post_doc, static_cast<const double *>(ix->post_val), ix->nnz,
                                                 ix->post_pk, meta);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

extern "C" int b2r_index_pack_status(const b2r_index *ix, void *stream, float *u_max_out) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(ix && ix->post_pk, "b2r_index_pack_status: no packed copy");
    PackMeta h = {};
    B2R_CUDA(cudaMemcpyAsync(&h, meta_of(ix), sizeof(h), cudaMemcpyDeviceToHost, st));
    B2R_CUDA(cudaStreamSynchronize(st));
    if (u_max_out) {
        float f;
        memcpy(&f, &h.max_bits, 4);
        *u_max_out = f;
    }
    if (h.bad) {
        set_error("b2r_index_pack: a posting value is not finite; the approximate pre-filter is off for this index");
        return B2R_ERR_UNSUPPORTED;
    }
    return B2R_OK;
}
