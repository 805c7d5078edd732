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
//
// Three epilogues.  DENSE writes the f32 score vector (the reference's return value).  MAXIMA and FUSED are
// the search path, which never materialises scores: (1) every 64th / 16th doc tile is scored with the
// MAXIMA epilogue, which emits one maximum per lane (a group of <= tile_docs/256 documents); the k-th largest
// of those group maxima is a lower bound T of the k-th best score of the whole shard (k groups, hence k
// documents, reach it).  (2) ALL tiles are scored with the FUSED epilogue, which appends the few documents
// whose key beats T to a per-query candidate list; the exact top-k is the top-k of that list.  A query whose
// list overflows is rescored exhaustively by the gated DENSE + top-k kernels that follow (they exit at once
// for every other query), so the result is exact for any corpus order; the gate is evaluated on the device,
// nothing synchronises.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

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

__device__ __forceinline__ double2 ld_stream_v2f64(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream_v2f32(const float2 *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

constexpr int SC_THREADS = 32 * B2R_SUBTILES;  // one warp per sub-tile of the CTA's doc tile

__device__ __forceinline__ uint32_t sc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sc_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tSC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra SC_DONE;\n\tbra SC_WAIT;\n\tSC_DONE:\n\t}\n"
        :
        : "r"(bar), "r"(parity)
        : "memory");
}

// acc[rel] += contribution of one posting (reference order of operations, no FMA contraction).
// FIRST: this is the first term that touches the warp's (freshly cleared) sub-tile, so the accumulator is
// known to be +0.0 and is not read back: a store instead of a read-modify-write on the shared-memory pipe.
// `0.0 + x` is still evaluated so that the stored bits are those of the reference's `acc += x`.
template <int KIND, bool FIRST>
__device__ __forceinline__ void apply_posting(double *acc_w, uint32_t rel, double u_or_w, float w_idf, float w_q,
                                              double w_idf64, double w_q64) {  // BM25 uses the doubles, IMPACT the floats
    const double a = FIRST ? 0.0 : acc_w[rel];
    if (KIND == B2R_KIND_BM25) {
        acc_w[rel] = __dadd_rn(a, __dmul_rn(__dmul_rn(w_idf64, u_or_w), w_q64));
    } else {
        // reference (fastmath) evaluates (tf * qtf) * idf in f32, then widens: see oracle/np_oracle.py
        float c = __fmul_rn(__fmul_rn((float)u_or_w, w_q), w_idf);
        acc_w[rel] = __dadd_rn(a, (double)c);
    }
}

template <int KIND>
__device__ __forceinline__ double load_val(const void *post_val, uint32_t p) {
    if (KIND == B2R_KIND_BM25) return ld_stream_f64(static_cast<const double *>(post_val) + p);
    return (double)ld_stream_f32(static_cast<const float *>(post_val) + p);
}

// type in which a staged term's idf and query weight travel between lanes
template <int KIND>
struct TermW {
    using type = float;
};
#ifndef SC_WT64
#define SC_WT64 0  // 1: BM25 term weights are widened to f64 once per lane at staging (needs > 40 registers: spills at 6 CTAs/SM)
#endif
#if SC_WT64
template <>
struct TermW<B2R_KIND_BM25> {
    using type = double;
};
#endif

// One DENSE (non-slab) term applied by one warp to its own sub-tile: [beg, end) are exactly the warp's postings.
template <int KIND, bool FIRST>
__device__ __forceinline__ void apply_dense_term(uint32_t beg, uint32_t end, int lane, uint32_t my_doc0,
                                                 const uint32_t *__restrict__ post_doc, const void *__restrict__ post_val,
                                                 double *acc_w, typename TermW<KIND>::type w_idf_t,
                                                 typename TermW<KIND>::type w_q_t) {
    // BM25 multiplies in f64: the widening is done once per term and warp (f32 -> f64 conversions run on a slow pipe)
    const double w_idf64 = (double)w_idf_t, w_q64 = (double)w_q_t;
    const float w_idf = (float)w_idf_t, w_q = (float)w_q_t;  // (exact: the values are f32 numbers)
    uint32_t p = beg + lane;
    for (; p + 96 < end; p += 128) {  // 4 independent postings in flight per lane
        uint32_t d0 = ld_stream_u32(post_doc + p), d1 = ld_stream_u32(post_doc + p + 32);
        uint32_t d2 = ld_stream_u32(post_doc + p + 64), d3 = ld_stream_u32(post_doc + p + 96);
        double u0 = load_val<KIND>(post_val, p), u1 = load_val<KIND>(post_val, p + 32);
        double u2 = load_val<KIND>(post_val, p + 64), u3 = load_val<KIND>(post_val, p + 96);
        apply_posting<KIND, FIRST>(acc_w, d0 - my_doc0, u0, w_idf, w_q, w_idf64, w_q64);
        apply_posting<KIND, FIRST>(acc_w, d1 - my_doc0, u1, w_idf, w_q, w_idf64, w_q64);
        apply_posting<KIND, FIRST>(acc_w, d2 - my_doc0, u2, w_idf, w_q, w_idf64, w_q64);
        apply_posting<KIND, FIRST>(acc_w, d3 - my_doc0, u3, w_idf, w_q, w_idf64, w_q64);
    }
    for (; p < end; p += 32) {
        uint32_t d = ld_stream_u32(post_doc + p);
        double u = load_val<KIND>(post_val, p);
        apply_posting<KIND, FIRST>(acc_w, d - my_doc0, u, w_idf, w_q, w_idf64, w_q64);
    }
}

// ------------------------------------------------------------------------------------------------------------
// The scorer: TILE-STATIONARY.  One CTA = one doc tile x a range of queries; it first copies the tile's head-term
// slabs (include/b200ret.h: the sub-tile's posting values in document order, 0 = no posting) into shared memory --
// once per CTA, i.e. once per (tile, query range) instead of once per (tile, query) -- and then its warps walk the
// queries of the range.  A warp owns one sub-tile of tile_docs/8 documents of one QUERY SLOT (a CTA runs up to
// four queries side by side, 8 warps each); warps never synchronise with each other after the slab copy.
// Per (query, sub-tile) the warp applies the query's terms in ascending term id (the reference's summation order:
// CSR rows are sorted by term id), each term in one of three ways:
//   slab    the term is a head term and its segment of this sub-tile has a slab: lane l owns documents
//           {64 c + 2 l, 64 c + 2 l + 1} and keeps their accumulators in REGISTERS; a posting costs one 8-byte
//           shared-memory read and three f64 operations -- no document id, no read-modify-write, no L2 traffic.
//           r += (idf * 0) * q leaves r bit-identical (r is never -0.0: it starts at +0.0), so absent documents
//           need no mask; terms with a non-finite weight never take this path.
//   dense   the term has sub-tile offsets (dense_ptr): the warp streams its own bank-scheduled segment from L2 and
//           does read-modify-write on the sub-tile's accumulators in shared memory;
//   sparse  every warp scans the tile's small block and keeps the postings of its own doc range.
// The accumulators move between registers and shared memory only when they have to: a run of slab terms stays in
// registers, a sparse block without a posting in the warp's doc range does not disturb them, and a sub-tile whose
// last applied term was a slab runs its epilogue straight from the registers.  Shared-memory accumulators are
// zero-filled lazily (only sub-tiles that see a non-slab posting ever need them).
// Why: with one CTA per (query, tile) every query re-fetched the head terms' postings from L2 (26 GB of L2->SM
// traffic per 1024-query step against ~12 TB/s of fabric: the round-1 kernel's real bound).
enum { SC_OUT_DENSE = 0, SC_OUT_FUSED = 1, SC_OUT_MAXIMA = 2 };
enum { SC_TILES_ALL = 0, SC_TILES_SAMPLE = 1 };
constexpr int SC_GROUPS_PER_TILE = 32 * B2R_SUBTILES;  // MAXIMA: one group maximum per lane
#ifndef SC_MAX_SLOTS_DEF
#define SC_MAX_SLOTS_DEF 4
#endif
constexpr int SC_MAX_SLOTS = SC_MAX_SLOTS_DEF;         // queries a CTA works on side by side
constexpr int SC_REC_CAP = 2048;                       // term records of a CTA's query range staged in shared memory
constexpr int SC_SLAB_SUB = B2R_SLAB_TILE_DOCS / B2R_SUBTILES;   // 256 documents per slab
constexpr int SC_SLAB_PAIRS = SC_SLAB_SUB / 64;        // double2 accumulators per lane in the register layout

struct ScoreOut {
    // DENSE: scores[queries, scores_stride], column = out_tile * tile_docs + doc in tile
    // MAXIMA: scores[queries, scores_stride], column = out_tile * SC_GROUPS_PER_TILE + warp * 32 + lane
    float *scores;
    int64_t scores_stride;
    const int32_t *gate;    // optional: run only for queries with gate[q_local] > gate_cap
    int32_t gate_cap;
    // FUSED
    const uint64_t *thr_keys;  // [queries]: keep documents whose key is > thr_keys[q] (0 = no threshold)
    uint64_t *cand;         // [queries, cap]
    int32_t *cand_cnt;      // [queries]
    int32_t cap;
    uint32_t n_docs;        // documents in this shard (tail of the last tile is padding)
    uint32_t doc_id_base;
};

template <int KIND, int OUT, bool SLABS>
__global__ void __launch_bounds__(SC_MAX_SLOTS * SC_THREADS, 1)
score_tiles_kernel(const uint32_t *__restrict__ post_doc, const void *__restrict__ post_val,
                   const uint32_t *__restrict__ blk_ptr, const int32_t *__restrict__ dense_id,
                   const uint32_t *__restrict__ dense_ptr, const int32_t *__restrict__ slab_idx,
                   const void *__restrict__ slab_val, int n_slabs, int n_tiles, int tile_docs,
                   const int32_t *__restrict__ q_ptr, const int32_t *__restrict__ q_terms,
                   const float *__restrict__ q_weights, const float *__restrict__ idf, int q0, int nq, int range_len,
                   int tile_mode, int tile_step, int diag, ScoreOut o) {
    using val_t = typename std::conditional<KIND == B2R_KIND_BM25, double, float>::type;
    using val2_t = typename std::conditional<KIND == B2R_KIND_BM25, double2, float2>::type;
    extern __shared__ __align__(128) unsigned char sc_smem[];
    __shared__ __align__(8) uint64_t cbar;              // "the slabs of this tile have landed"
    __shared__ int cache_ok[B2R_HEAD_TERMS * B2R_SUBTILES];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slots = blockDim.x / SC_THREADS;
    const int slot = warp / B2R_SUBTILES, w = warp % B2R_SUBTILES;
    const int sub = tile_docs / B2R_SUBTILES;
    const int y = blockIdx.y;
    const int tile = tile_mode == SC_TILES_ALL ? y : y * tile_step;
    const size_t n_seg = (size_t)n_tiles * B2R_SUBTILES;
    const size_t dense_row = n_seg + 1;
    const int r0 = blockIdx.x * range_len, r1 = min(nq, r0 + range_len);
    // layout: accumulators f64[slots][tile_docs], (SLABS) the slab cache val_t[B2R_HEAD_TERMS][tile_docs], then the
    // term records of the CTA's query range int4[SC_REC_CAP] = {idf, query weight, term id, dense row}
    double *acc_w = reinterpret_cast<double *>(sc_smem) + (size_t)slot * tile_docs + (size_t)w * sub;
    const val_t *cache = reinterpret_cast<const val_t *>(sc_smem + (size_t)slots * tile_docs * sizeof(double));
    int4 *rec = reinterpret_cast<int4 *>(sc_smem + (size_t)slots * tile_docs * sizeof(double) +
                                         (SLABS ? (size_t)B2R_HEAD_TERMS * tile_docs * sizeof(val_t) : 0));

    if (OUT == SC_OUT_DENSE && o.gate != nullptr) {   // a gated launch is expected to find nothing to do
        int any = 0;
        for (int ql = r0 + (int)threadIdx.x; ql < r1; ql += blockDim.x) any |= o.gate[ql] > o.gate_cap;
        if (!__syncthreads_or(any)) return;
    }
    // The term records of the range are staged ONCE per CTA (every (query, sub-tile) task would otherwise walk the
    // dependent chain q_terms -> idf / dense_id on its own, four exposed L2 latencies per task).
    const int tbase = q_ptr[q0 + r0];
    {
        const int n_rec = min(q_ptr[q0 + r1] - tbase, SC_REC_CAP);
        for (int i = threadIdx.x; i < n_rec; i += blockDim.x) {
            const int t = q_terms[tbase + i];
            rec[i] = make_int4(__float_as_int(idf[t]), __float_as_int(q_weights[tbase + i]), t, dense_id[t]);
        }
    }
    if (!SLABS) __syncthreads();
    if (SLABS) {
        const uint32_t cbar_a = sc_smem_u32(&cbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cbar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t bytes = 0;
            for (int i = lane; i < B2R_HEAD_TERMS * B2R_SUBTILES; i += 32) {
                const int h = i / B2R_SUBTILES, sseg = i % B2R_SUBTILES;
                const int32_t sid = slab_idx[(size_t)h * n_seg + (size_t)tile * B2R_SUBTILES + sseg];
                const int ok = sid >= 0 && sid < n_slabs;
                cache_ok[i] = ok;
                if (ok) {
                    const uint32_t nb = (uint32_t)(SC_SLAB_SUB * sizeof(val_t));
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            sc_smem_u32(cache + (size_t)h * tile_docs + (size_t)sseg * SC_SLAB_SUB)),
                        "l"(static_cast<const val_t *>(slab_val) + (size_t)sid * SC_SLAB_SUB), "r"(nb), "r"(cbar_a)
                        : "memory");
                    bytes += nb;
                }
            }
#pragma unroll
            for (int ofs = 16; ofs; ofs >>= 1) bytes += __shfl_xor_sync(full, bytes, ofs);
            // (a copy that completes before this arrive only drives the transaction count negative for a moment:
            // the phase cannot complete before the one pending arrival below)
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cbar_a), "r"(bytes) : "memory");
        }
        __syncthreads();                 // cache_ok is visible
        sc_mbar_wait(cbar_a, 0);         // the slab bytes are visible
    }

    auto load_rec = [&](int j) -> int4 {   // record of term j (index into q_terms)
        if (j - tbase < SC_REC_CAP) return rec[j - tbase];
        const int t = q_terms[j];
        return make_int4(__float_as_int(idf[t]), __float_as_int(q_weights[j]), t, dense_id[t]);
    };
    auto offsets_of = [&](const int4 rc, uint32_t &beg, uint32_t &end) {   // postings of the term in my sub-tile / tile
        if (rc.w >= 0) {
            const uint32_t *row = dense_ptr + (size_t)rc.w * dense_row + (size_t)tile * B2R_SUBTILES + w;
            beg = row[0];
            end = row[1];
        } else {
            const size_t e = (size_t)rc.z * n_tiles + tile;
            beg = blk_ptr[e];
            end = blk_ptr[e + 1];
        }
    };
    // software pipeline over the warp's queries: the offsets of the NEXT query's (first 32) terms and the q_ptr pair of
    // the one after it are in flight while the current query is applied
    int qs = 0, qe = 0, nqs = 0, nqe = 0;
    uint32_t pbeg = 0, pend = 0;
    if (r0 + slot < r1) {
        qs = q_ptr[q0 + r0 + slot];
        qe = q_ptr[q0 + r0 + slot + 1];
        if (lane < min(32, qe - qs)) offsets_of(load_rec(qs + lane), pbeg, pend);
    }
    if (r0 + slot + slots < r1) {
        nqs = q_ptr[q0 + r0 + slot + slots];
        nqe = q_ptr[q0 + r0 + slot + slots + 1];
    }
    for (int ql = r0 + slot; ql < r1; ql += slots) {
        uint32_t nbeg = 0, nend = 0;
        int nnqs = 0, nnqe = 0;
        if (ql + slots < r1 && lane < min(32, nqe - nqs)) offsets_of(load_rec(nqs + lane), nbeg, nend);
        if (ql + 2 * slots < r1) {
            nnqs = q_ptr[q0 + ql + 2 * slots];
            nnqe = q_ptr[q0 + ql + 2 * slots + 1];
        }
        if (!(OUT == SC_OUT_DENSE && o.gate != nullptr && o.gate[ql] <= o.gate_cap)) {
        const uint32_t my_doc0 = (uint32_t)tile * (uint32_t)tile_docs + (uint32_t)w * (uint32_t)sub;
        // FUSED: the candidate threshold of this query
        uint64_t thr = 0;
        double thr_lo = 0.0;
        if (OUT == SC_OUT_FUSED) {
            thr = o.thr_keys[ql];
            const uint32_t thr_hi = (uint32_t)(thr >> 32);
            // A document can only beat thr if f32(acc) >= the threshold score.  Instead of converting all
            // accumulators of the tile, they are compared in f64 against the f32 value just below the threshold
            // score: acc < pred(thr_f) implies f32(acc) <= pred(thr_f) < thr_f (rounding is monotone).  Only
            // survivors are converted and keyed.  thr_hi == 0: no threshold yet.
            // ordered encoding: -1 = next smaller f32; 0x7fffffff would be -0.0 (== +0.0 in the ranking): skip it;
            // at or below -inf (0x007fffff) there is nothing smaller: no filter
            uint32_t thr_ord = thr_hi > 0x007fffffu ? thr_hi - 1u : 0u;
            if (thr_ord == 0x7fffffffu) thr_ord = 0x7ffffffeu;
            thr_lo = thr_ord ? (double)unord_f32(thr_ord) : -__longlong_as_double(0x7ff0000000000000ll);
            // "strictly positive scores only" (kth_of_maxima's positive floor): nothing below 2^-150 rounds to a
            // positive f32, so the untouched documents (acc == 0) never reach the conversion path
            if (thr == ((0x80000000ull << 32) | 0xFFFFFFFFull)) thr_lo = __longlong_as_double(0x3690000000000000ll);
        }

        // where the sub-tile's accumulators are (warp-uniform): nowhere yet (all +0.0), registers, or shared memory
        bool first = true;   // no term has touched this sub-tile yet
        bool in_reg = false;
        double2 r[SC_SLAB_PAIRS];
        auto to_smem = [&]() {  // make shared memory hold the accumulators
            if (SLABS && in_reg) {
#pragma unroll
                for (int c = 0; c < SC_SLAB_PAIRS; ++c) *reinterpret_cast<double2 *>(acc_w + 64 * c + 2 * lane) = r[c];
                in_reg = false;
                __syncwarp();
            } else if (first) {
                for (int i = lane * 2; i < sub; i += 64) *reinterpret_cast<double2 *>(acc_w + i) = make_double2(0.0, 0.0);
                __syncwarp();
            }
        };

        for (int j0 = qs; j0 < qe; j0 += 32) {
            const int nt = min(32, qe - j0);
            // lane j stages term j0 + j for this (tile, sub-tile); the offsets of the first 32 terms were prefetched
            uint32_t my_beg = 0, my_end = 0;
            int my_kind = 0;   // 0 sparse, 1 dense, 2 + h: slab of head row h
            typename TermW<KIND>::type my_idf = 0, my_qw = 0;
            if (lane < nt) {
                const int4 rc = load_rec(j0 + lane);
                my_idf = __int_as_float(rc.x);
                my_qw = __int_as_float(rc.y);
                if (j0 == qs) {
                    my_beg = pbeg;
                    my_end = pend;
                } else {
                    offsets_of(rc, my_beg, my_end);
                }
                if (rc.w >= 0) {
                    my_kind = 1;
                    // slabs add (idf * 0) * q for absent documents: only exact when both weights are finite
                    if (SLABS && rc.w < B2R_HEAD_TERMS && cache_ok[rc.w * B2R_SUBTILES + w] && isfinite((float)my_idf) &&
                        isfinite((float)my_qw))
                        my_kind = 2 + rc.w;
                }
            }
            for (int j = 0; j < nt; ++j) {
                const uint32_t beg = __shfl_sync(full, my_beg, j), end = __shfl_sync(full, my_end, j);
                if (beg == end) continue;  // warp-uniform
                const typename TermW<KIND>::type w_idf = __shfl_sync(full, my_idf, j), w_q = __shfl_sync(full, my_qw, j);
                const int kind = __shfl_sync(full, my_kind, j);
                if (diag && kind == 0) continue;   // B2R_SCORE_DIAG=1: measurement only (sparse terms skipped: wrong results)
                if (SLABS && kind >= 2) {
                    if (!in_reg) {
                        if (first) {
#pragma unroll
                            for (int c = 0; c < SC_SLAB_PAIRS; ++c) r[c] = make_double2(0.0, 0.0);
                        } else {  // (shared memory was written under __syncwarp)
#pragma unroll
                            for (int c = 0; c < SC_SLAB_PAIRS; ++c)
                                r[c] = *reinterpret_cast<const double2 *>(acc_w + 64 * c + 2 * lane);
                        }
                        in_reg = true;
                    }
                    const val2_t *cv = reinterpret_cast<const val2_t *>(cache + (size_t)(kind - 2) * tile_docs +
                                                                        (size_t)w * SC_SLAB_SUB) + lane;
                    if (KIND == B2R_KIND_BM25) {
                        const double wi = (double)w_idf, wq = (double)w_q;
#pragma unroll
                        for (int c = 0; c < SC_SLAB_PAIRS; ++c) {
                            const val2_t u = cv[32 * c];
                            r[c].x = __dadd_rn(r[c].x, __dmul_rn(__dmul_rn(wi, (double)u.x), wq));
                            r[c].y = __dadd_rn(r[c].y, __dmul_rn(__dmul_rn(wi, (double)u.y), wq));
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < SC_SLAB_PAIRS; ++c) {
                            const val2_t u = cv[32 * c];
                            // reference (fastmath) evaluates (tf * qtf) * idf in f32, then widens
                            r[c].x = __dadd_rn(r[c].x, (double)__fmul_rn(__fmul_rn((float)u.x, (float)w_q), (float)w_idf));
                            r[c].y = __dadd_rn(r[c].y, (double)__fmul_rn(__fmul_rn((float)u.y, (float)w_q), (float)w_idf));
                        }
                    }
                    first = false;
                    continue;
                }
                if (kind != 0) {
                    const bool was_first = first;
                    to_smem();
                    if (was_first) apply_dense_term<KIND, true>(beg, end, lane, my_doc0, post_doc, post_val, acc_w, w_idf, w_q);
                    else apply_dense_term<KIND, false>(beg, end, lane, my_doc0, post_doc, post_val, acc_w, w_idf, w_q);
                    first = false;
                    __syncwarp();
                } else {
                    // sparse term: the tile's block is shared by the 8 warps of every slot (kept in L1); a block with
                    // no posting in this warp's doc range leaves the accumulators where they are
                    const double w_idf64 = (double)w_idf, w_q64 = (double)w_q;
                    bool touched = false;
                    for (uint32_t p0 = beg; p0 < end; p0 += 32) {
                        const uint32_t p = p0 + lane;
                        const uint32_t rel = p < end ? __ldg(post_doc + p) - my_doc0 : 0xFFFFFFFFu;
                        const bool hit = rel < (uint32_t)sub;
                        if (!__any_sync(full, hit)) continue;
                        if (!touched) to_smem();
                        if (hit) {
                            const double u = load_val<KIND>(post_val, p);
                            apply_posting<KIND, false>(acc_w, rel, u, (float)w_idf, (float)w_q, w_idf64, w_q64);
                        }
                        touched = true;
                        first = false;
                    }
                    if (touched) __syncwarp();
                }
            }
        }

        // ---- epilogue over the sub-tile's accumulators: pair (i, i + 1), i = 64 c + 2 lane
        double m = -__longlong_as_double(0x7ff0000000000000ll);
        float *out_row = nullptr;
        if (OUT == SC_OUT_DENSE) out_row = o.scores + (int64_t)ql * o.scores_stride + (int64_t)y * tile_docs + w * sub;
        auto emit = [&](const double2 a, const int i) {
            if (OUT == SC_OUT_DENSE) {
                *reinterpret_cast<float2 *>(out_row + i) = make_float2(__double2float_rn(a.x), __double2float_rn(a.y));
            } else if (OUT == SC_OUT_MAXIMA) {
                // one maximum per lane over the (valid) documents it reads; f32(max) == max(f32): rounding is monotone
                const uint32_t doc = my_doc0 + i;
                if (doc < o.n_docs) m = fmax(m, a.x);
                if (doc + 1 < o.n_docs) m = fmax(m, a.y);
            } else {
                if (a.x >= thr_lo || a.y >= thr_lo) {
                    const double av[2] = {a.x, a.y};
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t doc = my_doc0 + i + c;
                        if (av[c] >= thr_lo && doc < o.n_docs) {
                            const uint64_t key = make_key(ord_f32(__double2float_rn(av[c])), o.doc_id_base + doc);
                            if (key > thr) {
                                const int slot_c = atomicAdd(o.cand_cnt + ql, 1);
                                if (slot_c < o.cap) o.cand[(int64_t)ql * o.cap + slot_c] = key;
                            }
                        }
                    }
                }
            }
        };
        if (SLABS && in_reg) {
#pragma unroll
            for (int c = 0; c < SC_SLAB_PAIRS; ++c) emit(r[c], 64 * c + 2 * lane);
        } else if (first) {   // no posting of the query in this sub-tile: every score is +0.0
            for (int i = lane * 2; i < sub; i += 64) emit(make_double2(0.0, 0.0), i);
        } else {
            for (int i = lane * 2; i < sub; i += 64) emit(*reinterpret_cast<const double2 *>(acc_w + i), i);
        }
        if (OUT == SC_OUT_MAXIMA)
            o.scores[(int64_t)ql * o.scores_stride + (int64_t)y * SC_GROUPS_PER_TILE + w * 32 + lane] = __double2float_rn(m);
        __syncwarp();   // the next query's writes must not overtake this epilogue's reads
        }  // (gate)
        qs = nqs;
        qe = nqe;
        nqs = nnqs;
        nqe = nnqe;
        pbeg = nbeg;
        pend = nend;
    }  // query loop
}

// ------------------------------------------------------------------------------------------------------------
// The scorer for tile_docs == B2R_SLAB_TILE_DOCS (the default layout): same tile-stationary scheme as above, written
// for instruction count -- the round-1 kernel and the generic kernel above spend most of their issue slots on per-
// (query, sub-tile) overhead (term staging, shuffles, threshold set-up), not on postings.  Here everything that does
// not depend on the warp is computed ONCE per CTA and chunk of queries and kept in shared memory:
//   rec[i]  one 16-byte record per query term {idf, query weight, first posting, kind word}, in term order
//           kind word: [7:0] 0 sparse / 1 dense / 2 + h head row h; [15:8] sub-tile mask (sparse: sub-tiles of this
//           tile that hold a posting of the term; head: sub-tiles whose segment has a slab); [31:16] sparse: postings
//           of the term in this tile
//   cum[i]  dense and head terms: the 9 sub-tile offsets of the term's postings in this tile, relative, u16
//   qthr[q] FUSED: the query's candidate threshold as a key and as the f64 bound the accumulators are compared with
// A (query, sub-tile) task is then a loop over records read by broadcast loads: no shuffles, no global loads before
// the postings themselves, all loop bounds constant.  A sparse term whose block has no posting in the warp's
// sub-tile costs one record load and a bit test.
constexpr int T2K_TILE = B2R_SLAB_TILE_DOCS;
constexpr int T2K_SUB = T2K_TILE / B2R_SUBTILES;        // 256 documents per warp
constexpr int T2K_PAIRS = T2K_SUB / 64;                 // double2 accumulators per lane
constexpr int T2K_SLOTS = 4;
constexpr int T2K_REC_CAP = 704;                        // term records per chunk of queries
constexpr int T2K_Q_CAP = 128;                          // queries per chunk
constexpr int T2K_CUM = 10;                             // u16 per record in cum[] (9 used)
static_assert(T2K_PAIRS == SC_SLAB_PAIRS && T2K_SUB == SC_SLAB_SUB, "slab layout");

struct __align__(16) T2KRec {
    float idf, w;
    uint32_t base, kind;
};
struct __align__(16) T2KThr {
    uint64_t thr;
    double lo;
};

template <int KIND>
constexpr size_t t2k_smem_bytes() {
    return (size_t)T2K_SLOTS * T2K_TILE * 8 + (size_t)B2R_HEAD_TERMS * T2K_TILE * (KIND == B2R_KIND_BM25 ? 8 : 4) +
           (size_t)T2K_REC_CAP * sizeof(T2KRec) + (size_t)T2K_REC_CAP * T2K_CUM * 2 + (size_t)(T2K_Q_CAP + 4) * 4 +
           (size_t)T2K_Q_CAP * sizeof(T2KThr);
}

// record of query term j for doc tile `tile` (see above); okm[h] = sub-tiles of this tile in which head row h has a slab
__device__ __forceinline__ T2KRec t2k_make_record(int j, int tile, int n_tiles, const int32_t *__restrict__ q_terms,
                                                  const float *__restrict__ q_weights, const float *__restrict__ idf,
                                                  const int32_t *__restrict__ dense_id, const uint32_t *__restrict__ dense_ptr,
                                                  const uint32_t *__restrict__ blk_ptr, const uint32_t *__restrict__ post_doc,
                                                  const int *okm, bool slabs, uint16_t *cum_out) {
    T2KRec rc;
    const int t = q_terms[j];
    rc.w = q_weights[j];
    rc.idf = idf[t];
    const int32_t did = dense_id[t];
    if (did >= 0) {
        const uint32_t *row = dense_ptr + (size_t)did * ((size_t)n_tiles * B2R_SUBTILES + 1) + (size_t)tile * B2R_SUBTILES;
        const uint32_t o0 = row[0];
        cum_out[0] = 0;
#pragma unroll
        for (int sg = 1; sg <= B2R_SUBTILES; ++sg) cum_out[sg] = (uint16_t)(row[sg] - o0);
        rc.base = o0;
        rc.kind = 1;
        // slabs add (idf * 0) * q for absent documents: only exact when both weights are finite
        if (slabs && did < B2R_HEAD_TERMS && isfinite(rc.idf) && isfinite(rc.w)) rc.kind = (2u + did) | ((uint32_t)okm[did] << 8);
    } else {
        const size_t e = (size_t)t * n_tiles + tile;
        const uint32_t beg = blk_ptr[e], n = blk_ptr[e + 1] - beg;
        uint32_t mask = n > 16 ? 0xFFu : 0u;
        if (n <= 16)
            for (uint32_t p = 0; p < n; ++p) mask |= 1u << (((post_doc[beg + p] - (uint32_t)tile * T2K_TILE) / T2K_SUB) & 7);
        rc.base = beg;
        rc.kind = (mask << 8) | (n << 16);
    }
    return rc;
}

template <int KIND, int OUT>
__global__ void __launch_bounds__(T2K_SLOTS * SC_THREADS, 1)
score_t2k_kernel(const uint32_t *__restrict__ post_doc, const void *__restrict__ post_val,
                 const uint32_t *__restrict__ blk_ptr, const int32_t *__restrict__ dense_id,
                 const uint32_t *__restrict__ dense_ptr, const int32_t *__restrict__ slab_idx,
                 const void *__restrict__ slab_val, int n_slabs, int n_tiles, const int32_t *__restrict__ q_ptr,
                 const int32_t *__restrict__ q_terms, const float *__restrict__ q_weights, const float *__restrict__ idf,
                 int q0, int nq, int range_len, int tile_mode, int tile_step, ScoreOut o) {
    using val_t = typename std::conditional<KIND == B2R_KIND_BM25, double, float>::type;
    using val2_t = typename std::conditional<KIND == B2R_KIND_BM25, double2, float2>::type;
    extern __shared__ __align__(128) unsigned char sc_smem[];
    __shared__ __align__(8) uint64_t cbar;   // "the slabs of this tile have landed"
    __shared__ int okm[B2R_HEAD_TERMS];      // per head row: sub-tiles of this tile whose segment has a slab
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = warp / B2R_SUBTILES, w = warp % B2R_SUBTILES;
    const int y = blockIdx.y;
    const int tile = tile_mode == SC_TILES_ALL ? y : y * tile_step;
    const size_t n_seg = (size_t)n_tiles * B2R_SUBTILES;
    const int r0 = blockIdx.x * range_len, r1 = min(nq, r0 + range_len);
    const bool slabs = slab_idx != nullptr && n_slabs > 0;

    double *acc_w = reinterpret_cast<double *>(sc_smem) + (size_t)slot * T2K_TILE + (size_t)w * T2K_SUB;
    unsigned char *sp = sc_smem + (size_t)T2K_SLOTS * T2K_TILE * 8;
    const val_t *cache = reinterpret_cast<const val_t *>(sp);
    sp += (size_t)B2R_HEAD_TERMS * T2K_TILE * sizeof(val_t);
    T2KRec *rec = reinterpret_cast<T2KRec *>(sp);
    sp += (size_t)T2K_REC_CAP * sizeof(T2KRec);
    uint16_t *cum = reinterpret_cast<uint16_t *>(sp);
    sp += (size_t)T2K_REC_CAP * T2K_CUM * 2;
    int *qoff = reinterpret_cast<int *>(sp);
    sp += (size_t)(T2K_Q_CAP + 4) * 4;
    T2KThr *qthr = reinterpret_cast<T2KThr *>(sp);

    if (OUT == SC_OUT_DENSE && o.gate != nullptr) {   // a gated launch is expected to find nothing to do
        int any = 0;
        for (int ql = r0 + (int)threadIdx.x; ql < r1; ql += blockDim.x) any |= o.gate[ql] > o.gate_cap;
        if (!__syncthreads_or(any)) return;
    }
    // ---- the tile's slabs -> shared memory (one bulk copy per (head row, sub-tile) that has one)
    const uint32_t cbar_a = sc_smem_u32(&cbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cbar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < B2R_HEAD_TERMS) okm[threadIdx.x] = 0;
    __syncthreads();
    if (warp == 0 && slabs) {
        uint32_t bytes = 0;
        for (int i = lane; i < B2R_HEAD_TERMS * B2R_SUBTILES; i += 32) {
            const int h = i / B2R_SUBTILES, sseg = i % B2R_SUBTILES;
            const int32_t sid = slab_idx[(size_t)h * n_seg + (size_t)tile * B2R_SUBTILES + sseg];
            if (sid >= 0 && sid < n_slabs) {
                atomicOr(&okm[h], 1 << sseg);
                const uint32_t nb = (uint32_t)(T2K_SUB * sizeof(val_t));
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        sc_smem_u32(cache + (size_t)h * T2K_TILE + (size_t)sseg * T2K_SUB)),
                    "l"(static_cast<const val_t *>(slab_val) + (size_t)sid * T2K_SUB), "r"(nb), "r"(cbar_a)
                    : "memory");
                bytes += nb;
            }
        }
#pragma unroll
        for (int ofs = 16; ofs; ofs >>= 1) bytes += __shfl_xor_sync(full, bytes, ofs);
        // (a copy that completes before this arrive only drives the transaction count negative for a moment:
        // the phase cannot complete before the one pending arrival below)
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cbar_a), "r"(bytes) : "memory");
    } else if (threadIdx.x == 0 && !slabs) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cbar_a) : "memory");
    }
    __syncthreads();   // okm is complete

    const uint32_t my_doc0 = (uint32_t)tile * T2K_TILE + (uint32_t)w * T2K_SUB;
    bool cache_ready = false;
    for (int cur = r0; cur < r1;) {
        // ---- chunk of queries [cur, cur + c): as many as fit the record and query capacities (at least one)
        const int cnt = min(r1 - cur, T2K_Q_CAP);
        const int tb = q_ptr[q0 + cur];
        int fits = 0;
        if ((int)threadIdx.x < cnt) fits = q_ptr[q0 + cur + threadIdx.x + 1] - tb <= T2K_REC_CAP;
        int c = __syncthreads_count(fits);   // (q_ptr is monotone: the queries that fit form a prefix)
        if (c < 1) c = 1;
        const int n_staged = min(q_ptr[q0 + cur + c] - tb, T2K_REC_CAP);
        for (int i = threadIdx.x; i < n_staged; i += blockDim.x)
            rec[i] = t2k_make_record(tb + i, tile, n_tiles, q_terms, q_weights, idf, dense_id, dense_ptr, blk_ptr, post_doc,
                                     okm, slabs, cum + (size_t)i * T2K_CUM);
        for (int i = threadIdx.x; i <= c; i += blockDim.x) qoff[i] = q_ptr[q0 + cur + i] - tb;
        if (OUT == SC_OUT_FUSED) {
            for (int i = threadIdx.x; i < c; i += blockDim.x) {
                const uint64_t thr = o.thr_keys[cur + i];
                const uint32_t thr_hi = (uint32_t)(thr >> 32);
                // A document can only beat thr if f32(acc) >= the threshold score.  The accumulators are compared in
                // f64 against the f32 value just below the threshold score: acc < pred(thr_f) implies f32(acc) <=
                // pred(thr_f) < thr_f (rounding is monotone); only survivors are converted and keyed.  Ordered
                // encoding: -1 = next smaller f32; 0x7fffffff would be -0.0 (== +0.0 in the ranking): skip it; at or
                // below -inf (0x007fffff) there is nothing smaller: no filter (thr_hi == 0: no threshold at all)
                uint32_t thr_ord = thr_hi > 0x007fffffu ? thr_hi - 1u : 0u;
                if (thr_ord == 0x7fffffffu) thr_ord = 0x7ffffffeu;
                double lo = thr_ord ? (double)unord_f32(thr_ord) : -__longlong_as_double(0x7ff0000000000000ll);
                // "strictly positive scores only" (kth_of_maxima's positive floor): nothing below 2^-150 rounds to a
                // positive f32, so the untouched documents (acc == 0) never reach the conversion path
                if (thr == ((0x80000000ull << 32) | 0xFFFFFFFFull)) lo = __longlong_as_double(0x3690000000000000ll);
                qthr[i].thr = thr;
                qthr[i].lo = lo;
            }
        }
        __syncthreads();
        if (!cache_ready) {
            sc_mbar_wait(cbar_a, 0);   // the slab bytes are visible
            cache_ready = true;
        }

        for (int qi = slot; qi < c; qi += T2K_SLOTS) {
            const int ql = cur + qi;
            if (OUT == SC_OUT_DENSE && o.gate != nullptr && o.gate[ql] <= o.gate_cap) continue;
            const int i0 = qoff[qi], i1 = qoff[qi + 1];
            // where the sub-tile's accumulators are (warp-uniform): 0 nowhere yet (all +0.0), 1 registers, 2 shared memory
            int st = 0;
            double2 r[T2K_PAIRS];
#pragma unroll
            for (int cc = 0; cc < T2K_PAIRS; ++cc) r[cc] = make_double2(0.0, 0.0);
            auto to_smem = [&]() {
                if (st == 2) return;
#pragma unroll
                for (int cc = 0; cc < T2K_PAIRS; ++cc) *reinterpret_cast<double2 *>(acc_w + 64 * cc + 2 * lane) = r[cc];
                st = 2;     // (r[] holds zeros in state 0)
                __syncwarp();
            };
            for (int i = i0; i < i1; ++i) {
                T2KRec rc;
                uint16_t cum_l[T2K_CUM];
                const bool staged = i < n_staged;
                if (staged) rc = rec[i];
                else rc = t2k_make_record(tb + i, tile, n_tiles, q_terms, q_weights, idf, dense_id, dense_ptr, blk_ptr, post_doc,
                                          okm, slabs, cum_l);     // (a query longer than the record capacity)
                const uint32_t kind = rc.kind & 0xFFu;
                if (kind >= 2u && ((rc.kind >> (8 + w)) & 1u)) {
                    // ---- slab: register accumulators
                    if (st == 2) {
#pragma unroll
                        for (int cc = 0; cc < T2K_PAIRS; ++cc)
                            r[cc] = *reinterpret_cast<const double2 *>(acc_w + 64 * cc + 2 * lane);
                    }
                    st = 1;
                    const val2_t *cv = reinterpret_cast<const val2_t *>(cache + (size_t)(kind - 2u) * T2K_TILE +
                                                                        (size_t)w * T2K_SUB) + lane;
                    if (KIND == B2R_KIND_BM25) {
                        const double wi = (double)rc.idf, wq = (double)rc.w;
#pragma unroll
                        for (int cc = 0; cc < T2K_PAIRS; ++cc) {
                            const val2_t u = cv[32 * cc];
                            r[cc].x = __dadd_rn(r[cc].x, __dmul_rn(__dmul_rn(wi, (double)u.x), wq));
                            r[cc].y = __dadd_rn(r[cc].y, __dmul_rn(__dmul_rn(wi, (double)u.y), wq));
                        }
                    } else {
#pragma unroll
                        for (int cc = 0; cc < T2K_PAIRS; ++cc) {
                            const val2_t u = cv[32 * cc];
                            // reference (fastmath) evaluates (tf * qtf) * idf in f32, then widens
                            r[cc].x = __dadd_rn(r[cc].x, (double)__fmul_rn(__fmul_rn((float)u.x, rc.w), rc.idf));
                            r[cc].y = __dadd_rn(r[cc].y, (double)__fmul_rn(__fmul_rn((float)u.y, rc.w), rc.idf));
                        }
                    }
                } else if (kind >= 1u) {
                    // ---- dense term (or a head term whose segment has no slab here): my bank-scheduled segment
                    uint32_t cb, ce;
                    if (staged) {
                        cb = cum[(size_t)i * T2K_CUM + w];
                        ce = cum[(size_t)i * T2K_CUM + w + 1];
                    } else {
                        cb = cum_l[w];
                        ce = cum_l[w + 1];
                    }
                    if (cb == ce) continue;
                    to_smem();
                    apply_dense_term<KIND, false>(rc.base + cb, rc.base + ce, lane, my_doc0, post_doc, post_val, acc_w, rc.idf, rc.w);
                    __syncwarp();
                } else {
                    // ---- sparse term: the tile's block, shared by all warps (kept in L1); the mask says whether it has
                    // a posting in my sub-tile (blocks of more than 16 postings are always scanned)
                    if (!((rc.kind >> (8 + w)) & 1u)) continue;
                    const uint32_t n = rc.kind >> 16;
                    to_smem();
                    const double w_idf64 = (double)rc.idf, w_q64 = (double)rc.w;
                    for (uint32_t p0 = 0; p0 < n; p0 += 32) {
                        const uint32_t p = p0 + lane;
                        if (p < n) {
                            const uint32_t rel = __ldg(post_doc + rc.base + p) - my_doc0;
                            if (rel < (uint32_t)T2K_SUB) {
                                const double u = load_val<KIND>(post_val, rc.base + p);
                                apply_posting<KIND, false>(acc_w, rel, u, rc.idf, rc.w, w_idf64, w_q64);
                            }
                        }
                    }
                    __syncwarp();
                }
            }

            // ---- epilogue over the sub-tile's accumulators: pair (i, i + 1), i = 64 cc + 2 lane
            if (st == 2) {
#pragma unroll
                for (int cc = 0; cc < T2K_PAIRS; ++cc) r[cc] = *reinterpret_cast<const double2 *>(acc_w + 64 * cc + 2 * lane);
            }
            if (OUT == SC_OUT_DENSE) {
                float *out_row = o.scores + (int64_t)ql * o.scores_stride + (int64_t)y * T2K_TILE + w * T2K_SUB + 2 * lane;
#pragma unroll
                for (int cc = 0; cc < T2K_PAIRS; ++cc)
                    *reinterpret_cast<float2 *>(out_row + 64 * cc) =
                        make_float2(__double2float_rn(r[cc].x), __double2float_rn(r[cc].y));
            } else if (OUT == SC_OUT_MAXIMA) {
                // one maximum per lane over the (valid) documents it owns; f32(max) == max(f32): rounding is monotone
                double m = -__longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
                for (int cc = 0; cc < T2K_PAIRS; ++cc) {
                    const uint32_t doc = my_doc0 + 64 * cc + 2 * lane;
                    if (doc < o.n_docs) m = fmax(m, r[cc].x);
                    if (doc + 1 < o.n_docs) m = fmax(m, r[cc].y);
                }
                o.scores[(int64_t)ql * o.scores_stride + (int64_t)y * SC_GROUPS_PER_TILE + w * 32 + lane] = __double2float_rn(m);
            } else {
                const T2KThr th = qthr[qi];
                bool any = false;
#pragma unroll
                for (int cc = 0; cc < T2K_PAIRS; ++cc) any |= (r[cc].x >= th.lo) | (r[cc].y >= th.lo);
                if (any) {
#pragma unroll
                    for (int cc = 0; cc < T2K_PAIRS; ++cc) {
                        const double av[2] = {r[cc].x, r[cc].y};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const uint32_t doc = my_doc0 + 64 * cc + 2 * lane + e;
                            if (av[e] >= th.lo && doc < o.n_docs) {
                                const uint64_t key = make_key(ord_f32(__double2float_rn(av[e])), o.doc_id_base + doc);
                                if (key > th.thr) {
                                    const int slot_c = atomicAdd(o.cand_cnt + ql, 1);
                                    if (slot_c < o.cap) o.cand[(int64_t)ql * o.cap + slot_c] = key;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();   // the next query's writes must not overtake this epilogue's reads
        }  // queries of the chunk
        __syncthreads();   // the records are about to be replaced
        cur += c;
    }
}

constexpr size_t SC_SMEM_MAX = 226 * 1024;   // dynamic shared memory of a scorer CTA (227 KB minus the static part)

struct ScoreLaunch {
    const b2r_index *ix;
    const int32_t *q_ptr, *q_terms;
    const float *q_weights, *idf;
    cudaStream_t st;
};

static bool g_slabs_enabled = true;   // b2r_set_slabs: test / profiling hook (same results either way)

// CTAs per doc tile (= query ranges): enough CTAs for several waves over the 148 SMs, but every CTA should walk
// enough queries to pay for its slab copy (B2R_SCORE_RANGES overrides: tuning experiments only)
static int ranges_override() {
    const char *e = getenv("B2R_SCORE_RANGES");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 65535 ? v : 0;
}
static const int g_ranges_override = ranges_override();
static const int g_generic_scorer = [] {   // B2R_GENERIC_SCORER=1: tile 2048 through the generic kernel (A/B only)
    const char *e = getenv("B2R_GENERIC_SCORER");
    return e ? atoi(e) : 0;
}();
static const int g_score_diag = [] {
    const char *e = getenv("B2R_SCORE_DIAG");
    return e ? atoi(e) : 0;
}();

template <int KIND, int OUT, bool SLABS>
static int launch_score_impl(const ScoreLaunch &L, int q0, int nq, int tile_mode, int tile_step, int n_y, const ScoreOut &o) {
    const b2r_index *ix = L.ix;
    const size_t acc_bytes = (size_t)ix->tile_docs * sizeof(double);
    const size_t cache_bytes = SLABS ? (size_t)B2R_HEAD_TERMS * ix->tile_docs * (KIND == B2R_KIND_BM25 ? 8 : 4) : 0;
    const size_t rec_bytes = (size_t)SC_REC_CAP * sizeof(int4);
    int slots = (int)((SC_SMEM_MAX - rec_bytes - cache_bytes) / acc_bytes);
    slots = slots > SC_MAX_SLOTS ? SC_MAX_SLOTS : (slots < 1 ? 1 : slots);
    if (slots > nq) slots = nq;
    const size_t smem = (size_t)slots * acc_bytes + cache_bytes + rec_bytes;
    int ranges = g_ranges_override ? g_ranges_override : (148 * 8 + n_y - 1) / n_y;
    const int max_ranges = (nq + 4 * slots - 1) / (4 * slots);   // >= 4 queries per slot and CTA
    if (ranges > max_ranges) ranges = max_ranges;
    if (ranges < 1) ranges = 1;
    const int range_len = (nq + ranges - 1) / ranges;
    ranges = (nq + range_len - 1) / range_len;
    auto kern = score_tiles_kernel<KIND, OUT, SLABS>;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM_MAX));
        attr_set = true;
    }
    dim3 grid((unsigned)ranges, (unsigned)n_y);
    kern<<<grid, slots * SC_THREADS, smem, L.st>>>(ix->post_doc, ix->post_val, ix->blk_ptr, ix->dense_id, ix->dense_ptr,
                                                   ix->slab_idx, ix->slab_val, ix->n_slabs, ix->n_tiles, ix->tile_docs,
                                                   L.q_ptr, L.q_terms, L.q_weights, L.idf, q0, nq, range_len, tile_mode,
                                                   tile_step, OUT == SC_OUT_FUSED ? g_score_diag : 0, o);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

template <int KIND, int OUT>
static int launch_score_t2k(const ScoreLaunch &L, int q0, int nq, int tile_mode, int tile_step, int n_y, const ScoreOut &o) {
    const b2r_index *ix = L.ix;
    // CTAs per doc tile (= query ranges): several waves over the 148 SMs, but at least ~32 queries per CTA so that the
    // slab copy and the record staging are paid for
    int ranges = g_ranges_override ? g_ranges_override : (148 * 8 + n_y - 1) / n_y;
    const int max_ranges = (nq + 31) / 32;
    if (ranges > max_ranges) ranges = max_ranges;
    if (ranges < 1) ranges = 1;
    const int range_len = (nq + ranges - 1) / ranges;
    ranges = (nq + range_len - 1) / range_len;
    auto kern = score_t2k_kernel<KIND, OUT>;
    constexpr size_t smem = t2k_smem_bytes<KIND>();
    static_assert(smem <= SC_SMEM_MAX, "t2k shared memory");
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const bool slabs = g_slabs_enabled && ix->slab_idx && ix->slab_val && ix->n_slabs > 0;
    dim3 grid((unsigned)ranges, (unsigned)n_y);
    kern<<<grid, T2K_SLOTS * SC_THREADS, smem, L.st>>>(ix->post_doc, ix->post_val, ix->blk_ptr, ix->dense_id, ix->dense_ptr,
                                                       slabs ? ix->slab_idx : nullptr, ix->slab_val, slabs ? ix->n_slabs : 0,
                                                       ix->n_tiles, L.q_ptr, L.q_terms, L.q_weights, L.idf, q0, nq, range_len,
                                                       tile_mode, tile_step, o);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

template <int OUT>
static int launch_score(const ScoreLaunch &L, int q0, int nq, int tile_mode, int tile_step, int n_y, const ScoreOut &o) {
    if (nq == 0 || n_y == 0) return B2R_OK;
    const b2r_index *ix = L.ix;
    if (ix->tile_docs == T2K_TILE && !g_generic_scorer) {
        if (ix->kind == B2R_KIND_BM25) return launch_score_t2k<B2R_KIND_BM25, OUT>(L, q0, nq, tile_mode, tile_step, n_y, o);
        return launch_score_t2k<B2R_KIND_IMPACT, OUT>(L, q0, nq, tile_mode, tile_step, n_y, o);
    }
    if (ix->kind == B2R_KIND_BM25) return launch_score_impl<B2R_KIND_BM25, OUT, false>(L, q0, nq, tile_mode, tile_step, n_y, o);
    return launch_score_impl<B2R_KIND_IMPACT, OUT, false>(L, q0, nq, tile_mode, tile_step, n_y, o);
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

// ---- fused-selection plan ------------------------------------------------------------------------
constexpr int FUSED_MIN_TILES = 8;    // below this the plain score + select path is used
constexpr int FUSED_MAX_K = 128;

static bool g_fused_enabled = true;
// optional CUDA-event bracket around the fused scoring launch (bench.py's roofline of the dominant kernel)
static bool g_profile = false;
static cudaEvent_t g_ev[2] = {nullptr, nullptr};

struct FusedPlan {
    bool on;
    int step;  // every step-th doc tile forms the threshold sample
    int n_sample, cap;
    int64_t n_groups;  // group maxima per query = n_sample * SC_GROUPS_PER_TILE
};

static FusedPlan fused_plan(const b2r_index *ix, int k, bool want_scores) {
    FusedPlan p = {};
    p.on = g_fused_enabled && !want_scores && k >= 1 && k <= FUSED_MAX_K && ix->n_tiles >= FUSED_MIN_TILES;
    if (!p.on) return p;
    // The threshold is (about) the k-th best of a 1/r sample (r = n_tiles / n_sample <= step), so a query collects
    // ~ k * r candidates (negative-binomial: sigma ~ sqrt(k) * r); the cap is the power of two above mean + 6 sigma.
    p.step = k <= 16 ? 64 : 16;
    p.n_sample = (ix->n_tiles + p.step - 1) / p.step;
    {
        const double r = (double)ix->n_tiles / p.n_sample;
        const double want = k * r + 6.0 * sqrt((double)k) * r + k;
        p.cap = 256;
        while (p.cap < want && p.cap < 4096) p.cap <<= 1;
    }
    p.n_groups = (int64_t)p.n_sample * SC_GROUPS_PER_TILE;
    // tile 0 is full (n_tiles >= 8) and holds min(tile_docs / 2, 256) non-empty groups
    const int groups_tile0 = ix->tile_docs / 2 < SC_GROUPS_PER_TILE ? ix->tile_docs / 2 : SC_GROUPS_PER_TILE;
    if (groups_tile0 < k) p.on = false;
    return p;
}

// workspace bytes needed to run `qc` queries in one pass
static size_t pass_bytes(const b2r_index *ix, const FusedPlan &fp, int64_t qc, int k) {
    const size_t full = align_up((size_t)padded_docs(ix) * 4 * (size_t)qc, 256) + topk_ws_bytes(qc, ix->n_docs, k);
    if (!fp.on) return full;
    return full + align_up((size_t)fp.n_groups * 4 * (size_t)qc, 256) + align_up((size_t)qc * 8, 256) +
           align_up((size_t)qc * fp.cap * 8, 256) + align_up((size_t)qc * 4, 256) + 256;
}

}  // namespace b2r

using namespace b2r;

// test / profiling hook: 0 disables the fused-selection path (plain score + select is used)
extern "C" void b2r_set_fused_selection(int enabled) { b2r::g_fused_enabled = enabled != 0; }
// test / profiling hook: 0 makes the scorer ignore an index's slabs (shared-memory accumulators only)
extern "C" void b2r_set_slabs(int enabled) { b2r::g_slabs_enabled = enabled != 0; }

extern "C" int b2r_set_profiling(int enabled) {
    if (enabled && !g_ev[0]) {
        B2R_CUDA(cudaEventCreate(&g_ev[0]));
        B2R_CUDA(cudaEventCreate(&g_ev[1]));
    }
    g_profile = enabled != 0;
    return B2R_OK;
}

extern "C" int b2r_profile_fused_ms(float *ms, int32_t *n_tiles_scored) {
    B2R_CHECK_ARG(ms && g_ev[0], "b2r_profile_fused_ms: profiling was never enabled");
    B2R_CUDA(cudaEventSynchronize(g_ev[1]));
    B2R_CUDA(cudaEventElapsedTime(ms, g_ev[0], g_ev[1]));
    (void)n_tiles_scored;
    return B2R_OK;
}

extern "C" int b2r_fused_plan(const b2r_index *ix, int32_t k, int32_t *n_sample_tiles, int32_t *tile_step,
                              int32_t *cap) {
    int rc = check_index(ix);
    if (rc) return rc;
    FusedPlan fp = fused_plan(ix, k, false);
    if (n_sample_tiles) *n_sample_tiles = fp.on ? fp.n_sample : 0;
    if (tile_step) *tile_step = fp.on ? fp.step : 0;
    if (cap) *cap = fp.on ? fp.cap : 0;
    return B2R_OK;
}

extern "C" int b2r_search_workspace(const b2r_index *ix, int32_t n_queries, int32_t k, size_t *min_bytes,
                                    size_t *full_bytes) {
    int rc = check_index(ix);
    if (rc) return rc;
    B2R_CHECK_ARG(n_queries >= 0 && k >= 0 && k <= B2R_TOPK_MAX_FAST, "b2r_search_workspace: bad n_queries/k");
    const int kk = k > 0 ? k : 1;
    const int64_t nq = n_queries > 0 ? n_queries : 1;
    const FusedPlan fp = fused_plan(ix, kk, false);
    const size_t keys = align_up((size_t)nq * kk * 8, 256);
    if (min_bytes) *min_bytes = pass_bytes(ix, fp, 1, kk) + keys + 512;
    if (full_bytes) *full_bytes = pass_bytes(ix, fp, nq, kk) + keys + 512;
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
    ScoreLaunch L = {ix, q_ptr, q_terms, q_weights, idf, st};

    if (scores_out) {
        ScoreOut o = {};
        o.scores = scores_out;
        o.scores_stride = scores_stride;
        rc = launch_score<SC_OUT_DENSE>(L, 0, n_queries, SC_TILES_ALL, 1, ix->n_tiles, o);
        if (rc) return rc;
        if (k > 0) {
            rc = topk_scores_rows(scores_out, n_queries, ix->n_docs, scores_stride, k, ix->doc_id_base, keys, wp, left,
                                  st);
            if (rc) return rc;
        }
        if (k > 0) return decode_keys(keys, (int64_t)n_queries * k, idx_out, val_out, nullptr, 0, k, 0, st);
        return B2R_OK;
    }

    // queries are processed in chunks sized by the workspace
    const FusedPlan fp = fused_plan(ix, k, false);
    int64_t qc = n_queries;
    while (qc > 1 && pass_bytes(ix, fp, qc, k) > left) qc = (qc + 1) / 2;
    if (pass_bytes(ix, fp, qc, k) > left) {
        set_error("b2r_search_batch: workspace too small (%zu bytes left, one query needs %zu)", left,
                  pass_bytes(ix, fp, 1, k));
        return B2R_ERR_WORKSPACE;
    }
    float *full = static_cast<float *>(carve((size_t)pad * 4 * (size_t)qc));
    const size_t tk_full_bytes = topk_ws_bytes(qc, ix->n_docs, k);
    void *tk_full = carve(tk_full_bytes);
    float *maxima = nullptr;
    uint64_t *thr = nullptr, *cand = nullptr;
    int32_t *cand_cnt = nullptr;
    if (fp.on) {
        maxima = static_cast<float *>(carve((size_t)fp.n_groups * 4 * (size_t)qc));
        thr = static_cast<uint64_t *>(carve((size_t)qc * 8));
        cand = static_cast<uint64_t *>(carve((size_t)qc * fp.cap * 8));
        cand_cnt = static_cast<int32_t *>(carve((size_t)qc * 4));
    }

    for (int64_t q0 = 0; q0 < n_queries; q0 += qc) {
        const int nq = (int)((n_queries - q0) < qc ? (n_queries - q0) : qc);
        uint64_t *kout = keys + q0 * k;
        TopkOpts gate;
        if (fp.on) {
            // 1. threshold: group maxima of every step-th tile, then the k-th largest of them per query
            ScoreOut so = {};
            so.scores = maxima;
            so.scores_stride = fp.n_groups;
            so.n_docs = (uint32_t)ix->n_docs;
            rc = launch_score<SC_OUT_MAXIMA>(L, (int)q0, nq, SC_TILES_SAMPLE, fp.step, fp.n_sample, so);
            if (rc) return rc;
            rc = kth_of_maxima(maxima, nq, fp.n_groups, fp.n_groups, k, false, true, thr, st);
            if (rc) return rc;
            // 2. every tile: score, keep only the documents that reach the threshold
            B2R_CUDA(cudaMemsetAsync(cand_cnt, 0, (size_t)nq * 4, st));
            ScoreOut fo = {};
            fo.thr_keys = thr;
            fo.cand = cand;
            fo.cand_cnt = cand_cnt;
            fo.cap = fp.cap;
            fo.n_docs = (uint32_t)ix->n_docs;
            fo.doc_id_base = (uint32_t)ix->doc_id_base;
            if (g_profile) B2R_CUDA(cudaEventRecord(g_ev[0], st));
            rc = launch_score<SC_OUT_FUSED>(L, (int)q0, nq, SC_TILES_ALL, 1, ix->n_tiles, fo);
            if (rc) return rc;
            if (g_profile) B2R_CUDA(cudaEventRecord(g_ev[1], st));
            // 3. exact top-k of the candidates
            rc = topk_of_lists(cand, nq, fp.cap, cand_cnt, k, k, kout, st);
            if (rc) return rc;
            // 4. exact fallback, gated on the device to the queries whose list overflowed
            gate.gate = cand_cnt;
            gate.gate_cap = fp.cap;
        }
        ScoreOut o = {};
        o.scores = full;
        o.scores_stride = pad;
        o.gate = gate.gate;
        o.gate_cap = gate.gate_cap;
        rc = launch_score<SC_OUT_DENSE>(L, (int)q0, nq, SC_TILES_ALL, 1, ix->n_tiles, o);
        if (rc) return rc;
        rc = topk_scores_rows(full, nq, ix->n_docs, pad, k, ix->doc_id_base, kout, tk_full, tk_full_bytes, st, gate);
        if (rc) return rc;
    }
    return decode_keys(keys, (int64_t)n_queries * k, idx_out, val_out, nullptr, 0, k, 0, st);
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
