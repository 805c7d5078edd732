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
// Four epilogues.  DENSE writes the f32 score vector (the reference's return value).  MAXIMA, FUSED and EXHAUSTIVE
// are the search path, which never materialises scores: (1) every 64th / 16th doc tile is scored with the
// MAXIMA epilogue, which emits one maximum per lane (a group of <= tile_docs/256 documents); the k-th largest
// of those group maxima is a lower bound T of the k-th best score of the whole shard (k groups, hence k
// documents, reach it).  (2) ALL tiles are scored with the FUSED epilogue, which appends the few documents
// whose key beats T to a per-query candidate list; the exact top-k is the top-k of that list.  A query whose
// list overflows (or comes up short) is rescored by the EXHAUSTIVE launch that follows: its CTAs leave at once for
// every other query, and for a marked query they score every tile again and keep a streaming top-k in shared
// memory (no score vector), so the result is exact for any corpus order with a workspace that does not grow with
// the corpus; the gate is evaluated on the device, nothing synchronises.
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

constexpr int SC_THREADS = 32 * B2R_SUBTILES;  // one warp per sub-tile of the CTA's doc tile

// The accumulators are cleared by the copy engine (cp.async.bulk from this zero page), not by stores:
// the LSU / shared-memory data pipe is the scorer's bottleneck and the bulk copy does not go through it.
__device__ __align__(128) double g_zero_page[16384 / B2R_SUBTILES];  // one sub-tile of the largest tile

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

// One term applied by one warp to its own sub-tile.  dense: [beg, end) are exactly the warp's postings;
// otherwise [beg, end) is the tile's (small) block and the warp keeps the postings of its doc range.
template <int KIND, bool FIRST>
__device__ __forceinline__ void apply_term(int dense, uint32_t beg, uint32_t end, int lane, int sub, uint32_t my_doc0,
                                           const uint32_t *__restrict__ post_doc, const void *__restrict__ post_val,
                                           double *acc_w, typename TermW<KIND>::type w_idf_t,
                                           typename TermW<KIND>::type w_q_t) {
    // BM25 multiplies in f64: the widening was done ONCE per lane when the query was staged (f32 -> f64 conversions
    // run on a slow pipe, and this function is entered once per term, warp and tile)
    const double w_idf64 = (double)w_idf_t, w_q64 = (double)w_q_t;
    const float w_idf = (float)w_idf_t, w_q = (float)w_q_t;  // (exact: the values are f32 numbers)
    if (dense) {
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
    } else {
        for (uint32_t p = beg + lane; p < end; p += 32) {
            const uint32_t rel = __ldg(post_doc + p) - my_doc0;  // blocks are shared by the 8 warps: keep in L1
            if (rel < (uint32_t)sub) {
                double u = load_val<KIND>(post_val, p);
                apply_posting<KIND, FIRST>(acc_w, rel, u, w_idf, w_q, w_idf64, w_q64);
            }
        }
    }
}

// One CTA = one (query, doc tile); one WARP = one sub-tile of tile_docs/8 docs whose f64 accumulators
// it alone touches.  A warp applies the query's terms in ascending term id to its own sub-tile, so the
// only synchronisation is __syncwarp: no CTA barrier, no atomics, no load imbalance between warps
// (a dense term's postings are split by sub-tile through dense_ptr; a sparse term's small block is
// scanned by every warp, each keeping the postings that fall in its range).
enum { SC_OUT_DENSE = 0, SC_OUT_FUSED = 1, SC_OUT_MAXIMA = 2, SC_OUT_EXHAUSTIVE = 3 };
constexpr int SC_XK_CAP = 8192;    // EXHAUSTIVE: keys a CTA keeps in shared memory (k best so far + pending candidates)
constexpr int SC_XK_CHUNK = 4096;  // EXHAUSTIVE: documents examined between two capacity checks
constexpr int SC_X_PARTS = 32;     // EXHAUSTIVE: most CTAs per query, each with its own share of the doc tiles
enum { SC_TILES_ALL = 0, SC_TILES_SAMPLE = 1 };
constexpr int SC_GROUPS_PER_TILE = 32 * B2R_SUBTILES;  // MAXIMA: one group maximum per lane

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
    // EXHAUSTIVE: exact top-k with no score vector for the queries the selection kernels marked (overflowed or short
    // candidate list): marked[0] = their number, marked[1 + i] = their indices inside the chunk.  A fixed, small grid
    // walks the work items (marked query, part): an unmarked batch costs one wave of CTAs that read one integer.
    const int32_t *marked;
    int32_t x_n_parts;      // parts a marked query's doc tiles are split into (<= SC_X_PARTS)
    int32_t k;
    uint64_t *x_parts;      // [queries, SC_X_PARTS, k] the k best of every part's tiles
    int32_t *x_done;        // [queries] parts finished (zeroed by the caller); the last one merges
    uint64_t *keys_final;   // [queries, k] (any of the three may be null)
    int64_t *idx_final;
    float *val_final;
};

template <int KIND, int OUT>
#ifndef SC_MIN_CTAS
#define SC_MIN_CTAS 6  // 6 CTAs/SM is what 32 KB of accumulators per CTA allows (40 registers per thread)
#endif
__global__ void __launch_bounds__(SC_THREADS, SC_MIN_CTAS)
score_tiles_kernel(const uint32_t *__restrict__ post_doc, const void *__restrict__ post_val,
                   const uint32_t *__restrict__ blk_ptr, const int32_t *__restrict__ dense_id,
                   const uint32_t *__restrict__ dense_ptr, int n_tiles, int tile_docs,
                   const int32_t *__restrict__ q_ptr, const int32_t *__restrict__ q_terms,
                   const float *__restrict__ q_weights, const float *__restrict__ idf, int q0, int tile_mode,
                   int tile_step, int n_y, ScoreOut o) {
    extern __shared__ double acc[];  // [tile_docs]
    __shared__ __align__(8) uint64_t zbar[B2R_SUBTILES];  // per warp: "my accumulators have been cleared"
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (OUT == SC_OUT_DENSE && o.gate != nullptr && o.gate[blockIdx.x] <= o.gate_cap) return;
    // work items: every other epilogue has exactly one per CTA, (query blockIdx.x, tiles blockIdx.y + i gridDim.y);
    // EXHAUSTIVE loops over (marked query, part) pairs
    const int n_items = OUT == SC_OUT_EXHAUSTIVE ? o.marked[0] * o.x_n_parts : 1;
    const int y_stride = OUT == SC_OUT_EXHAUSTIVE ? o.x_n_parts : (int)gridDim.y;
    // EXHAUSTIVE: the CTA's k best keys so far live sorted in xk[0, x_kept), pending candidates behind them
    uint64_t *xk = reinterpret_cast<uint64_t *>(acc + tile_docs);
    __shared__ int x_cnt, x_last;
    __shared__ uint64_t x_thr;
    auto x_compact = [&]() {   // all threads: sort, keep the k best, raise the threshold
        const int n = x_cnt;
        int P = 32;
        while (P < n) P <<= 1;
        for (int i = n + (int)threadIdx.x; i < P; i += SC_THREADS) xk[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc<SC_THREADS>(xk, P);
        if (threadIdx.x == 0) {
            x_cnt = n < o.k ? n : o.k;
            x_thr = n >= o.k ? xk[o.k - 1] : 0ull;
        }
        __syncthreads();
    };
    const int sub = tile_docs / B2R_SUBTILES;
    double *acc_w = acc + w * sub;
    const size_t dense_row = (size_t)n_tiles * B2R_SUBTILES + 1;
    const uint32_t zbar_a = sc_smem_u32(&zbar[w]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(zbar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t zphase = 0;
  for (int item = OUT == SC_OUT_EXHAUSTIVE ? (int)blockIdx.x : 0; item < n_items;
       item += OUT == SC_OUT_EXHAUSTIVE ? (int)gridDim.x : 1) {
    const int ql = OUT == SC_OUT_EXHAUSTIVE ? o.marked[1 + item / o.x_n_parts] : (int)blockIdx.x;  // query inside the chunk
    const int y_first = OUT == SC_OUT_EXHAUSTIVE ? item % o.x_n_parts : (int)blockIdx.y;
    const int q = q0 + ql;
    if (OUT == SC_OUT_EXHAUSTIVE) {
        __syncthreads();   // (the previous item's shared state is no longer read)
        if (threadIdx.x == 0) {
            x_cnt = 0;
            x_thr = 0ull;
        }
        __syncthreads();
    }
    const int qs = q_ptr[q], qe = q_ptr[q + 1];
    auto tile_of = [&](int y) -> int {
        return tile_mode == SC_TILES_ALL ? y : y * tile_step;
    };
    // A CTA walks several doc tiles (stride gridDim.y).  A query of <= 32 terms is staged ONCE, lane j keeping
    // term j's weights and the base of its offset row; per tile only the two offsets of the warp's posting range
    // are fetched, one tile ahead, so the dependent chain q_terms -> dense_id -> offsets -> postings is paid once
    // per CTA instead of once per tile.
    const bool staged = qe - qs <= 32;
    const uint32_t *my_row = nullptr;  // dense: offsets per sub-tile; sparse: offsets per tile
    typename TermW<KIND>::type my_idf = 0, my_qw = 0;
    int my_dense = 0;
    uint32_t nxt_beg = 0, nxt_end = 0;
    if (staged && lane < qe - qs) {
        const int t = q_terms[qs + lane];
        my_qw = q_weights[qs + lane];
        my_idf = idf[t];
        const int32_t did = dense_id[t];
        my_dense = did >= 0;
        my_row = my_dense ? dense_ptr + (size_t)did * dense_row + w : blk_ptr + (size_t)t * n_tiles;
        if (y_first < n_y) {
            const size_t i0 = (size_t)tile_of(y_first) * (my_dense ? B2R_SUBTILES : 1);
            nxt_beg = my_row[i0];
            nxt_end = my_row[i0 + 1];
        }
    }
  for (int y = y_first; y < n_y; y += y_stride) {
    if (lane == 0) {  // clear my sub-tile's accumulators with one bulk copy; overlaps the term staging below
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(zbar_a), "r"((uint32_t)(sub * 8))
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                sc_smem_u32(acc_w)),
            "l"(g_zero_page), "r"((uint32_t)(sub * 8)), "r"(zbar_a)
            : "memory");
    }
    const int tile = tile_of(y);
    const uint32_t my_doc0 = (uint32_t)tile * (uint32_t)tile_docs + (uint32_t)w * (uint32_t)sub;
    const size_t my_sub = (size_t)tile * B2R_SUBTILES + w;
    uint32_t my_beg = nxt_beg, my_end = nxt_end;
    if (staged && my_row != nullptr && y + y_stride < n_y) {  // offsets of the next tile: used one iteration later
        const size_t i1 = (size_t)tile_of(y + y_stride) * (my_dense ? B2R_SUBTILES : 1);
        nxt_beg = my_row[i1];
        nxt_end = my_row[i1 + 1];
    }

    bool cleared = false;
    bool first = true;  // no term has touched this warp's sub-tile yet (warp-uniform)

    for (int j0 = qs; j0 < qe; j0 += 32) {
        const int nt = min(32, qe - j0);
        if (!staged) {  // long query: lane j stages term j0+j for this tile (dense: my sub-tile; sparse: the tile block)
            my_beg = my_end = 0;
            my_dense = 0;
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
        }
        if (!cleared) {  // the staging loads above were issued before this wait
            sc_mbar_wait(zbar_a, zphase);
            cleared = true;
        }
        for (int j = 0; j < nt; ++j) {
            const uint32_t beg = __shfl_sync(full, my_beg, j), end = __shfl_sync(full, my_end, j);
            if (beg == end) continue;  // warp-uniform
            const int dense = __shfl_sync(full, my_dense, j);
            const typename TermW<KIND>::type w_idf = __shfl_sync(full, my_idf, j), w_q = __shfl_sync(full, my_qw, j);
            if (first) apply_term<KIND, true>(dense, beg, end, lane, sub, my_doc0, post_doc, post_val, acc_w, w_idf, w_q);
            else apply_term<KIND, false>(dense, beg, end, lane, sub, my_doc0, post_doc, post_val, acc_w, w_idf, w_q);
            first = false;
            __syncwarp();
        }
    }

    if (!cleared) sc_mbar_wait(zbar_a, zphase);  // query without terms
    zphase ^= 1;
    if (OUT == SC_OUT_DENSE) {
        float *out = o.scores + (int64_t)ql * o.scores_stride + (int64_t)y * tile_docs + w * sub;
        for (int i = lane * 2; i < sub; i += 64) {
            double2 a = *reinterpret_cast<const double2 *>(acc_w + i);
            *reinterpret_cast<float2 *>(out + i) = make_float2(__double2float_rn(a.x), __double2float_rn(a.y));
        }
    } else if (OUT == SC_OUT_MAXIMA) {
        // one maximum per lane over the (valid) documents it reads; f32(max) == max(f32): rounding is monotone
        double m = -__longlong_as_double(0x7ff0000000000000ll);
        for (int i = lane * 2; i < sub; i += 64) {
            const double2 a = *reinterpret_cast<const double2 *>(acc_w + i);
            const uint32_t doc = my_doc0 + i;
            if (doc < o.n_docs) m = fmax(m, a.x);
            if (doc + 1 < o.n_docs) m = fmax(m, a.y);
        }
        o.scores[(int64_t)ql * o.scores_stride + (int64_t)y * SC_GROUPS_PER_TILE + w * 32 + lane] = __double2float_rn(m);
    } else if (OUT == SC_OUT_EXHAUSTIVE) {
        // every document whose key beats the k-th best so far becomes a pending candidate; the doc tile is examined
        // in chunks that always fit behind the keys already held (x_cnt <= SC_XK_CAP - chunk before each chunk)
        const int chunk = tile_docs < SC_XK_CHUNK ? tile_docs : SC_XK_CHUNK;
        for (int c0 = 0; c0 < tile_docs; c0 += chunk) {
            if (w * sub >= c0 && w * sub < c0 + chunk) {
                const uint64_t thr = x_thr;
                for (int i = lane * 2; i < sub; i += 64) {
                    const double2 a = *reinterpret_cast<const double2 *>(acc_w + i);
                    const double av[2] = {a.x, a.y};
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t doc = my_doc0 + i + c;
                        if (doc < o.n_docs) {
                            const uint64_t key = make_key(ord_f32(__double2float_rn(av[c])), o.doc_id_base + doc);
                            if (key > thr) xk[atomicAdd(&x_cnt, 1)] = key;
                        }
                    }
                }
            }
            // (one barrier: the last thread to arrive sees every append of the chunk, and x_cnt only grows)
            if (__syncthreads_or(x_cnt > SC_XK_CAP - chunk)) x_compact();
        }
    } else {
        const uint64_t thr = o.thr_keys[ql];
        const uint32_t thr_hi = (uint32_t)(thr >> 32);
        // A document can only beat thr if f32(acc) >= the threshold score.  Instead of converting all 4096
        // accumulators of the tile, they are compared in f64 against the f32 value just below the threshold
        // score: acc < pred(thr_f) implies f32(acc) <= pred(thr_f) < thr_f (rounding is monotone).  Only
        // survivors are converted and keyed (measured: -2 % kernel time).  thr_hi == 0: no threshold yet.
        // ordered encoding: -1 = next smaller f32; 0x7fffffff would be -0.0 (== +0.0 in the ranking): skip it;
        // at or below -inf (0x007fffff) there is nothing smaller: no filter
        uint32_t thr_ord = thr_hi > 0x007fffffu ? thr_hi - 1u : 0u;
        if (thr_ord == 0x7fffffffu) thr_ord = 0x7ffffffeu;
        double thr_lo = thr_ord ? (double)unord_f32(thr_ord) : -__longlong_as_double(0x7ff0000000000000ll);
        // "strictly positive scores only" (kth_of_maxima's positive floor): nothing below 2^-150 rounds to a
        // positive f32, so the untouched documents (acc == 0) never reach the conversion path
        if (thr == ((0x80000000ull << 32) | 0xFFFFFFFFull)) thr_lo = __longlong_as_double(0x3690000000000000ll);
        // (measured: testing 8 documents per step through an fmax tree is slower than this plain pair loop)
        for (int i = lane * 2; i < sub; i += 64) {
            const double2 a = *reinterpret_cast<const double2 *>(acc_w + i);
            if (a.x >= thr_lo || a.y >= thr_lo) {
                const double av[2] = {a.x, a.y};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t doc = my_doc0 + i + c;
                    if (av[c] >= thr_lo && doc < o.n_docs) {
                        const uint64_t key = make_key(ord_f32(__double2float_rn(av[c])), o.doc_id_base + doc);
                        if (key > thr) {
                            const int slot = atomicAdd(o.cand_cnt + ql, 1);
                            if (slot < o.cap) o.cand[(int64_t)ql * o.cap + slot] = key;
                        }
                    }
                }
            }
        }
    }
    // the next tile's bulk clear (async proxy) must not overtake this tile's accumulator reads (generic proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }  // tile loop
    if (OUT == SC_OUT_EXHAUSTIVE) {
        // this part's k best -> x_parts; the last part of the query to finish merges all of them
        __syncthreads();
        x_compact();
        const int k = o.k;
        uint64_t *mine = o.x_parts + ((int64_t)ql * y_stride + y_first) * k;
        for (int i = threadIdx.x; i < k; i += SC_THREADS) mine[i] = i < x_cnt ? xk[i] : 0ull;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) x_last = atomicAdd(o.x_done + ql, 1) == y_stride - 1;
        __syncthreads();
        if (!x_last) continue;   // (CTA-uniform)
        __threadfence();
        const int n = y_stride * k;   // <= SC_X_PARTS * 128 keys
        const uint64_t *all = o.x_parts + (int64_t)ql * y_stride * k;
        for (int i = threadIdx.x; i < n; i += SC_THREADS) xk[i] = __ldcg(all + i);
        if (threadIdx.x == 0) x_cnt = n;
        __syncthreads();
        x_compact();
        for (int i = threadIdx.x; i < k; i += SC_THREADS) {
            const uint64_t key = xk[i];   // (0 beyond the number of documents: sorted last)
            const int64_t at = (int64_t)ql * k + i;
            if (o.keys_final) o.keys_final[at] = key;
            if (o.idx_final) o.idx_final[at] = key ? (int64_t)(0xFFFFFFFFu - (uint32_t)key) : -1;
            if (o.val_final) o.val_final[at] = key ? unord_f32((uint32_t)(key >> 32)) : __int_as_float(0xff800000);
        }
    }
  }  // work items
}

struct ScoreLaunch {
    const b2r_index *ix;
    const int32_t *q_ptr, *q_terms;
    const float *q_weights, *idf;
    cudaStream_t st;
};

// doc tiles walked by one CTA of the scorer (B2R_SCORE_TILES_PER_CTA overrides it: tuning experiments only)
static int tiles_per_cta_default() {
    const char *e = getenv("B2R_SCORE_TILES_PER_CTA");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 64 ? v : 4;
}
static const int g_tiles_per_cta = tiles_per_cta_default();

template <int OUT>
static int launch_score(const ScoreLaunch &L, int q0, int nq, int tile_mode, int tile_step, int n_y, const ScoreOut &o) {
    if (nq == 0 || n_y == 0) return B2R_OK;
    const b2r_index *ix = L.ix;
    const size_t smem = (size_t)ix->tile_docs * sizeof(double) + (OUT == SC_OUT_EXHAUSTIVE ? (size_t)SC_XK_CAP * 8 : 0);
    // a gated launch is expected to do nothing: keep its grid tiny (each CTA walks n_y / 4 tiles if it runs)
    // tiles per CTA: enough CTAs must remain to fill the GPU a few times over (148 SMs x 6 CTAs)
    int per_cta = g_tiles_per_cta;
    while (per_cta > 1 && (int64_t)nq * (n_y / per_cta) < 148 * 6 * 4) per_cta >>= 1;
    int grid_y = (OUT == SC_OUT_DENSE && o.gate != nullptr) ? (n_y < 4 ? n_y : 4) : (n_y + per_cta - 1) / per_cta;
    // EXHAUSTIVE: a fixed grid (two 96 KB CTAs per SM) walks the (marked query, part) work items; normally there are
    // none and every CTA leaves after reading the count.  A marked query's doc tiles are split into one part per 8
    // tiles, at most SC_X_PARTS: what the step waits for is the latency of the few marked queries.
    if (OUT == SC_OUT_EXHAUSTIVE) grid_y = 1;
    dim3 grid((unsigned)(OUT == SC_OUT_EXHAUSTIVE ? 148 * 2 : nq), (unsigned)grid_y);
    if (ix->kind == B2R_KIND_BM25) {
        auto kern = score_tiles_kernel<B2R_KIND_BM25, OUT>;
        B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SC_THREADS, smem, L.st>>>(ix->post_doc, ix->post_val, ix->blk_ptr, ix->dense_id, ix->dense_ptr,
                                               ix->n_tiles, ix->tile_docs, L.q_ptr, L.q_terms, L.q_weights, L.idf, q0,
                                               tile_mode, tile_step, n_y, o);
    } else {
        auto kern = score_tiles_kernel<B2R_KIND_IMPACT, OUT>;
        B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SC_THREADS, smem, L.st>>>(ix->post_doc, ix->post_val, ix->blk_ptr, ix->dense_id, ix->dense_ptr,
                                               ix->n_tiles, ix->tile_docs, L.q_ptr, L.q_terms, L.q_weights, L.idf, q0,
                                               tile_mode, tile_step, n_y, o);
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

// ---- fused-selection plan ------------------------------------------------------------------------
constexpr int FUSED_MIN_TILES = 8;    // below this the plain score + select path is used
constexpr int FUSED_MAX_K = 128;

static bool g_fused_enabled = true;
static int g_fused_cap = 0;
// B2R_SAMPLE_STEP: every step-th doc tile forms the threshold sample (default 64; tuning experiments only)
static const int g_sample_step = [] {
    const char *e = getenv("B2R_SAMPLE_STEP");
    const int v = e ? atoi(e) : 0;
    return v >= 2 && v <= 1024 ? v : 0;
}();
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
    // every 64th tile; for k > 16 on shards of fewer than 1024 tiles every 16th (a handful of sample tiles gives a
    // loose threshold, and the longer candidate lists then cost more in topk_of_lists than the sample saves)
    p.step = (k <= 16 || ix->n_tiles >= 1024) ? 64 : 16;
    if (g_sample_step > 0) p.step = g_sample_step;
    p.n_sample = (ix->n_tiles + p.step - 1) / p.step;
    {
        const double r = (double)ix->n_tiles / p.n_sample;
        const double want = k * r + 6.0 * sqrt((double)k) * r + k;
        p.cap = 256;
        while (p.cap < want && p.cap < 4096) p.cap <<= 1;
        // beyond 4096 keys in steps of 1024: the list is staged in shared memory by topk_of_lists, whose occupancy
        // follows the capacity
        if (want > 4096) p.cap = want >= 16384 ? 16384 : (int)((want + 1023) / 1024) * 1024;
    }
    if (g_fused_cap > 0) {   // b2r_set_fused_cap: test hook (tiny lists force the exhaustive fallback)
        p.cap = g_fused_cap < k ? k : g_fused_cap;
        if (p.cap > 16384) p.cap = 16384;
    }
    p.n_groups = (int64_t)p.n_sample * SC_GROUPS_PER_TILE;
    // tile 0 is full (n_tiles >= 8) and holds min(tile_docs / 2, 256) non-empty groups
    const int groups_tile0 = ix->tile_docs / 2 < SC_GROUPS_PER_TILE ? ix->tile_docs / 2 : SC_GROUPS_PER_TILE;
    if (groups_tile0 < k) p.on = false;
    return p;
}

// workspace bytes needed to run `qc` queries in one pass.  The fused path never holds a score vector: group maxima
// of the sample, thresholds, candidate lists, and the part lists of the exhaustive fallback -- C3 (8.8M docs, top-100,
// 1024 queries): 0.2 GB.  Only the plain path (small shards, k > 128) keeps qc score rows.
static size_t pass_bytes(const b2r_index *ix, const FusedPlan &fp, int64_t qc, int k) {
    if (!fp.on) return align_up((size_t)padded_docs(ix) * 4 * (size_t)qc, 256) + topk_ws_bytes(qc, ix->n_docs, k);
    return align_up((size_t)fp.n_groups * 4 * (size_t)qc, 256) + align_up((size_t)qc * 8, 256) +
           align_up((size_t)qc * fp.cap * 8, 256) + 2 * align_up((size_t)qc * 4, 256) +
           align_up((size_t)qc * SC_X_PARTS * k * 8, 256) + align_up((size_t)(qc + 1) * 4, 256) + 256;
}

}  // namespace b2r

using namespace b2r;

// test / profiling hook: 0 disables the fused-selection path (plain score + select is used)
extern "C" void b2r_set_fused_selection(int enabled) { b2r::g_fused_enabled = enabled != 0; }
// test hook: candidate-list capacity of the fused path (0 = the plan's own); a tiny value makes every list overflow
extern "C" void b2r_set_fused_cap(int cap) { b2r::g_fused_cap = cap > 0 ? cap : 0; }

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
    float *full = nullptr, *maxima = nullptr;
    void *tk_full = nullptr;
    size_t tk_full_bytes = 0;
    uint64_t *thr = nullptr, *cand = nullptr, *x_parts = nullptr;
    int32_t *cand_cnt = nullptr, *x_done = nullptr, *marked = nullptr;
    if (fp.on) {
        maxima = static_cast<float *>(carve((size_t)fp.n_groups * 4 * (size_t)qc));
        thr = static_cast<uint64_t *>(carve((size_t)qc * 8));
        cand = static_cast<uint64_t *>(carve((size_t)qc * fp.cap * 8));
        cand_cnt = static_cast<int32_t *>(carve((size_t)qc * 4));
        x_done = static_cast<int32_t *>(carve((size_t)qc * 4));
        x_parts = static_cast<uint64_t *>(carve((size_t)qc * SC_X_PARTS * k * 8));
        marked = static_cast<int32_t *>(carve((size_t)(qc + 1) * 4));
    } else {
        full = static_cast<float *>(carve((size_t)pad * 4 * (size_t)qc));
        tk_full_bytes = topk_ws_bytes(qc, ix->n_docs, k);
        tk_full = carve(tk_full_bytes);
    }

    for (int64_t q0 = 0; q0 < n_queries; q0 += qc) {
        const int nq = (int)((n_queries - q0) < qc ? (n_queries - q0) : qc);
        uint64_t *kout = keys + q0 * k;
        if (fp.on) {
            // 1. threshold: group maxima of every step-th tile, then the k-th largest of them per query (the same
            //    kernel resets the query's candidate and fallback counters)
            // (with a packed copy of the index, steps 1-3 run on the f32 pre-filter of score_approx.cu: same launches,
            //  same buffers, bit-identical results)
            const bool approx = approx_usable(ix, k);
            if (approx) {
                rc = approx_maxima(ix, q_ptr, q_terms, q_weights, idf, (int)q0, nq, fp.step, fp.n_sample, maxima,
                                   fp.n_groups, st);
            } else {
                ScoreOut so = {};
                so.scores = maxima;
                so.scores_stride = fp.n_groups;
                so.n_docs = (uint32_t)ix->n_docs;
                rc = launch_score<SC_OUT_MAXIMA>(L, (int)q0, nq, SC_TILES_SAMPLE, fp.step, fp.n_sample, so);
            }
            if (rc) return rc;
            rc = kth_of_maxima(maxima, nq, fp.n_groups, fp.n_groups, k, false, true, thr, st, cand_cnt, x_done, marked);
            if (rc) return rc;
            // 2. every tile: score, keep only the documents that reach the threshold
            ScoreOut fo = {};
            fo.thr_keys = thr;
            fo.cand = cand;
            fo.cand_cnt = cand_cnt;
            fo.cap = fp.cap;
            fo.n_docs = (uint32_t)ix->n_docs;
            fo.doc_id_base = (uint32_t)ix->doc_id_base;
            if (g_profile) B2R_CUDA(cudaEventRecord(g_ev[0], st));
            if (approx)
                rc = approx_fused(ix, q_ptr, q_terms, q_weights, idf, (int)q0, nq, thr, cand, cand_cnt, fp.cap, st);
            else
                rc = launch_score<SC_OUT_FUSED>(L, (int)q0, nq, SC_TILES_ALL, 1, ix->n_tiles, fo);
            if (rc) return rc;
            if (g_profile) B2R_CUDA(cudaEventRecord(g_ev[1], st));
            // 3. exact top-k of the candidates, ranked keys and their decoded form in one launch; a list that
            //    overflowed or came up short marks its query (cand_cnt > cap)
            if (approx)   // (the candidates carry approximate scores: the survivors are rescored in f64 before ranking)
                rc = approx_select(ix, q_ptr, q_terms, q_weights, idf, (int)q0, nq, thr, cand, cand_cnt, fp.cap, k,
                                   kout, idx_out ? idx_out + q0 * k : nullptr, val_out ? val_out + q0 * k : nullptr,
                                   marked, st);
            else
                rc = topk_of_lists(cand, nq, fp.cap, cand_cnt, k, k, kout, st, idx_out ? idx_out + q0 * k : nullptr,
                                   val_out ? val_out + q0 * k : nullptr, marked);
            if (rc) return rc;
            // 4. exact fallback for the marked queries only (every other CTA leaves at once): exhaustive scoring with
            //    a streaming top-k in shared memory -- no score vector, no workspace that grows with the corpus
            ScoreOut xo = {};
            xo.marked = marked;
            xo.x_n_parts = ix->n_tiles / 8 < 1 ? 1 : (ix->n_tiles / 8 > SC_X_PARTS ? SC_X_PARTS : ix->n_tiles / 8);
            xo.n_docs = (uint32_t)ix->n_docs;
            xo.doc_id_base = (uint32_t)ix->doc_id_base;
            xo.k = k;
            xo.x_parts = x_parts;
            xo.x_done = x_done;
            xo.keys_final = kout;
            xo.idx_final = idx_out ? idx_out + q0 * k : nullptr;
            xo.val_final = val_out ? val_out + q0 * k : nullptr;
            rc = launch_score<SC_OUT_EXHAUSTIVE>(L, (int)q0, nq, SC_TILES_ALL, 1, ix->n_tiles, xo);
            if (rc) return rc;
            continue;
        }
        ScoreOut o = {};
        o.scores = full;
        o.scores_stride = pad;
        rc = launch_score<SC_OUT_DENSE>(L, (int)q0, nq, SC_TILES_ALL, 1, ix->n_tiles, o);
        if (rc) return rc;
        rc = topk_scores_rows(full, nq, ix->n_docs, pad, k, ix->doc_id_base, kout, tk_full, tk_full_bytes, st);
        if (rc) return rc;
    }
    if (fp.on) return B2R_OK;   // (decoded by the selection kernels)
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
