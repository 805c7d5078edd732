// Term-major index builder: doc-major CSR (what RetrievalService.build_bm25_index produces,
// rag_system/core/retrieval.py:176-184) -> tile-partitioned posting lists in HBM.
//
//   pass 1  count   : histogram of postings per (term, doc tile)            atomics on blk_ptr
//   pass 2  scan    : exclusive prefix sum over the n_vocab*n_tiles table   3 kernels, in place
//   pass 3  scatter : write (local doc, value) into its block (scratch copy)  atomics on the cursor
//   pass 4  order   : rank every posting inside its block by doc id (bitmap + popcount, one warp per
//                     block) and write the final arrays.  Doc-ascending blocks make the scorer's f64
//                     shared-memory read-modify-write nearly bank-conflict free for the dense head
//                     terms that carry most of the postings (ncu: 389M of 729M shared wavefronts were
//                     conflict replays with unordered blocks), and make the layout deterministic.
//
//   pass 5  dense   : sub-tile offsets for the dense terms (binary search in the doc-ordered blocks)
//   pass 6  schedule: re-order every dense (term, sub-tile) segment round-robin over the 16 accumulator
//                     bank slots, so that a half-warp's 16 postings hit 16 different slots
//
// The BM25 posting value is the query-independent factor of the reference formula
// (retrieval.py:58,70-72), evaluated with the same f64 operations in the same order:
//   u = (tf * (k1 + 1.0)) / (tf + k1 * (1.0 - b + b * dl / avgdl))
// so that scoring is  acc += (idf * u) * qtf  -- bit-identical to the reference's per-posting term.
// A document may occur at most once per term list (a CSR row without duplicate column ids, which is
// what scipy's constructor guarantees); a duplicate is reported through the build status flag.
#include "common.cuh"

#include <stdlib.h>
#include <type_traits>

namespace b2r {

constexpr int BLD_THREADS = 256;

// scratch layout: [0] int32 status flag (1 = term id out of range, 2 = duplicate (doc, term)),
// scan partials, then the unordered copy of the postings (doc u32[nnz], value f64|f32[nnz])
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = BLD_THREADS * SCAN_ITEMS;  // 4096 table entries per CTA

static size_t scan_chunks_for(size_t n_entries) { return (n_entries + SCAN_CHUNK - 1) / SCAN_CHUNK; }

__global__ void __launch_bounds__(BLD_THREADS)
build_count_kernel(const int32_t *__restrict__ indices, const int64_t *__restrict__ indptr, int64_t n_docs,
                   int32_t n_vocab, int n_tiles, int tile_shift, uint32_t *__restrict__ cnt /* table + 1 */,
                   int32_t *__restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = warp; d < n_docs; d += n_warps) {
        const int64_t beg = indptr[d], end = indptr[d + 1];
        const int64_t tile = d >> tile_shift;
        for (int64_t p = beg + lane; p < end; p += 32) {
            int32_t t = indices[p];
            if (t < 0 || t >= n_vocab) {
                *flag = 1;
                continue;
            }
            atomicAdd(&cnt[(size_t)t * n_tiles + tile], 1u);
        }
    }
}

// --- exclusive scan over `n` u32 entries, in place ---------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (BLD_THREADS / 32) ? warp_sums[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (BLD_THREADS / 32)) warp_sums[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    return warp_sums[wid] + inc - v;
}

__global__ void __launch_bounds__(BLD_THREADS)
scan_reduce_kernel(const uint32_t *__restrict__ tab, size_t n, uint32_t *__restrict__ chunk_sums) {
    __shared__ uint32_t ws[BLD_THREADS / 32];
    __shared__ uint32_t tot;
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += tab[base + i];
    block_exclusive_scan(s, ws, &tot);
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(BLD_THREADS)
scan_chunks_kernel(uint32_t *__restrict__ chunk_sums, size_t n_chunks) {
    __shared__ uint32_t ws[BLD_THREADS / 32];
    __shared__ uint32_t tot;
    uint32_t carry = 0;
    for (size_t base = 0; base < n_chunks; base += BLD_THREADS) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < n_chunks ? chunk_sums[i] : 0;
        uint32_t ex = block_exclusive_scan(v, ws, &tot);
        if (i < n_chunks) chunk_sums[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BLD_THREADS)
scan_apply_kernel(uint32_t *__restrict__ tab, size_t n, const uint32_t *__restrict__ chunk_sums) {
    __shared__ uint32_t ws[BLD_THREADS / 32];
    __shared__ uint32_t tot;
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? tab[base + i] : 0;
        s += v[i];
    }
    uint32_t run = block_exclusive_scan(s, ws, &tot) + chunk_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) tab[base + i] = run;
        run += v[i];
    }
}

template <int KIND>
__global__ void __launch_bounds__(BLD_THREADS)
build_scatter_kernel(const float *__restrict__ tf, const int32_t *__restrict__ indices,
                     const int64_t *__restrict__ indptr, const float *__restrict__ doc_len, int64_t n_docs,
                     int32_t n_vocab, int n_tiles, int tile_shift, double k1, double b, double avgdl,
                     uint32_t *__restrict__ cursor /* table + 1 */, uint32_t *__restrict__ post_doc,
                     void *__restrict__ post_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double k1p1 = __dadd_rn(k1, 1.0);
    const double one_m_b = __dsub_rn(1.0, b);
    for (int64_t d = warp; d < n_docs; d += n_warps) {
        const int64_t beg = indptr[d], end = indptr[d + 1];
        const int64_t tile = d >> tile_shift;
        double norm = 0.0;
        if (KIND == B2R_KIND_BM25) {
            // k1 * (1.0 - b + b * dl / avgdl)  == k1 * ((1.0 - b) + ((b * dl) / avgdl))
            double dl = (double)doc_len[d];
            norm = __dmul_rn(k1, __dadd_rn(one_m_b, __ddiv_rn(__dmul_rn(b, dl), avgdl)));
        }
        for (int64_t p = beg + lane; p < end; p += 32) {
            int32_t t = indices[p];
            if (t < 0 || t >= n_vocab) continue;
            uint32_t slot = atomicAdd(&cursor[(size_t)t * n_tiles + tile], 1u);
            post_doc[slot] = (uint32_t)d;
            float f = tf[p];
            if (KIND == B2R_KIND_BM25) {
                double fd = (double)f;
                static_cast<double *>(post_val)[slot] = __ddiv_rn(__dmul_rn(fd, k1p1), __dadd_rn(fd, norm));
            } else {
                static_cast<float *>(post_val)[slot] = f;
            }
        }
    }
}


// ---- pass 4: order every block by doc id -------------------------------------------------------
// A warp owns 32 consecutive table entries at a time and walks the non-empty ones.  Ranking inside a
// block is a counting sort over the tile's doc range: presence bitmap (tile_docs bits) in shared
// memory, per-word prefix popcounts, rank(d) = prefix[word] + popc(bits below d).
template <int KIND>
__global__ void __launch_bounds__(BLD_THREADS)
build_order_kernel(const uint32_t *__restrict__ blk_ptr, size_t n_blocks, int n_tiles, int tile_docs,
                   const uint32_t *__restrict__ tmp_doc, const void *__restrict__ tmp_val,
                   uint32_t *__restrict__ post_doc, void *__restrict__ post_val, int32_t *__restrict__ flag) {
    extern __shared__ uint32_t order_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int W = max(tile_docs >> 5, 32);              // bitmap words per block (>= one word per lane)
    uint32_t *bitmap = order_smem + (size_t)wib * 2 * W;
    uint32_t *prefix = bitmap + W;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const unsigned full = 0xffffffffu;

    for (size_t base = warp * 32; base < n_blocks; base += n_warps * 32) {
        const size_t e = base + lane;
        uint32_t beg = 0, cnt = 0;
        if (e < n_blocks) {
            beg = blk_ptr[e];
            cnt = blk_ptr[e + 1] - beg;
        }
        unsigned todo = __ballot_sync(full, cnt > 0);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t s = __shfl_sync(full, beg, j);
            const uint32_t n = __shfl_sync(full, cnt, j);
            const uint32_t doc0 = (uint32_t)((base + j) % (size_t)n_tiles) * (uint32_t)tile_docs;
            if (n <= 32) {
                uint32_t d = 0xFFFFFFFFu;
                double vd = 0.0;
                float vf = 0.0f;
                if ((uint32_t)lane < n) {
                    d = tmp_doc[s + lane];
                    if (KIND == B2R_KIND_BM25) vd = static_cast<const double *>(tmp_val)[s + lane];
                    else vf = static_cast<const float *>(tmp_val)[s + lane];
                }
                uint32_t rank = 0, dup = 0;
                for (uint32_t i = 0; i < n; ++i) {
                    uint32_t di = __shfl_sync(full, d, (int)i);
                    rank += (di < d);
                    dup += (di == d);
                }
                if ((uint32_t)lane < n) {
                    if (dup > 1) *flag = 2;
                    post_doc[s + rank] = d;
                    if (KIND == B2R_KIND_BM25) static_cast<double *>(post_val)[s + rank] = vd;
                    else static_cast<float *>(post_val)[s + rank] = vf;
                }
                continue;
            }
            for (int w = lane; w < W; w += 32) bitmap[w] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32) {
                uint32_t l = tmp_doc[s + i] - doc0;
                atomicOr(&bitmap[l >> 5], 1u << (l & 31));
            }
            __syncwarp();
            uint32_t running = 0;
            for (int w0 = 0; w0 < W; w0 += 32) {
                uint32_t c = __popc(bitmap[w0 + lane]);
                uint32_t inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(full, inc, o);
                    if (lane >= o) inc += t;
                }
                prefix[w0 + lane] = running + inc - c;
                running += __shfl_sync(full, inc, 31);
            }
            if (running != n && lane == 0) *flag = 2;   // two postings of one (doc, term)
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32) {
                uint32_t d = tmp_doc[s + i];
                uint32_t l = d - doc0;
                uint32_t r = prefix[l >> 5] + __popc(bitmap[l >> 5] & ((1u << (l & 31)) - 1u));
                post_doc[s + r] = d;
                if (KIND == B2R_KIND_BM25)
                    static_cast<double *>(post_val)[s + r] = static_cast<const double *>(tmp_val)[s + i];
                else
                    static_cast<float *>(post_val)[s + r] = static_cast<const float *>(tmp_val)[s + i];
            }
            __syncwarp();
        }
    }
}

// ---- pass 5: sub-tile offsets for dense terms ---------------------------------------------------
// A term is dense when it averages >= B2R_DENSE_MIN_PER_TILE postings per tile; at most
// nnz / (B2R_DENSE_MIN_PER_TILE * n_tiles) terms can be, which bounds the table.
__global__ void __launch_bounds__(BLD_THREADS)
dense_select_kernel(const uint32_t *__restrict__ blk_ptr, int32_t n_vocab, int n_tiles, uint32_t min_df,
                    int32_t n_dense_max, int32_t *__restrict__ dense_id, int32_t *__restrict__ counter) {
    int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vocab) return;
    uint32_t df = blk_ptr[(size_t)(t + 1) * n_tiles] - blk_ptr[(size_t)t * n_tiles];
    int32_t id = -1;
    if (df >= min_df) {
        id = atomicAdd(counter, 1);
        if (id >= n_dense_max) id = -1;  // cannot happen (counting argument); stay safe
    }
    dense_id[t] = id;
}

// One warp per 4 tiles of a dense term: lane = (tile offset << 3) | sub-tile; each lane binary-searches
// the first posting of its sub-tile in the doc-ordered block.
__global__ void __launch_bounds__(BLD_THREADS)
dense_fill_kernel(const uint32_t *__restrict__ blk_ptr, const uint32_t *__restrict__ post_doc,
                  const int32_t *__restrict__ dense_id, int32_t n_vocab, int n_tiles, int tile_docs,
                  uint32_t *__restrict__ dense_ptr) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x;
    const int32_t id = dense_id[t];
    if (id < 0) return;
    const int sub_docs = tile_docs / B2R_SUBTILES;
    const size_t row = (size_t)id * ((size_t)n_tiles * B2R_SUBTILES + 1);
    const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int T0 = warp * 4; T0 < n_tiles; T0 += n_warps * 4) {
        const int T = T0 + (lane >> 3), s = lane & 7;
        if (T >= n_tiles) continue;
        uint32_t lo = blk_ptr[(size_t)t * n_tiles + T], hi = blk_ptr[(size_t)t * n_tiles + T + 1];
        const uint32_t target = (uint32_t)T * (uint32_t)tile_docs + (uint32_t)s * (uint32_t)sub_docs;
        while (lo < hi) {  // lower_bound(target)
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (post_doc[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        dense_ptr[row + (size_t)T * B2R_SUBTILES + s] = lo;
    }
    if (threadIdx.x == 0) dense_ptr[row + (size_t)n_tiles * B2R_SUBTILES] = blk_ptr[(size_t)(t + 1) * n_tiles];
}

// ---- pass 6: bank schedule of the dense sub-tile segments -----------------------------------------
// The scorer's hot loop is a warp doing acc[doc] += x on f64 accumulators in shared memory, lane i of a
// half-warp taking posting i of a group of 16 consecutive postings: the access is conflict-free when
// the 16 documents are distinct mod 16 (16 eight-byte slots span the 32 banks).  With doc-ascending
// segments a group of density rho spans ~16/rho documents and replays ~2x (ncu: 163 M of 461 M shared
// wavefronts).  The order of postings INSIDE a (term, sub-tile) segment is free (a document occurs at most
// once per term, and the reference's summation order is across terms), so each segment is re-ordered
// round-robin over the 16 residue classes: posting number j of class r goes to row j, rows are packed.
// Every full row is conflict-free; only the tail, where some classes are exhausted, still replays.
constexpr int SCHED_THREADS = 128;
constexpr int SCHED_SLOTS_MAX = 32;
// SLOTS = 16: residue classes doc mod 16 (8-byte accumulator slots per 128-byte bank row: the f64 scorer).
// SLOTS = 32: classes doc mod 32, rows in class order -- a full row of 32 postings hits 32 different 4-byte banks
// (the f32 pre-filter of score_approx.cu) AND its two halves (classes 0..15, 16..31) are distinct mod 16, so the
// f64 scorer's half-warps stay conflict-free as well.

template <int KIND, int SLOTS>
__global__ void __launch_bounds__(SCHED_THREADS)
bank_schedule_kernel(const int32_t *__restrict__ dense_id, const uint32_t *__restrict__ dense_ptr, int n_tiles,
                     int tile_docs, uint32_t *__restrict__ post_doc, void *__restrict__ post_val) {
    using val_t = typename std::conditional<KIND == B2R_KIND_BM25, double, float>::type;
    extern __shared__ __align__(16) unsigned char sched_smem[];
    const int32_t id = dense_id[blockIdx.x];
    if (id < 0) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, n_warps = SCHED_THREADS / 32;
    const int sub = tile_docs / B2R_SUBTILES, words = sub >> 5;
    const size_t per_warp = (size_t)sub * (sizeof(val_t) + 4) + (size_t)words * 4 + SCHED_SLOTS_MAX * 4;
    unsigned char *mine = sched_smem + (size_t)wib * ((per_warp + 15) / 16 * 16);
    val_t *s_val = reinterpret_cast<val_t *>(mine);
    uint32_t *s_doc = reinterpret_cast<uint32_t *>(mine + (size_t)sub * sizeof(val_t));
    uint32_t *bitmap = s_doc + sub;
    uint32_t *cls_cnt = bitmap + words;
    val_t *g_val = static_cast<val_t *>(post_val);
    const size_t n_seg = (size_t)n_tiles * B2R_SUBTILES;
    const uint32_t *row = dense_ptr + (size_t)id * (n_seg + 1);
    for (size_t seg = wib; seg < n_seg; seg += n_warps) {
        const uint32_t lo = row[seg], n = row[seg + 1] - lo;
        if (n <= (uint32_t)SLOTS || n > (uint32_t)sub) continue;  // one group: nothing to gain (n > sub: bad input, flagged)
        const uint32_t doc0 = (uint32_t)seg * (uint32_t)sub;
        for (int w = lane; w < words; w += 32) bitmap[w] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t d = post_doc[lo + i];
            s_doc[i] = d;
            s_val[i] = g_val[lo + i];
            const uint32_t l = (d - doc0) & (uint32_t)(sub - 1);
            atomicOr(&bitmap[l >> 5], 1u << (l & 31));
        }
        __syncwarp();
        if (lane < SLOTS) {
            const uint32_t m = SLOTS == 16 ? (0x00010001u << lane) : (1u << lane);
            uint32_t c = 0;
            for (int w = 0; w < words; ++w) c += __popc(bitmap[w] & m);
            cls_cnt[lane] = c;
        }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t d = s_doc[i];
            const uint32_t l = (d - doc0) & (uint32_t)(sub - 1);
            const uint32_t r = l & (uint32_t)(SLOTS - 1), wl = l >> 5;
            const uint32_t cls = SLOTS == 16 ? (0x00010001u << r) : (1u << r);
            uint32_t j = 0;  // postings of my class that precede me
            for (uint32_t w = 0; w < wl; ++w) j += __popc(bitmap[w] & cls);
            if (SLOTS == 16 && (l & 16)) j += (bitmap[wl] >> r) & 1u;
            uint32_t pos = 0;  // rows 0..j-1 in full, then the classes below mine that reach row j
#pragma unroll
            for (int c = 0; c < SLOTS; ++c) {
                const uint32_t cc = cls_cnt[c];
                pos += min(cc, j) + ((uint32_t)c < r && cc > j ? 1u : 0u);
            }
            if (pos < n) {
                post_doc[lo + pos] = d;
                g_val[lo + pos] = s_val[i];
            }
        }
        __syncwarp();
    }
}

static int g_bank_schedule = 32;   // 0 = off, 16 / 32 = residue classes (b2r_set_bank_schedule)

template <int KIND, int SLOTS>
static int launch_bank_schedule(const b2r_index *ix, size_t smem, cudaStream_t st) {
    B2R_CUDA(cudaFuncSetAttribute(bank_schedule_kernel<KIND, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    bank_schedule_kernel<KIND, SLOTS><<<(unsigned)ix->n_vocab, SCHED_THREADS, smem, st>>>(
        ix->dense_id, ix->dense_ptr, ix->n_tiles, ix->tile_docs, ix->post_doc, ix->post_val);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

// postings per tile from which a term gets sub-tile offsets (B2R_DENSE_MIN overrides it: tuning experiments only)
static int dense_min_per_tile() {
    static const int v = [] {
        const char *e = getenv("B2R_DENSE_MIN");
        const int x = e ? atoi(e) : 0;
        return x >= 1 && x <= 4096 ? x : B2R_DENSE_MIN_PER_TILE;
    }();
    return v;
}

static int tile_shift_of(int tile_docs) {
    int s = 0;
    while ((1 << s) < tile_docs) ++s;
    return s;
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_index_sizes_for(int64_t nnz, int64_t n_docs, int32_t n_vocab, int32_t tile_docs, int32_t kind,
                                   b2r_index_sizes *out) {
    B2R_CHECK_ARG(out, "b2r_index_sizes_for: null out");
    B2R_CHECK_ARG(nnz >= 0 && nnz < 0xFFFF0000ll, "index: nnz=%lld must be < 2^32 - 65536 per shard", (long long)nnz);
    B2R_CHECK_ARG(n_docs >= 1 && n_docs < 0xFFFFFFFFll && n_vocab >= 1, "index: bad n_docs/n_vocab");
    B2R_CHECK_ARG(tile_docs >= 256 && tile_docs <= 16384 && (tile_docs & (tile_docs - 1)) == 0,
                  "index: tile_docs=%d must be a power of two in [256,16384]", tile_docs);
    B2R_CHECK_ARG(kind == B2R_KIND_BM25 || kind == B2R_KIND_IMPACT, "index: unknown kind %d", kind);
    int64_t n_tiles = (n_docs + tile_docs - 1) / tile_docs;
    B2R_CHECK_ARG(n_tiles <= 65535, "index: %lld doc tiles exceed 65535; raise tile_docs", (long long)n_tiles);
    size_t entries = (size_t)n_vocab * (size_t)n_tiles + 1;
    out->post_doc_bytes = align_up((size_t)(nnz > 0 ? nnz : 1) * 4, 256);
    out->post_val_bytes = align_up((size_t)(nnz > 0 ? nnz : 1) * (kind == B2R_KIND_BM25 ? 8 : 4), 256);
    out->blk_ptr_bytes = align_up(entries * 4, 256);
    out->scratch_bytes = 256 + align_up(scan_chunks_for(entries) * 4, 256) + out->post_doc_bytes + out->post_val_bytes;
    out->n_dense_max = nnz / ((int64_t)b2r::dense_min_per_tile() * n_tiles) + 1;
    out->dense_id_bytes = align_up((size_t)n_vocab * 4, 256);
    out->dense_ptr_bytes = align_up((size_t)out->n_dense_max * ((size_t)n_tiles * B2R_SUBTILES + 1) * 4, 256);
    return B2R_OK;
}

extern "C" int b2r_index_build(const b2r_index *ix, const float *tf, const int32_t *indices, const int64_t *indptr,
                               const float *doc_len, double k1, double b, double avgdl, void *scratch,
                               size_t scratch_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(ix && ix->post_doc && ix->post_val && ix->blk_ptr && ix->dense_id && ix->dense_ptr,
                  "b2r_index_build: index buffers not set");
    B2R_CHECK_ARG(indptr && scratch, "b2r_index_build: null input");
    B2R_CHECK_ARG(ix->nnz == 0 || (tf && indices), "b2r_index_build: null postings");
    B2R_CHECK_ARG(ix->kind != B2R_KIND_BM25 || doc_len, "b2r_index_build: BM25 index needs doc_len");
    b2r_index_sizes sz;
    int rc = b2r_index_sizes_for(ix->nnz, ix->n_docs, ix->n_vocab, ix->tile_docs, ix->kind, &sz);
    if (rc) return rc;
    B2R_CHECK_ARG(ix->n_tiles == (ix->n_docs + ix->tile_docs - 1) / ix->tile_docs, "b2r_index_build: n_tiles mismatch");
    B2R_CHECK_ARG(ix->n_dense_max == sz.n_dense_max, "b2r_index_build: n_dense_max mismatch");
    if (scratch_bytes < sz.scratch_bytes) {
        set_error("b2r_index_build: scratch too small (%zu < %zu)", scratch_bytes, sz.scratch_bytes);
        return B2R_ERR_WORKSPACE;
    }
    if (ix->kind == B2R_KIND_BM25)
        B2R_CHECK_ARG(avgdl > 0.0 || ix->nnz == 0, "b2r_index_build: avgdl must be positive");

    const size_t entries = (size_t)ix->n_vocab * (size_t)ix->n_tiles + 1;
    const size_t n_chunks = scan_chunks_for(entries);
    int32_t *flag = static_cast<int32_t *>(scratch);
    uint32_t *chunk_sums = reinterpret_cast<uint32_t *>(static_cast<char *>(scratch) + 256);
    const int shift = tile_shift_of(ix->tile_docs);
    char *tmp_base = static_cast<char *>(scratch) + 256 + align_up(n_chunks * 4, 256);
    uint32_t *tmp_doc = reinterpret_cast<uint32_t *>(tmp_base);
    void *tmp_val = tmp_base + sz.post_doc_bytes;

    B2R_CUDA(cudaMemsetAsync(scratch, 0, 256, st));
    B2R_CUDA(cudaMemsetAsync(ix->blk_ptr, 0, entries * 4, st));
    // one warp per document row; enough CTAs to fill the machine several times over
    int64_t warps_needed = ix->n_docs;
    int64_t blocks = (warps_needed * 32 + BLD_THREADS - 1) / BLD_THREADS;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    // counts go to table[1 + i] so that, after the exclusive scan and the scatter's atomic cursor
    // walk, table[i] = start and table[i + 1] = end of block i.
    build_count_kernel<<<(unsigned)blocks, BLD_THREADS, 0, st>>>(indices, indptr, ix->n_docs, ix->n_vocab, ix->n_tiles,
                                                                shift, ix->blk_ptr + 1, flag);
    B2R_LAUNCH_CHECK();
    const size_t n_scan = entries - 1;
    if (n_scan > 0) {
        scan_reduce_kernel<<<(unsigned)n_chunks, BLD_THREADS, 0, st>>>(ix->blk_ptr + 1, n_scan, chunk_sums);
        B2R_LAUNCH_CHECK();
        scan_chunks_kernel<<<1, BLD_THREADS, 0, st>>>(chunk_sums, scan_chunks_for(n_scan));
        B2R_LAUNCH_CHECK();
        scan_apply_kernel<<<(unsigned)scan_chunks_for(n_scan), BLD_THREADS, 0, st>>>(ix->blk_ptr + 1, n_scan, chunk_sums);
        B2R_LAUNCH_CHECK();
    }
    if (ix->kind == B2R_KIND_BM25)
        build_scatter_kernel<B2R_KIND_BM25><<<(unsigned)blocks, BLD_THREADS, 0, st>>>(
            tf, indices, indptr, doc_len, ix->n_docs, ix->n_vocab, ix->n_tiles, shift, k1, b, avgdl, ix->blk_ptr + 1,
            tmp_doc, tmp_val);
    else
        build_scatter_kernel<B2R_KIND_IMPACT><<<(unsigned)blocks, BLD_THREADS, 0, st>>>(
            tf, indices, indptr, doc_len, ix->n_docs, ix->n_vocab, ix->n_tiles, shift, k1, b, avgdl, ix->blk_ptr + 1,
            tmp_doc, tmp_val);
    B2R_LAUNCH_CHECK();
    {
        const size_t n_blocks = entries - 1;
        const int words = ix->tile_docs / 32 > 32 ? ix->tile_docs / 32 : 32;
        const size_t smem = (size_t)(BLD_THREADS / 32) * 2 * words * sizeof(uint32_t);
        size_t want = (n_blocks + 32 * (BLD_THREADS / 32) - 1) / (32 * (BLD_THREADS / 32));
        unsigned ob = (unsigned)(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
        if (ix->kind == B2R_KIND_BM25) {
            B2R_CUDA(cudaFuncSetAttribute(build_order_kernel<B2R_KIND_BM25>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            build_order_kernel<B2R_KIND_BM25><<<ob, BLD_THREADS, smem, st>>>(ix->blk_ptr, n_blocks, ix->n_tiles,
                                                                            ix->tile_docs, tmp_doc, tmp_val,
                                                                            ix->post_doc, ix->post_val, flag);
        } else {
            B2R_CUDA(cudaFuncSetAttribute(build_order_kernel<B2R_KIND_IMPACT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            build_order_kernel<B2R_KIND_IMPACT><<<ob, BLD_THREADS, smem, st>>>(ix->blk_ptr, n_blocks, ix->n_tiles,
                                                                              ix->tile_docs, tmp_doc, tmp_val,
                                                                              ix->post_doc, ix->post_val, flag);
        }
        B2R_LAUNCH_CHECK();
    }
    {
        int32_t *counter = flag + 1;  // scratch[4..8): number of dense terms
        const uint32_t min_df = (uint32_t)dense_min_per_tile() * (uint32_t)ix->n_tiles;
        dense_select_kernel<<<(unsigned)((ix->n_vocab + BLD_THREADS - 1) / BLD_THREADS), BLD_THREADS, 0, st>>>(
            ix->blk_ptr, ix->n_vocab, ix->n_tiles, min_df, ix->n_dense_max, ix->dense_id, counter);
        B2R_LAUNCH_CHECK();
        dense_fill_kernel<<<(unsigned)ix->n_vocab, BLD_THREADS, 0, st>>>(ix->blk_ptr, ix->post_doc, ix->dense_id,
                                                                         ix->n_vocab, ix->n_tiles, ix->tile_docs,
                                                                         ix->dense_ptr);
        B2R_LAUNCH_CHECK();
        if (g_bank_schedule) {
            const int sub = ix->tile_docs / B2R_SUBTILES;
            const size_t vb = ix->kind == B2R_KIND_BM25 ? 8 : 4;
            const size_t per_warp = ((size_t)sub * (vb + 4) + (size_t)(sub >> 5) * 4 + SCHED_SLOTS_MAX * 4 + 15) / 16 * 16;
            const size_t smem = per_warp * (SCHED_THREADS / 32);
            // 32 classes need sub-tiles of at least 64 documents to be worth a row; smaller tiles keep 16
            const bool wide = g_bank_schedule == 32 && sub >= 64;
            int rc2;
            if (ix->kind == B2R_KIND_BM25)
                rc2 = wide ? launch_bank_schedule<B2R_KIND_BM25, 32>(ix, smem, st)
                           : launch_bank_schedule<B2R_KIND_BM25, 16>(ix, smem, st);
            else
                rc2 = wide ? launch_bank_schedule<B2R_KIND_IMPACT, 32>(ix, smem, st)
                           : launch_bank_schedule<B2R_KIND_IMPACT, 16>(ix, smem, st);
            if (rc2) return rc2;
        }
    }
    return B2R_OK;
}

// test / profiling hook: 0 = keep the dense segments doc-ascending (no bank schedule); applies to later builds
// (16 = residue classes doc mod 16 only, the round-1 schedule; any other non-zero value = the default, 32 classes)
extern "C" void b2r_set_bank_schedule(int enabled) { b2r::g_bank_schedule = enabled == 0 ? 0 : (enabled == 16 ? 16 : 32); }

extern "C" int b2r_index_build_status(const void *scratch, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t flag = 0;
    B2R_CUDA(cudaMemcpyAsync(&flag, scratch, sizeof(flag), cudaMemcpyDeviceToHost, st));
    B2R_CUDA(cudaStreamSynchronize(st));
    if (flag == 1) {
        set_error("b2r_index_build: a term id in `indices` is outside [0, n_vocab)");
        return B2R_ERR_DATA;
    }
    if (flag) {
        set_error("b2r_index_build: a CSR row lists the same term id twice (sum duplicates first)");
        return B2R_ERR_DATA;
    }
    return B2R_OK;
}
