// INT8 dense similarity scan (reference: quantized_dot_product_batch,
// rag_system/core/retriever_registry.py:90-117, and the scan + argpartition of
// QuantizedEmbeddingRetriever.search, :495-515).
//
// The contraction is the one dense GEMM-shaped op of the hot path, so it runs on the 5th-generation
// tensor cores: tcgen05.mma kind::i8 (M = 128 documents x N = 128 queries x K = 32 bytes per
// instruction), operands staged by TMA into 128B-swizzled K-major shared-memory tiles (the query tile
// resident, document K-chunks through an mbarrier ring), int32 accumulators double-buffered in tensor
// memory, read back with tcgen05.ld by the epilogue warps, which apply the reference's f64 scale chain
// f32((f64(dot) * f64(qs)) * f64(ds)) -- exact int32 dots, so results are bit-identical.
// Embedding widths that are not a multiple of 128 bytes (or exceed 768) take the shared-memory tiled
// dp4a kernel below instead (same results).
// b2r_int8_scan_topk walks the corpus in document chunks: dots of one chunk go to a workspace
// tile [n_q, chunk], the streaming top-k (topk.cu) reduces it to k keys per query, and the
// per-chunk winners are merged at the end, so [n_q, n_docs] is never materialised.
#include <cuda.h>

#include "common.cuh"

#include <math.h>
#include <stdlib.h>

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


// =================================================================================================
// tcgen05 (UMMA) INT8 path
// =================================================================================================
constexpr int MM_M = 128;        // documents per tile  = UMMA M = TMEM lanes
constexpr int MM_N = 128;        // queries per tile    = UMMA N = TMEM columns (int32)
constexpr int MM_KC = 128;       // bytes of K per shared-memory chunk = one 128B swizzle row
constexpr int MM_UK = 32;        // bytes of K per tcgen05.mma kind::i8
constexpr int MM_EPI_WARPS = 16;   // 4 TMEM lane quadrants x 4 column groups of 32
constexpr int MM_THREADS = 32 * (MM_EPI_WARPS + 2);  // warps 0-15 epilogue, warp 16 TMA producer, warp 17 MMA issuer
constexpr int MM_MAX_KC = 6;     // dim <= 768
constexpr int MM_CHUNK_BYTES = MM_M * MM_KC;  // 16 KB: [128 rows][128 B], 8-row x 128 B swizzle atoms

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
// [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024 B)
// | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) @4, a/b format INT8 (1) @7/@10,
// K-major A and B, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t umma_idesc_s8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_s8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
        :
        : "r"(bar), "r"(parity)
        : "memory");
}

// ---- TMA + mbarrier helpers ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D tiled bulk tensor copy global -> shared (128B-swizzled box), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// same, delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in cta_mask
__device__ __forceinline__ void tma_load_2d_mcast(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                                  uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
          "h"(cta_mask)
        : "memory");
}
// arrive on the barrier at this offset in every CTA of cta_mask when all prior MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_mcast(uint64_t *bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {  // arrive on bar when all prior MMAs have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

constexpr int MM_MAX_STAGES = 8;
struct MmBars {
    uint64_t full[MM_MAX_STAGES];   // TMA landed a document K-chunk
    uint64_t empty[MM_MAX_STAGES];  // the MMAs that read it have completed
    uint64_t tfull[2];              // an accumulator (128 TMEM columns) is complete
    uint64_t tempty[2];             // the epilogue has drained it
    uint64_t bfull;                 // the resident query tile has landed
};

// Warp-specialised, persistent: CTA (x = query tile of 128, y) walks document tiles y, y+gridDim.y, ...
//   warp 16 lane 0: TMA producer  -- document tile K-chunks (128 docs x 128 B, SWIZZLE_128B) into a ring
//   warp 17 lane 0: MMA issuer    -- 4 x tcgen05.mma kind::i8 (M128 N128 K32) per chunk, accumulators in TMEM,
//                                    double buffered (2 x 128 columns) so the epilogue overlaps the next tile
//   warps 0-15    : epilogue      -- tcgen05.ld (one document row per thread; warp w reads TMEM lane quadrant
//                                    w % 4 and the 32 columns of group w / 4); DENSE: f64 scale chain, f32
//                                    stores; FUSED: f32 pre-filter, exact chain for the survivors only
// Query tiles are the fast grid dimension and form a thread-block cluster (up to 8 CTAs): every document
// K-chunk is fetched from L2 ONCE per cluster and TMA-multicast into the same ring slot of every CTA (the
// CTAs take turns issuing), which divides the L2->SM operand traffic -- the limiter of a 128 x 128 tile --
// by the cluster size.  A ring slot is reused only after the MMAs of ALL CTAs of the cluster have read it
// (multicast tcgen05.commit onto every CTA's `empty` barrier, arrival count = cluster size).
enum { MM_OUT_DENSE = 0, MM_OUT_FUSED = 1, MM_OUT_MAXIMA = 2 };
enum { MM_TILES_ALL = 0, MM_TILES_SAMPLE = 1 };
constexpr int MM_MAX_GROUPS = 148 * MM_M;  // MAXIMA: one running maximum per (CTA row y, document row of the tile)

struct MmOut {
    // DENSE: f32 scores, column = logical tile * 128 + row (logical == actual unless a sample is being scored)
    // MAXIMA: f32 approximate group maxima, column = blockIdx.y * 128 + row (one group = the documents a
    //         thread sees while its CTA walks the sample tiles)
    float *out;
    int64_t out_stride;
    // optional gate (both epilogues): the CTA runs only if some query of its tile has gate[q] > gate_cap
    const int32_t *gate;
    int32_t gate_cap;
    // FUSED: keep only documents whose key beats thr_keys[q] (0 = no threshold)
    const uint64_t *thr_keys;
    uint64_t *cand;      // [n_q, cap]
    int32_t *cand_cnt;   // [n_q]
    int32_t cap;
    uint32_t doc_id_base;
    // B2R_INT8_DIAG (measurement only, results are garbage): 1 = the epilogue releases every accumulator
    // unread, 2 = additionally no MMA is issued (pure TMA ring rate)
    int32_t diag;
};

template <int OUT>
__global__ void __launch_bounds__(MM_THREADS, 1)
int8_mma_kernel(const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_q, int n_q,
                int64_t n_docs, int n_kc, int n_stages, const float *__restrict__ q_scale,
                const float *__restrict__ d_scale, int tile_mode, int tile_step, int64_t n_logical, MmOut o) {
    extern __shared__ uint8_t mm_smem_raw[];
    __shared__ MmBars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ double qs_s[MM_N];
    __shared__ uint64_t thr_key_s[MM_N];
    __shared__ float2 flt_s[MM_N];  // FUSED pre-filter: {f32 query scale, lowered f32 threshold}
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(mm_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sB = smem;                          // [n_kc][128 queries][128 B]
    uint8_t *sA = smem + n_kc * MM_CHUNK_BYTES;  // [n_stages][128 docs][128 B]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * MM_N;
    const int64_t n_tiles = n_logical;  // logical tiles this launch covers; tile_of() maps them to document tiles
    auto tile_of = [&](int64_t y) -> int64_t {
        return tile_mode == MM_TILES_ALL ? y : y * tile_step;
    };
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    if (o.gate != nullptr) {  // device-side gate of the exhaustive fallback: uniform over the whole cluster
        const int cq0 = (int)(blockIdx.x - crank) * MM_N;
        bool mine = false;
        for (int i = tid; i < (int)csize * MM_N; i += MM_THREADS)
            mine |= cq0 + i < n_q && o.gate[cq0 + i] > o.gate_cap;
        if (!__syncthreads_or(mine)) return;
    }

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(2 * MM_N))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < MM_MAX_STAGES; ++i) {
            mbar_init(&bars.full[i], 1);
            mbar_init(&bars.empty[i], csize);  // one multicast commit from every CTA of the cluster
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars.tfull[i], 1);
            mbar_init(&bars.tempty[i], MM_EPI_WARPS);  // one arrival per epilogue warp
        }
        mbar_init(&bars.bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < MM_N) {
        const bool qv = q0 + tid < n_q;
        qs_s[tid] = qv ? (double)q_scale[q0 + tid] : 0.0;
        if (OUT == MM_OUT_FUSED) {
            const uint64_t thr = qv ? o.thr_keys[q0 + tid] : ~0ull;
            const uint32_t hi = (uint32_t)(thr >> 32);
            thr_key_s[tid] = thr;
            // Pre-filter in f32 (the exact f64 chain costs two 64-bit conversions per output, which paces the
            // whole kernel).  p = fl(fl(dot * qs) * ds) differs from the exact score s by < 2^-22 |s| as long as
            // nothing under/overflows (guaranteed for |qs|, |ds| in [1e-15, 1e15]: |dot| < 2^24 is exact in f32),
            // so  s >= thr  implies  p >= thr - 2^-20 |thr|.  Anything else (odd scales, no threshold yet,
            // NaN) lowers the filter to -inf: every document then takes the exact path, which is always right.
            const float qf = qv ? q_scale[q0 + tid] : 0.0f;
            const float thr_f = hi ? unord_f32(hi) : __int_as_float(0xff800000);
            const bool sane = fabsf(qf) >= 1e-15f && fabsf(qf) <= 1e15f && hi != 0 && fabsf(thr_f) <= 3e38f;
            float lo_f = sane ? __fsub_rn(__fsub_rn(thr_f, __fmul_rn(fabsf(thr_f), 9.5367431640625e-07f)), 1e-37f)
                              : __int_as_float(0xff800000);
            if (!qv) lo_f = __int_as_float(0x7f800000);  // no query in this column: nothing passes
            flt_s[tid] = make_float2(qf, lo_f);
        } else if (OUT == MM_OUT_MAXIMA) {  // an odd scale gives NaN products, which fmaxf ignores
            const float qf = qv ? q_scale[q0 + tid] : 0.0f;
            const bool sane = qv && fabsf(qf) >= 1e-15f && fabsf(qf) <= 1e15f;
            flt_s[tid] = make_float2(sane ? qf : __int_as_float(0x7fc00000), 0.0f);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrival
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == MM_EPI_WARPS) {
        if (lane == 0) {  // ---- TMA producer
            mbar_expect_tx(&bars.bfull, (uint32_t)(n_kc * MM_CHUNK_BYTES));
            for (int kc = 0; kc < n_kc; ++kc) tma_load_2d(sB + kc * MM_CHUNK_BYTES, &map_q, kc * MM_KC, q0, &bars.bfull);
            int stage = 0;
            uint32_t ph = 0, turn = 0;
            for (int64_t t = blockIdx.y; t < n_tiles; t += gridDim.y) {
                for (int kc = 0; kc < n_kc; ++kc) {
                    // slot free in EVERY CTA of the cluster (passes at once the first time round)
                    mbar_wait(smem_u32(&bars.empty[stage]), ph ^ 1);
                    mbar_expect_tx(&bars.full[stage], (uint32_t)MM_CHUNK_BYTES);
                    if (csize == 1) {
                        tma_load_2d(sA + stage * MM_CHUNK_BYTES, &map_d, kc * MM_KC, (int)(tile_of(t) * MM_M),
                                    &bars.full[stage]);
                    } else if (turn == crank) {  // my turn: one L2 read, delivered to all CTAs of the cluster
                        tma_load_2d_mcast(sA + stage * MM_CHUNK_BYTES, &map_d, kc * MM_KC, (int)(tile_of(t) * MM_M),
                                          &bars.full[stage], cmask);
                    }
                    if (++turn == csize) turn = 0;
                    if (++stage == n_stages) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == MM_EPI_WARPS + 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = umma_idesc_s8(MM_M, MM_N);
            mbar_wait(smem_u32(&bars.bfull), 0);
            int stage = 0, acc = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int64_t t = blockIdx.y; t < n_tiles; t += gridDim.y) {
                mbar_wait(smem_u32(&bars.tempty[acc]), acc_ph ^ 1);  // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MM_N);
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(smem_u32(&bars.full[stage]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int ks = 0; ks < MM_KC / MM_UK; ++ks) {
                        const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * MM_CHUNK_BYTES + ks * MM_UK));
                        const uint64_t bd = umma_desc_sw128(smem_u32(sB + kc * MM_CHUNK_BYTES + ks * MM_UK));
                        if (o.diag < 2) umma_s8(d_tmem, ad, bd, idesc, (kc | ks) ? 1u : 0u);
                    }
                    // frees the smem slot (in every CTA of the cluster) when these MMAs are done
                    if (csize == 1) umma_commit(&bars.empty[stage]);
                    else umma_commit_mcast(&bars.empty[stage], cmask);
                    if (++stage == n_stages) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
                umma_commit(&bars.tfull[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else {  // ---- epilogue warps: TMEM lane quadrant = warp % 4 (document rows), column group = warp / 4
        const int quad = warp & 3, c0 = (warp >> 2) * 32;
        int acc = 0;
        uint32_t acc_ph = 0;
        auto scale_of = [&](int64_t t) -> float {  // scale of this thread's document row in logical tile t
            const int64_t d = tile_of(t) * MM_M + quad * 32 + lane;
            return (t < n_tiles && d < n_docs) ? __ldg(d_scale + d) : 0.0f;
        };
        float dsf_next = scale_of(blockIdx.y);
        float runmax[OUT == MM_OUT_MAXIMA ? 32 : 1];
#pragma unroll
        for (int j = 0; j < (OUT == MM_OUT_MAXIMA ? 32 : 1); ++j) runmax[j] = __int_as_float(0xff800000);
        for (int64_t t = blockIdx.y; t < n_tiles; t += gridDim.y) {
            const int64_t doc = tile_of(t) * MM_M + quad * 32 + lane;
            const bool doc_ok = doc < n_docs;
            const float dsf = dsf_next;
            dsf_next = scale_of(t + gridDim.y);  // one tile ahead: its latency hides behind this tile's work
            const double ds = (double)dsf;
            const bool ds_sane = fabsf(dsf) >= 1e-15f && fabsf(dsf) <= 1e15f;
            mbar_wait(smem_u32(&bars.tfull[acc]), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (o.diag >= 1) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.tempty[acc]);
            } else {
                const int cc = 1;  // (one 32-column group per warp: its only load is also its last)
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * MM_N + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                      "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                      "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // DENSE: this warp's share of the accumulator is in registers after the second load: hand it back.
                // FUSED re-reads single columns of survivors from TMEM and releases the buffer after that.
                if (OUT != MM_OUT_FUSED && cc == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.tempty[acc]);
                }
                if (OUT == MM_OUT_DENSE) {
                    float *optr = o.out + (int64_t)(q0 + c0) * o.out_stride + (t * MM_M + quad * 32 + lane);
                    if (q0 + c0 + 32 <= n_q) {  // full column block (warp-uniform)
                        if (doc_ok) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const double sc = __dmul_rn(__dmul_rn((double)(int32_t)v[j], qs_s[c0 + j]), ds);
                                optr[(int64_t)j * o.out_stride] = __double2float_rn(sc);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (q0 + c0 + j < n_q && doc_ok) {
                                const double sc = __dmul_rn(__dmul_rn((double)(int32_t)v[j], qs_s[c0 + j]), ds);
                                optr[(int64_t)j * o.out_stride] = __double2float_rn(sc);
                            }
                        }
                    }
                } else if (OUT == MM_OUT_MAXIMA) {
                    // threshold sample: running maximum of the f32 approximation p = fl(fl(dot * qs) * ds) per
                    // (thread = document row, query column); rows with an odd scale stay out of it
                    if (doc_ok && ds_sane) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float p = __fmul_rn(__fmul_rn((float)(int32_t)v[j], flt_s[c0 + j].x), dsf);
                            runmax[j] = fmaxf(runmax[j], p);
                        }
                    }
                } else {
                    // 1. f32 pre-filter over the 32 columns: one bit per column that may beat its threshold
                    uint32_t hit = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 f = flt_s[c0 + j];
                        const float p = __fmul_rn(__fmul_rn((float)(int32_t)v[j], f.x), dsf);
                        hit |= (p < f.y) ? 0u : (1u << j);  // NaN passes
                    }
                    if (!ds_sane) hit = 0xffffffffu;
                    if (!doc_ok) hit = 0;
                    // 2. exact f64 chain for the survivors only.  The loop runs over the columns in which ANY
                    // lane has a survivor (warp-uniform, usually none or one) and re-reads that column from TMEM,
                    // so the unrolled filter above stays branch-free and the rare path is a small rolled loop.
                    uint32_t cols = __reduce_or_sync(0xffffffffu, hit);
                    while (cols) {
                        const int j = __ffs(cols) - 1;
                        cols &= cols - 1;
                        uint32_t dot;
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(dot) : "r"(taddr + (uint32_t)j));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if ((hit >> j) & 1u) {
                            const float sc =
                                __double2float_rn(__dmul_rn(__dmul_rn((double)(int32_t)dot, qs_s[c0 + j]), ds));
                            const uint64_t key = make_key(ord_f32(sc), o.doc_id_base + (uint32_t)doc);
                            if (key > thr_key_s[c0 + j]) {
                                const int q = q0 + c0 + j;
                                const int slot = atomicAdd(o.cand_cnt + q, 1);
                                if (slot < o.cap) o.cand[(int64_t)q * o.cap + slot] = key;
                            }
                        }
                    }
                    if (cc == 1) {  // done with this accumulator buffer
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.tempty[acc]);
                    }
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
        }
        if (OUT == MM_OUT_MAXIMA) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (q0 + c0 + j < n_q)
                    o.out[(int64_t)(q0 + c0 + j) * o.out_stride + (int64_t)blockIdx.y * MM_M + quad * 32 + lane] = runmax[j];
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * MM_N))
                     : "memory");
    }
}

// the ONE shape predicate of the tcgen05 path: b2r_int8_scan_workspace (no pointers yet) and the call path share it
static bool mma_dim_ok(int dim) { return dim % MM_KC == 0 && dim / MM_KC >= 1 && dim / MM_KC <= MM_MAX_KC; }
static bool mma_shape_ok(int dim, const void *q8, const void *d8) {
    return mma_dim_ok(dim) && (reinterpret_cast<uintptr_t>(q8) & 15) == 0 && (reinterpret_cast<uintptr_t>(d8) & 15) == 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// =================================================================================================
// CTA-pair variant of the fused scan (tcgen05.mma cta_group::2, M = 256 documents x N = 256 queries)
// =================================================================================================
// The 128 x 128 single-CTA tile is bound by shared-memory bandwidth: per tile the tensor core fetches 192 KB of
// operands from shared memory while TMA writes 96 KB into it (profiles: the pipeline runs at the SUM of the TMA and
// MMA times).  A CTA pair (two SMs of one TPC, cluster of 2 along x) halves both per output: CTA r holds the
// query rows [q0 + 128 r, +128) resident (its half of the N = 256 operand), streams its own 128-document tile
// (its half of the M = 256 operand) and receives a 128 x 256 accumulator in its own TMEM.  Only the leader
// (cluster rank 0) issues MMAs; TMA completions of both CTAs are counted on the leader's barriers; MMA commits
// are multicast to both; both epilogues release the accumulator on the leader's barrier.
constexpr int MP_N = 256;  // queries per pair tile = accumulator columns per CTA

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {  // same offset in CTA `cta` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
// Remote arrive without a cluster-scope release fence (the release form costs a full ERRBAR per arrival: 14 % of the
// epilogue's stall samples).  Nothing in generic memory is published by this arrival: it only tells the leader's
// MMA issuer that the TMEM reads of an accumulator have completed, and those are ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync on this side and tcgen05.fence::after_thread_sync on the other.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into MY shared memory whose completion is counted on a barrier given by its shared::cluster address
// (the leader's): the .cta_group::2 form allows destination and barrier to live in different CTAs of the pair
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, int c0, int c1,
                                                 uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_s8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {  // both CTAs' barrier at this offset
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

__global__ void __launch_bounds__(MM_THREADS, 1)
int8_mma_pair_kernel(const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_q, int n_q,
                     int64_t n_docs, int n_kc, int n_stages, const float *__restrict__ q_scale,
                     const float *__restrict__ d_scale, int64_t n_pairs, MmOut o) {
    extern __shared__ uint8_t mm_smem_raw[];
    __shared__ MmBars bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ double qs_s[MP_N];
    __shared__ uint64_t thr_key_s[MP_N];
    __shared__ float2 flt_s[MP_N];
    // survivors are staged per warp and appended to the global candidate lists 32 at a time, so that the
    // round trip of the global atomics is paid once per 32 survivors and not on every accumulator release
    __shared__ uint64_t stage_key[MM_EPI_WARPS][64];
    __shared__ int32_t stage_q[MM_EPI_WARPS][64];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(mm_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sB = smem;                          // [n_kc][my 128 queries][128 B]
    uint8_t *sA = smem + n_kc * MM_CHUNK_BYTES;  // [n_stages][my 128 docs][128 B]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();     // 0 = leader
    const int q0 = (int)(blockIdx.x >> 1) * MP_N;  // first query of the pair tile
    auto doc_tile = [&](int64_t t) -> int64_t { return 2 * t + rank; };  // my 128-document tile of pair tile t

    if (warp == 0) {  // one warp of EACH CTA of the pair takes part in the allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(2 * MP_N))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < MM_MAX_STAGES; ++i) {
            mbar_init(&bars.full[i], 1);   // (used on the leader) its own arrive.expect_tx; bytes of both CTAs
            mbar_init(&bars.empty[i], 1);  // the leader's multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars.tfull[i], 1);                    // the leader's multicast commit
            mbar_init(&bars.tempty[i], 2 * MM_EPI_WARPS);    // (used on the leader) every epilogue warp of the pair
        }
        mbar_init(&bars.bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < MP_N; i += MM_THREADS) {
        const bool qv = q0 + i < n_q;
        const float qf = qv ? q_scale[q0 + i] : 0.0f;
        qs_s[i] = (double)qf;
        const uint64_t thr = qv ? o.thr_keys[q0 + i] : ~0ull;
        const uint32_t hi = (uint32_t)(thr >> 32);
        thr_key_s[i] = thr;
        const float thr_f = hi ? unord_f32(hi) : __int_as_float(0xff800000);
        const bool sane = fabsf(qf) >= 1e-15f && fabsf(qf) <= 1e15f && hi != 0 && fabsf(thr_f) <= 3e38f;
        float lo_f = sane ? __fsub_rn(__fsub_rn(thr_f, __fmul_rn(fabsf(thr_f), 9.5367431640625e-07f)), 1e-37f)
                          : __int_as_float(0xff800000);
        if (!qv) lo_f = __int_as_float(0x7f800000);
        flt_s[i] = make_float2(qf, lo_f);  // (see int8_mma_kernel for the pre-filter's error bound)
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers and TMEM are set up before any remote signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == MM_EPI_WARPS) {
        if (lane == 0) {  // ---- TMA producer (both CTAs): completions are counted on the LEADER's barriers
            const uint32_t bfull_l = mapa_u32(smem_u32(&bars.bfull), 0);
            if (rank == 0) mbar_expect_tx(&bars.bfull, (uint32_t)(2 * n_kc * MM_CHUNK_BYTES));
            for (int kc = 0; kc < n_kc; ++kc)
                tma_load_2d_pair(sB + kc * MM_CHUNK_BYTES, &map_q, kc * MM_KC, q0 + (int)rank * MM_N, bfull_l);
            int stage = 0;
            uint32_t ph = 0;
            for (int64_t t = blockIdx.y; t < n_pairs; t += gridDim.y) {
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(smem_u32(&bars.empty[stage]), ph ^ 1);  // my slot is free (leader's MMAs done with it)
                    if (rank == 0) mbar_expect_tx(&bars.full[stage], (uint32_t)(2 * MM_CHUNK_BYTES));
                    tma_load_2d_pair(sA + stage * MM_CHUNK_BYTES, &map_d, kc * MM_KC, (int)(doc_tile(t) * MM_M),
                                     mapa_u32(smem_u32(&bars.full[stage]), 0));
                    if (++stage == n_stages) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == MM_EPI_WARPS + 1) {
        if (lane == 0 && rank == 0) {  // ---- MMA issuer (leader only): 256 x 256 x 32 per instruction
            const uint32_t idesc = umma_idesc_s8(2 * MM_M, MP_N);
            mbar_wait(smem_u32(&bars.bfull), 0);
            int stage = 0, acc = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int64_t t = blockIdx.y; t < n_pairs; t += gridDim.y) {
                mbar_wait(smem_u32(&bars.tempty[acc]), acc_ph ^ 1);  // both epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MP_N);
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(smem_u32(&bars.full[stage]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int ks = 0; ks < MM_KC / MM_UK; ++ks) {
                        const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * MM_CHUNK_BYTES + ks * MM_UK));
                        const uint64_t bd = umma_desc_sw128(smem_u32(sB + kc * MM_CHUNK_BYTES + ks * MM_UK));
                        if (o.diag < 2) umma_s8_pair(d_tmem, ad, bd, idesc, (kc | ks) ? 1u : 0u);
                    }
                    umma_commit_pair(&bars.empty[stage]);  // frees the slot in both CTAs
                    if (++stage == n_stages) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
                umma_commit_pair(&bars.tfull[acc]);  // accumulator complete in both CTAs
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else {  // ---- epilogue warps: TMEM lane quadrant = warp % 4, column group of 64 = warp / 4
        const int quad = warp & 3, cg = warp >> 2;
        int acc = 0;
        uint32_t acc_ph = 0;
        auto scale_of = [&](int64_t t) -> float {
            const int64_t d = doc_tile(t) * MM_M + quad * 32 + lane;
            return (t < n_pairs && d < n_docs) ? __ldg(d_scale + d) : 0.0f;
        };
        float dsf_next = scale_of(blockIdx.y);
        int n_staged = 0;  // survivors in my warp's staging buffer (warp-uniform)
        for (int64_t t = blockIdx.y; t < n_pairs; t += gridDim.y) {
            const int64_t doc = doc_tile(t) * MM_M + quad * 32 + lane;
            const bool doc_ok = doc < n_docs;
            const float dsf = dsf_next;
            dsf_next = scale_of(t + gridDim.y);
            const double ds = (double)dsf;
            const bool ds_sane = fabsf(dsf) >= 1e-15f && fabsf(dsf) <= 1e15f;
            mbar_wait(smem_u32(&bars.tfull[acc]), acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = o.diag >= 1 ? 2 : 0; cc < 2; ++cc) {  // (B2R_INT8_DIAG >= 1: release the accumulator unread)
                const int c0 = cg * 64 + cc * 32;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * MP_N + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                      "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                      "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t hit = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 f = flt_s[c0 + j];
                    const float p = __fmul_rn(__fmul_rn((float)(int32_t)v[j], f.x), dsf);
                    hit |= (p < f.y) ? 0u : (1u << j);  // NaN passes
                }
                if (!ds_sane) hit = 0xffffffffu;
                if (!doc_ok) hit = 0;
                uint32_t cols = __reduce_or_sync(0xffffffffu, hit);
                while (cols) {  // exact f64 chain for the survivors, column by column (re-read from TMEM)
                    const int j = __ffs(cols) - 1;
                    cols &= cols - 1;
                    uint32_t dot;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(dot) : "r"(taddr + (uint32_t)j));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    bool pass = false;
                    uint64_t key = 0;
                    if ((hit >> j) & 1u) {
                        const float sc = __double2float_rn(__dmul_rn(__dmul_rn((double)(int32_t)dot, qs_s[c0 + j]), ds));
                        key = make_key(ord_f32(sc), o.doc_id_base + (uint32_t)doc);
                        pass = key > thr_key_s[c0 + j];
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, pass);
                    if (m) {  // warp-uniform: append to my warp's staging buffer (n_staged <= 31 before, <= 63 after)
                        if (pass) {
                            const int slot = n_staged + __popc(m & ((1u << lane) - 1u));
                            stage_key[warp][slot] = key;
                            stage_q[warp][slot] = q0 + c0 + j;
                        }
                        n_staged += __popc(m);
                        __syncwarp();
                        if (n_staged >= 32) {  // flush 32 survivors: one global atomic per lane, all in flight together
                            const uint64_t fk = stage_key[warp][lane];
                            const int fq = stage_q[warp][lane];
                            const int gslot = atomicAdd(o.cand_cnt + fq, 1);
                            if (gslot < o.cap) o.cand[(int64_t)fq * o.cap + gslot] = fk;
                            const uint64_t mk = stage_key[warp][32 + lane];
                            const int mq = stage_q[warp][32 + lane];
                            __syncwarp();
                            stage_key[warp][lane] = mk;  // move the tail (n_staged - 32 <= 31 entries) to the front
                            stage_q[warp][lane] = mq;
                            n_staged -= 32;
                            __syncwarp();
                        }
                    }
                }
            }
            // done with this accumulator buffer: tell the leader's MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&bars.tempty[acc]);
                else mbar_arrive_cluster(mapa_u32(smem_u32(&bars.tempty[acc]), 0));
            }
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
        }
        if (lane < n_staged) {  // what is left in the staging buffer
            const uint64_t fk = stage_key[warp][lane];
            const int fq = stage_q[warp][lane];
            const int gslot = atomicAdd(o.cand_cnt + fq, 1);
            if (gslot < o.cap) o.cand[(int64_t)fq * o.cap + gslot] = fk;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it or read its shared memory
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * MP_N))
                     : "memory");
    }
}

static int make_rowmajor_i8_map(CUtensorMap *map, const int8_t *base, int64_t n_rows, int dim) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        B2R_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) {
            set_error("int8 scan: cuTensorMapEncodeTiled is not available from this driver");
            return B2R_ERR_UNSUPPORTED;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n_rows};  // innermost first
    const cuuint64_t gstride[1] = {(cuuint64_t)dim};                   // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)MM_KC, (cuuint32_t)MM_M};   // 128 B x 128 rows, rows past the end read as 0
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("int8 scan: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return B2R_ERR_CUDA;
    }
    return B2R_OK;
}

static int g_int8_cluster = 2;  // largest cluster the launch may use (1 = no multicast); measured: 1 = 2 > 4, 8

template <int OUT>
static int launch_int8_mma(const int8_t *q8, int n_q, const int8_t *d8, int64_t n_docs, int dim, const float *qs,
                           const float *ds, int tile_mode, int tile_step, int64_t n_logical, const MmOut &o,
                           cudaStream_t st, int *gy_out = nullptr) {
    if (n_q == 0 || n_logical == 0) return B2R_OK;
    const int n_kc = dim / MM_KC;
    // ring depth: everything the 227 KB of shared memory leave next to the resident query tile (static arrays
    // and the 1 KB alignment slack come off first); B2R_INT8_STAGES overrides it (tuning experiments only)
    int n_stages = (232448 - 4096 - 1024 - n_kc * MM_CHUNK_BYTES) / MM_CHUNK_BYTES;
    if (n_stages > MM_MAX_STAGES) n_stages = MM_MAX_STAGES;
    static const int stage_cap = [] {
        const char *e = getenv("B2R_INT8_STAGES");
        const int v = e ? atoi(e) : 0;
        return v >= 2 && v <= MM_MAX_STAGES ? v : MM_MAX_STAGES;
    }();
    if (n_stages > stage_cap) n_stages = stage_cap;
    const size_t smem = (size_t)(n_kc + n_stages) * MM_CHUNK_BYTES + 1024;
    CUtensorMap map_d, map_q;
    int rc = make_rowmajor_i8_map(&map_d, d8, n_docs, dim);
    if (rc) return rc;
    rc = make_rowmajor_i8_map(&map_q, q8, n_q, dim);
    if (rc) return rc;
    B2R_CUDA(cudaFuncSetAttribute(int8_mma_kernel<OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = (n_q + MM_N - 1) / MM_N;
    // query tiles form a cluster (power of two, at most g_int8_cluster CTAs) that shares every document K-chunk
    // by multicast; gx is padded to a multiple of it (a padded CTA sees only out-of-range queries: zero
    // operand, no output)
    int csize = 1;
    while (csize * 2 <= g_int8_cluster && csize < gx) csize <<= 1;
    gx = (gx + csize - 1) / csize * csize;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(MM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // The kernel is persistent (a CTA walks document tiles with stride gridDim.y), so the grid must be ONE
    // wave: at most as many clusters as the GPU can hold at once (a cluster lives inside one GPC, so this is
    // fewer CTAs than SMs for large clusters) -- a second wave would double the run time.
    cfg.gridDim = dim3((unsigned)gx, 1);
    int max_clusters = 0;
    B2R_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, int8_mma_kernel<OUT>, &cfg));
    if (max_clusters < 1) max_clusters = 1;
    int64_t gy = (int64_t)max_clusters / (gx / csize);
    if (gy < 1) gy = 1;
    if (gy > n_logical) gy = n_logical;
    // the MAXIMA epilogue writes maxima[n_q][MM_MAX_GROUPS]: one row of MM_M maxima per CTA row y
    if (gy > MM_MAX_GROUPS / MM_M) gy = MM_MAX_GROUPS / MM_M;
    cfg.gridDim = dim3((unsigned)gx, (unsigned)gy);
    if (gy_out) *gy_out = (int)gy;
    static const int diag = [] {
        const char *e = getenv("B2R_INT8_DIAG");
        return e ? atoi(e) : 0;
    }();
    MmOut od = o;
    od.diag = diag;
    B2R_CUDA(cudaLaunchKernelEx(&cfg, int8_mma_kernel<OUT>, map_d, map_q, n_q, n_docs, n_kc, n_stages, qs, ds, tile_mode,
                                tile_step, n_logical, od));
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

static int g_int8_pair = 1;  // 1 = the fused scan of batches > 128 queries runs on CTA pairs (cta_group::2)

// fused scan of all documents on CTA pairs; same contract as launch_int8_mma<MM_OUT_FUSED>(MM_TILES_ALL)
static int launch_int8_pair(const int8_t *q8, int n_q, const int8_t *d8, int64_t n_docs, int dim, const float *qs,
                            const float *ds, const MmOut &o, cudaStream_t st) {
    if (n_q == 0 || n_docs == 0) return B2R_OK;
    const int n_kc = dim / MM_KC;
    const int64_t n_tiles = (n_docs + MM_M - 1) / MM_M, n_pairs = (n_tiles + 1) / 2;
    // ring depth: what 227 KB leave next to my half of the query tile, the static arrays (~18.6 KB) and alignment
    int n_stages = (232448 - 20480 - 1024 - n_kc * MM_CHUNK_BYTES) / MM_CHUNK_BYTES;
    if (n_stages > MM_MAX_STAGES) n_stages = MM_MAX_STAGES;
    const size_t smem = (size_t)(n_kc + n_stages) * MM_CHUNK_BYTES + 1024;
    CUtensorMap map_d, map_q;
    int rc = make_rowmajor_i8_map(&map_d, d8, n_docs, dim);
    if (rc) return rc;
    rc = make_rowmajor_i8_map(&map_q, q8, n_q, dim);
    if (rc) return rc;
    B2R_CUDA(cudaFuncSetAttribute(int8_mma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_qt = (n_q + MP_N - 1) / MP_N;  // pair tiles along the queries; grid x = 2 CTAs per pair tile
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(MM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(2 * n_qt), 1);
    int max_clusters = 0;  // persistent kernel: one wave of CTA pairs
    B2R_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, int8_mma_pair_kernel, &cfg));
    if (max_clusters < 1) max_clusters = 1;
    int64_t gy = (int64_t)max_clusters / n_qt;
    if (gy < 1) gy = 1;
    if (gy > n_pairs) gy = n_pairs;
    cfg.gridDim = dim3((unsigned)(2 * n_qt), (unsigned)gy);
    static const int diag = [] {
        const char *e = getenv("B2R_INT8_DIAG");
        return e ? atoi(e) : 0;
    }();
    MmOut od = o;
    od.diag = diag;
    B2R_CUDA(cudaLaunchKernelEx(&cfg, int8_mma_pair_kernel, map_d, map_q, n_q, n_docs, n_kc, n_stages, qs, ds, n_pairs, od));
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

static bool g_int8_use_mma = true;

// dense scores of all documents; `gate` (optional) restricts the work to queries with gate[q] > gate_cap
static int launch_int8_dot(const int8_t *q8, int n_q, const int8_t *d8, int64_t n_docs, int dim, const float *qs,
                           const float *ds, float *out, int64_t stride, cudaStream_t st,
                           const int32_t *gate = nullptr, int32_t gate_cap = 0) {
    if (n_q == 0 || n_docs == 0) return B2R_OK;
    if (g_int8_use_mma && mma_shape_ok(dim, q8, d8)) {
        MmOut o = {};
        o.out = out;
        o.out_stride = stride;
        o.gate = gate;
        o.gate_cap = gate_cap;
        return launch_int8_mma<MM_OUT_DENSE>(q8, n_q, d8, n_docs, dim, qs, ds, MM_TILES_ALL, 1,
                                             (n_docs + MM_M - 1) / MM_M, o, st);
    }
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
    int64_t c = ((int64_t)1 << 30) / (n_q > 0 ? n_q : 1);  // score tile of at most 4 GiB
    c = (c / 4096) * 4096;
    if (c < 4096) c = 4096;
    if (c > n_docs) c = (n_docs + 3) / 4 * 4;
    return c;
}

}  // namespace b2r

using namespace b2r;

// test / profiling hooks: b2r_set_int8_mma(0) forces the dp4a kernel for every shape;
// b2r_set_int8_cluster(c) caps the thread-block cluster (TMA multicast group) of the tcgen05 kernel at c CTAs
extern "C" void b2r_set_int8_mma(int enabled) { b2r::g_int8_use_mma = enabled != 0; }
extern "C" void b2r_set_int8_pair(int enabled) { b2r::g_int8_pair = enabled != 0; }
extern "C" void b2r_set_int8_cluster(int max_cluster) {
    b2r::g_int8_cluster = max_cluster < 1 ? 1 : max_cluster > 8 ? 8 : max_cluster;
}

extern "C" int b2r_int8_dot_batch(const int8_t *q8, int32_t n_q, const int8_t *d8, int64_t n_docs, int32_t dim,
                                  const float *q_scale, const float *d_scale, float *out, void *stream) {
    B2R_CHECK_ARG(q8 && d8 && q_scale && d_scale && out && dim >= 1 && n_q >= 0 && n_docs >= 0,
                  "b2r_int8_dot_batch: bad arguments");
    B2R_CHECK_ARG((int64_t)dim * 127 * 127 < 0x7FFFFFFFll, "b2r_int8_dot_batch: dim too large for int32 dots");
    return launch_int8_dot(q8, n_q, d8, n_docs, dim, q_scale, d_scale, out, n_docs, static_cast<cudaStream_t>(stream));
}

// fused selection for the scan (same scheme as the BM25 search path in score.cu): exact top-k of a strided
// sample of 128-document tiles gives a per-query threshold, the rest of the corpus is scanned with an
// epilogue that only appends documents beating it, overflowing queries fall back to the gated dense path
struct I8Fused {
    bool on;
    int step, cap;
    int64_t n_tiles, n_sample;
};
static bool g_int8_fused = true;
// B2R_INT8_SAMPLE_STEP: stride of the threshold sample for k > 16 (tuning experiments only)
static const int g_i8_step_k100 = [] {
    const char *e = getenv("B2R_INT8_SAMPLE_STEP");
    const int v = e ? atoi(e) : 0;
    return v >= 2 && v <= 256 ? v : 16;
}();

static I8Fused i8_fused_plan(int32_t n_q, int64_t n_docs, int dim, int k, bool shape_ok) {
    I8Fused p = {};
    p.n_tiles = (n_docs + MM_M - 1) / MM_M;
    // >= 256 tiles: the first sample tiles are full, so at least 128 >= k group maxima exist
    p.on = g_int8_fused && g_int8_use_mma && shape_ok && k <= 128 && p.n_tiles >= 256 && dim > 0 && n_q >= 1;
    if (!p.on) return p;
    p.step = k <= 16 ? 32 : g_i8_step_k100;
    p.n_sample = (p.n_tiles + p.step - 1) / p.step;
    {   // a query collects ~ k * r candidates, r = n_tiles / n_sample (sigma ~ sqrt(k) * r): cap = 2^m >= mean + 6 sigma
        const double r = (double)p.n_tiles / (double)p.n_sample;
        const double want = k * r + 6.0 * sqrt((double)k) * r + k;
        p.cap = 256;
        while (p.cap < want && p.cap < 16384) p.cap <<= 1;
    }
    return p;
}

static size_t i8_fused_bytes(const I8Fused &fp, int64_t nq, int /*k*/) {
    if (!fp.on) return 0;
    return align_up((size_t)nq * MM_MAX_GROUPS * 4, 256) + align_up((size_t)nq * 8, 256) +
           align_up((size_t)nq * fp.cap * 8, 256) + align_up((size_t)nq * 4, 256) + 256;
}

// enabled: 0 = plain chunked "dense tile + select" path; otherwise the fused-selection path (default)
extern "C" void b2r_set_int8_fused(int enabled) { g_int8_fused = enabled != 0; }

extern "C" int b2r_int8_scan_workspace(int32_t n_q, int64_t n_docs, int32_t dim, int32_t k, size_t *bytes) {
    B2R_CHECK_ARG(bytes && n_q >= 0 && n_docs >= 1 && k >= 1 && k <= B2R_TOPK_MAX_FAST,
                  "b2r_int8_scan_workspace: bad arguments");
    int64_t nq = n_q > 0 ? n_q : 1;
    int64_t chunk = i8_chunk_docs(n_q, n_docs);
    int64_t n_chunks = (n_docs + chunk - 1) / chunk;
    // (a misaligned q8 / d8 pointer later turns the fused path off: the workspace is then merely larger than needed)
    const I8Fused fp = i8_fused_plan(n_q, n_docs, dim, k, mma_dim_ok(dim));
    *bytes = align_up((size_t)nq * chunk * 4, 256) + topk_ws_bytes(nq, chunk, k) +
             align_up((size_t)n_chunks * nq * k * 8, 256) + topk_keys_ws_bytes(nq, n_chunks * k, k) +
             align_up((size_t)nq * k * 8, 256) + i8_fused_bytes(fp, nq, k) + 1024;
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

    TopkOpts gate;  // empty unless the fused path ran: then only overflowed queries take the dense path below
    const I8Fused fp = i8_fused_plan(n_q, n_docs, dim, k, mma_shape_ok(dim, q8, d8));
    if (fp.on) {
        float *maxima = static_cast<float *>(carve((size_t)n_q * MM_MAX_GROUPS * 4));
        uint64_t *thr = static_cast<uint64_t *>(carve((size_t)n_q * 8));
        uint64_t *cand = static_cast<uint64_t *>(carve((size_t)n_q * fp.cap * 8));
        int32_t *cand_cnt = static_cast<int32_t *>(carve((size_t)n_q * 4));
        // 1. threshold: every step-th 128-document tile is scanned with the MAXIMA epilogue (f32 approximation,
        //    one running maximum per thread and query); the k-th largest group maximum, lowered by the
        //    approximation margin, is a lower bound of the k-th best exact score
        MmOut so = {};
        so.out = maxima;
        so.out_stride = MM_MAX_GROUPS;
        int gy = 0;
        int rc = launch_int8_mma<MM_OUT_MAXIMA>(q8, n_q, d8, n_docs, dim, q_scale, d_scale, MM_TILES_SAMPLE, fp.step,
                                                fp.n_sample, so, st, &gy);
        if (rc) return rc;
        rc = kth_of_maxima(maxima, n_q, (int64_t)gy * MM_M, MM_MAX_GROUPS, k, true, false, thr, st);
        if (rc) return rc;
        // 2. scan every tile, keep the documents whose exact key reaches the threshold
        B2R_CUDA(cudaMemsetAsync(cand_cnt, 0, (size_t)n_q * 4, st));
        MmOut fo = {};
        fo.thr_keys = thr;
        fo.cand = cand;
        fo.cand_cnt = cand_cnt;
        fo.cap = fp.cap;
        fo.doc_id_base = (uint32_t)doc_id_base;
        if (g_int8_pair && n_q > MM_N)
            rc = launch_int8_pair(q8, n_q, d8, n_docs, dim, q_scale, d_scale, fo, st);
        else
            rc = launch_int8_mma<MM_OUT_FUSED>(q8, n_q, d8, n_docs, dim, q_scale, d_scale, MM_TILES_ALL, 1, fp.n_tiles,
                                               fo, st);
        if (rc) return rc;
        // 3. exact top-k of the candidates; overflowed queries fall through to the gated exhaustive path
        rc = topk_of_lists(cand, n_q, fp.cap, cand_cnt, k, 0, keys, st);
        if (rc) return rc;
        gate.gate = cand_cnt;
        gate.gate_cap = fp.cap;
    }
    // exhaustive chunked path: the whole job without the fused path, otherwise gated to overflowed queries
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t d0 = c * chunk;
        const int64_t nd = (n_docs - d0) < chunk ? (n_docs - d0) : chunk;
        int rc = launch_int8_dot(q8, n_q, d8 + d0 * dim, nd, dim, q_scale, d_scale + d0, tile, chunk, st, gate.gate,
                                 gate.gate_cap);
        if (rc) return rc;
        rc = topk_scores_rows(tile, n_q, nd, chunk, k, doc_id_base + d0,
                              n_chunks == 1 ? keys : part + c * (int64_t)n_q * k, tk, tk_ws, st, gate);
        if (rc) return rc;
    }
    if (n_chunks > 1) {
        int rc = topk_keys_rows(part, n_q, n_chunks * k, k, k, (int64_t)n_q * k, k, keys, mg, mg_ws, st, gate);
        if (rc) return rc;
    }
    return decode_keys(keys, (int64_t)n_q * k, idx_out, val_out, nullptr, 0, k, 0, st);
}
