// Per-row top-k selection (reference: fast_topk_selection, rag_system/core/retrieval.py:79-92).
//
// Streaming threshold select: a CTA streams one segment of a row through registers, keeps the
// current k best keys sorted at the front of a 2048-entry shared array and appends only elements
// that beat the running k-th best ("tau").  Nothing is ever fully sorted: only the k best plus the
// few hundred pending candidates go through a bitonic network.  The first sub-chunk bootstraps tau
// from per-thread group maxima (the k-th largest of >= k disjoint group maxima is a lower bound of
// the k-th largest element).  Scores stay f32 in registers: the usual cost per element is a quarter of
// an LDG.128 and one FMNMX, keys are only built for sub-chunks that hold a candidate, so the kernel
// is HBM-bound: 4 bytes per score.
//
// Ranking rule and key layout: common.cuh.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

namespace b2r {

constexpr int TK_THREADS = 256;
constexpr int TK_SN = 2048;       // shared key array (16 KB): [0,k) best, [k,SN) pending candidates
constexpr int TK_TARGET_CTAS = 148 * 8;

// Merge the c pending candidates into the sorted best list; returns the new threshold.
__device__ __forceinline__ uint64_t tk_flush(uint64_t *arr, int *cnt, int k, int c, uint64_t tau) {
    const int tid = threadIdx.x;
    const int total = k + c;
    int P = 2;
    while (P < total) P <<= 1;
    for (int i = total + tid; i < P; i += TK_THREADS) arr[i] = 0;
    __syncthreads();
    bitonic_sort_desc<TK_THREADS>(arr, P);
    uint64_t kth = arr[k - 1];
    if (tid == 0) *cnt = 0;
    __syncthreads();
    return kth > tau ? kth : tau;
}

// Element j of row r lives at  (j / piece_len) * piece_stride + r * row_stride + j % piece_len
// (piece_len >= n for an ordinary [rows, n] array; piece_len = k for shard-gathered [W, Q, k]).
template <bool KEYS_IN, int E>
__global__ void __launch_bounds__(TK_THREADS)
topk_stream_kernel(const float *__restrict__ scores, const uint64_t *__restrict__ keys_in, int64_t n,
                   int64_t row_stride, int64_t piece_len, int64_t piece_stride, uint32_t id_base, int k,
                   int64_t seg_len, uint64_t *__restrict__ out, int chunk_shift, uint32_t chunk_stride,
                   const int32_t *__restrict__ gate, int gate_cap) {
    __shared__ uint64_t arr[TK_SN];
    __shared__ int cnt;
    constexpr int SUB = TK_THREADS * E;
    const int tid = threadIdx.x;
    const int64_t row = blockIdx.x;
    if (gate != nullptr && gate[row] <= gate_cap) return;  // uniform: this row is not ours to compute
    const int seg = blockIdx.y, n_segs = gridDim.y;
    const int64_t seg_beg = (int64_t)seg * seg_len;
    const int64_t seg_end = min(n, seg_beg + seg_len);
    const int CAP = TK_SN - k;
    const int flush_at = max(1, k >> 2);  // merge pending candidates early: a fresh threshold keeps the fast path hot

    for (int i = tid; i < TK_SN; i += TK_THREADS) arr[i] = 0;
    if (tid == 0) cnt = 0;
    __syncthreads();

    uint64_t tau = 0;
    bool need_boot = (seg_end - seg_beg) > (int64_t)CAP;
    const float *rp = KEYS_IN ? nullptr : scores + row * row_stride;
    const bool vec_ok = !KEYS_IN && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0) && ((seg_len & 3) == 0);
    const bool single_piece = piece_len >= n;

    for (int64_t base = seg_beg; base < seg_end; base += SUB) {
        uint64_t key[E];
        if (KEYS_IN) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int64_t j = base + (int64_t)e * TK_THREADS + tid;
                uint64_t v = 0;
                if (j < seg_end) {
                    int64_t off = single_piece ? (row * row_stride + j)
                                               : ((j / piece_len) * piece_stride + row * row_stride + (j % piece_len));
                    v = keys_in[off];
                }
                key[e] = v;
            }
        } else {
            // scores stay f32 in registers; the common case (nothing beats the threshold) costs one
            // FSETP per element.  Keys are only built for sub-chunks that hold a candidate.
            float f[E];
            const bool vec = vec_ok && base + SUB <= seg_end;
            if (vec) {
#pragma unroll
                for (int v = 0; v < E / 4; ++v) {
                    float4 x = ldg_stream_f4(rp + base + ((int64_t)v * TK_THREADS + tid) * 4);
                    f[4 * v + 0] = x.x;
                    f[4 * v + 1] = x.y;
                    f[4 * v + 2] = x.z;
                    f[4 * v + 3] = x.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    int64_t j = base + (int64_t)e * TK_THREADS + tid;
                    f[e] = (j < seg_end) ? __ldg(rp + j) : __int_as_float(0x7fc00000);  // NaN = absent
                }
            }
            const uint32_t tau_hi = (uint32_t)(tau >> 32);
            const bool all = need_boot || tau_hi == 0;  // no usable threshold yet: every element is a candidate
            const float tau_f = unord_f32(tau_hi);      // an element can only pass if score >= tau_f
            if (!all) {
                float m = f[0];
#pragma unroll
                for (int e = 1; e < E; ++e) m = fmaxf(m, f[e]);  // fmaxf drops NaN
                if (!__syncthreads_or(m >= tau_f)) continue;
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int64_t j = vec ? base + ((int64_t)(e >> 2) * TK_THREADS + tid) * 4 + (e & 3)
                                : base + (int64_t)e * TK_THREADS + tid;
                const bool c = (all || f[e] >= tau_f) && j < seg_end;
                uint32_t g = (uint32_t)j;
                if (chunk_shift) g = (uint32_t)(j >> chunk_shift) * chunk_stride + (g & ((1u << chunk_shift) - 1u));
                key[e] = c ? make_key(ord_f32(f[e]), id_base + g) : 0ull;
            }
        }

        if (need_boot) {  // uniform; first sub-chunk of a long segment
            need_boot = false;
            int G = 1;
            while (G * TK_THREADS < k) G <<= 1;  // k <= 1024 -> G <= 4
            uint64_t m[4] = {0, 0, 0, 0};
#pragma unroll
            for (int e = 0; e < E; ++e) m[e & 3] = key[e] > m[e & 3] ? key[e] : m[e & 3];
            if (G <= 2) {
                m[0] = m[0] > m[2] ? m[0] : m[2];
                m[1] = m[1] > m[3] ? m[1] : m[3];
            }
            if (G == 1) m[0] = m[0] > m[1] ? m[0] : m[1];
            arr[tid] = m[0];
            if (G >= 2) arr[TK_THREADS + tid] = m[1];
            if (G >= 4) {
                arr[2 * TK_THREADS + tid] = m[2];
                arr[3 * TK_THREADS + tid] = m[3];
            }
            __syncthreads();
            bitonic_sort_desc<TK_THREADS>(arr, G * TK_THREADS);
            uint64_t t0 = arr[k - 1];
            __syncthreads();
            tau = t0 ? t0 - 1 : 0;
            for (int i = tid; i < G * TK_THREADS; i += TK_THREADS) arr[i] = 0;
            __syncthreads();
        }

        uint32_t pend = 0;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (key[e] > tau) pend |= 1u << e;
        if (!__syncthreads_or(pend != 0)) continue;

        while (true) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if ((pend >> e) & 1u) {
                    int slot = atomicAdd(&cnt, 1);
                    if (slot < CAP) {
                        arr[k + slot] = key[e];
                        pend &= ~(1u << e);
                    }
                }
            }
            __syncthreads();
            const int c = cnt;
            __syncthreads();
            const bool over = c > CAP;
            if (over || c >= flush_at) {
                tau = tk_flush(arr, &cnt, k, over ? CAP : c, tau);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (((pend >> e) & 1u) && key[e] <= tau) pend &= ~(1u << e);
            }
            if (!over) break;
        }
    }

    __syncthreads();
    const int c = cnt;
    __syncthreads();
    if (c > 0) tk_flush(arr, &cnt, k, c, tau);
    uint64_t *o = out + (row * n_segs + seg) * (int64_t)k;
    for (int i = tid; i < k; i += TK_THREADS) o[i] = arr[i];
}

// ------------------------------------------------------------------------------------------ plan
struct TkPlan {
    int n_segs;
    int64_t seg_len;
};

static TkPlan tk_plan(int64_t n_rows, int64_t n, int sub) {
    int64_t max_segs = (n + sub - 1) / sub;
    if (max_segs < 1) max_segs = 1;
    int64_t want = (TK_TARGET_CTAS + n_rows - 1) / n_rows;
    if (want < 1) want = 1;
    int64_t segs = want < max_segs ? want : max_segs;
    int64_t seg_len = (n + segs - 1) / segs;
    seg_len = (seg_len + sub - 1) / sub * sub;
    if (seg_len < sub) seg_len = sub;
    segs = (n + seg_len - 1) / seg_len;
    if (segs < 1) segs = 1;
    TkPlan p;
    p.n_segs = (int)segs;
    p.seg_len = seg_len;
    return p;
}

constexpr int TK_E_SCORES_DEFAULT = 16;
// scores per thread and sub-chunk of the streaming selector: 16 or 32 (B2R_TOPK_E overrides: tuning experiments only)
static int tk_e_scores() {
    static const int v = [] {
        const char *e = getenv("B2R_TOPK_E");
        const int x = e ? atoi(e) : 0;
        return x == 32 || x == 16 ? x : TK_E_SCORES_DEFAULT;
    }();
    return v;
}
constexpr int TK_E_KEYS = 8;

static size_t tk_keys_ws(int64_t n_rows, int64_t n, int32_t k) {
    size_t total = 0;
    while (true) {
        TkPlan p = tk_plan(n_rows, n, TK_THREADS * TK_E_KEYS);
        if (p.n_segs == 1) return total;
        total += align_up((size_t)n_rows * p.n_segs * k * 8, 256);
        n = (int64_t)p.n_segs * k;
    }
}

size_t topk_keys_ws_bytes(int64_t n_rows, int64_t n, int32_t k) { return tk_keys_ws(n_rows, n, k) + 256; }

size_t topk_ws_bytes(int64_t n_rows, int64_t n, int32_t k) {
    TkPlan p = tk_plan(n_rows, n, TK_THREADS * tk_e_scores());
    if (p.n_segs == 1) return 256;
    size_t l1 = align_up((size_t)n_rows * p.n_segs * k * 8, 256);
    return l1 + tk_keys_ws(n_rows, (int64_t)p.n_segs * k, k) + 256;
}

int topk_keys_rows(const uint64_t *keys_in, int64_t n_rows, int64_t n, int64_t row_stride, int64_t piece_len,
                   int64_t piece_stride, int32_t k, uint64_t *keys_out, void *ws, size_t ws_bytes,
                   cudaStream_t st, const TopkOpts &opts) {
    if (n_rows == 0) return B2R_OK;
    B2R_CHECK_ARG(k >= 1 && k <= B2R_TOPK_MAX_FAST, "top-k: k=%d outside [1,%d]", k, B2R_TOPK_MAX_FAST);
    char *wp = static_cast<char *>(ws);
    size_t left = ws_bytes;
    while (true) {
        TkPlan p = tk_plan(n_rows, n, TK_THREADS * TK_E_KEYS);
        uint64_t *dst = keys_out;
        if (p.n_segs > 1) {
            size_t need = align_up((size_t)n_rows * p.n_segs * k * 8, 256);
            if (need > left) {
                set_error("top-k merge: workspace too small (%zu > %zu)", need, left);
                return B2R_ERR_WORKSPACE;
            }
            dst = reinterpret_cast<uint64_t *>(wp);
            wp += need;
            left -= need;
        }
        dim3 grid((unsigned)n_rows, (unsigned)p.n_segs);
        topk_stream_kernel<true, TK_E_KEYS><<<grid, TK_THREADS, 0, st>>>(nullptr, keys_in, n, row_stride, piece_len,
                                                                         piece_stride, 0u, k, p.seg_len, dst, 0, 0u,
                                                                         opts.gate, opts.gate_cap);
        B2R_LAUNCH_CHECK();
        if (p.n_segs == 1) return B2R_OK;
        keys_in = dst;
        n = (int64_t)p.n_segs * k;
        row_stride = n;
        piece_len = n;
        piece_stride = 0;
    }
}

int topk_scores_rows(const float *scores, int64_t n_rows, int64_t n, int64_t row_stride, int32_t k,
                     int64_t doc_id_base, uint64_t *keys_out, void *ws, size_t ws_bytes, cudaStream_t st,
                     const TopkOpts &opts) {
    if (n_rows == 0) return B2R_OK;
    B2R_CHECK_ARG(k >= 1 && k <= B2R_TOPK_MAX_FAST, "top-k: k=%d outside [1,%d]", k, B2R_TOPK_MAX_FAST);
    B2R_CHECK_ARG(doc_id_base >= 0 && doc_id_base + n < 0xFFFFFFFFll, "top-k: global doc index exceeds 2^32-2");
    B2R_CHECK_ARG(n_rows < 0x7FFFFFFFll, "top-k: too many rows");
    TkPlan p = tk_plan(n_rows, n, TK_THREADS * tk_e_scores());
    uint64_t *dst = keys_out;
    char *wp = static_cast<char *>(ws);
    size_t left = ws_bytes;
    if (p.n_segs > 1) {
        size_t need = align_up((size_t)n_rows * p.n_segs * k * 8, 256);
        if (need > left) {
            set_error("top-k: workspace too small (%zu > %zu)", need, left);
            return B2R_ERR_WORKSPACE;
        }
        dst = reinterpret_cast<uint64_t *>(wp);
        wp += need;
        left -= need;
    }
    dim3 grid((unsigned)n_rows, (unsigned)p.n_segs);
    if (tk_e_scores() == 32) {
        topk_stream_kernel<false, 32><<<grid, TK_THREADS, 0, st>>>(
            scores, nullptr, n, row_stride, n, 0, (uint32_t)doc_id_base, k, p.seg_len, dst, opts.chunk_shift,
            opts.chunk_stride, opts.gate, opts.gate_cap);
    } else {
        topk_stream_kernel<false, 16><<<grid, TK_THREADS, 0, st>>>(
            scores, nullptr, n, row_stride, n, 0, (uint32_t)doc_id_base, k, p.seg_len, dst, opts.chunk_shift,
            opts.chunk_stride, opts.gate, opts.gate_cap);
    }
    B2R_LAUNCH_CHECK();
    if (p.n_segs == 1) return B2R_OK;
    int64_t n2 = (int64_t)p.n_segs * k;
    TopkOpts o2;
    o2.gate = opts.gate;
    o2.gate_cap = opts.gate_cap;
    return topk_keys_rows(dst, n_rows, n2, n2, n2, 0, k, keys_out, wp, left, st, o2);
}

// ------------------------------------------------------------------------------ top-k of short lists
// Among 256 histogram bins (thread t owns bin 255 - t) find the bin that holds the kk-th largest value: suffix sums
// high to low.  Returns through *bin / *kk_in_bin (shared); if fewer than kk values exist, bin 0 takes it.
// Called by all 256 threads; contains barriers.
__device__ __forceinline__ void select_bin_256(const uint32_t *hist, uint32_t kk, uint32_t *wsum, uint32_t *bin,
                                               uint32_t *kk_in_bin) {
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
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    inc += base;                     // values in bins >= mine
    const uint32_t above = inc - v;  // values in bins > mine
    if ((above < kk && inc >= kk) || (tid == 255 && inc < kk)) {
        *bin = 255u - (uint32_t)tid;
        *kk_in_bin = kk - above;
    }
    __syncthreads();
}

// One CTA per row: the first min(cnt[row], cap) keys of the row's candidate list (the rest of the list is never
// read, so it needs no clearing) are staged in shared memory; a radix select (8 passes of 8 bits) finds the k-th
// largest key, the k keys at or above it are compacted and only those are sorted.  Keys are distinct (they carry
// the document index), so exactly k survive.
constexpr int TL_THREADS = 256;
constexpr int TL_MAX = 16384;  // longest list (dynamic shared memory: 8 bytes per key, 128 KB at the maximum)

__global__ void __launch_bounds__(TL_THREADS)
topk_of_lists_kernel(const uint64_t *__restrict__ lists, int cap, int32_t *__restrict__ cnt, int k, int min_cnt,
                     uint64_t *__restrict__ out, int64_t *__restrict__ idx_out, float *__restrict__ val_out,
                     int32_t *__restrict__ marked) {
    extern __shared__ uint64_t arr[];   // [max(cap, 32)]
    __shared__ uint64_t best[B2R_TOPK_MAX_FAST];
    __shared__ uint32_t hist[256], wsum[8], s_bin, s_kk, s_n;
    const int row = blockIdx.x, tid = threadIdx.x;
    int c = cnt[row];
    __syncthreads();
    // a list shorter than min_cnt cannot supply the k best (its threshold was raised above the sample's k-th
    // best): mark the row for the exhaustive fallback exactly like an overflowed one (cnt > cap)
    if (tid == 0 && c < min_cnt) cnt[row] = cap + 1;
    if (tid == 0 && marked != nullptr && (c < min_cnt || c > cap)) marked[1 + atomicAdd(marked, 1)] = row;
    c = c < 0 ? 0 : (c > cap ? cap : c);
    const uint64_t *src = lists + (int64_t)row * cap;
    for (int i = tid; i < c; i += TL_THREADS) arr[i] = src[i];
    uint64_t *sortbuf = arr;
    int n = c;
    if (c > k) {
        uint64_t prefix = 0, mask = 0;
        uint32_t kk = (uint32_t)k;
        for (int shift = 56; shift >= 0; shift -= 8) {
            hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < c; i += TL_THREADS) {
                const uint64_t key = arr[i];
                if ((key & mask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            select_bin_256(hist, kk, wsum, &s_bin, &s_kk);
            prefix |= (uint64_t)s_bin << shift;
            kk = s_kk;
            mask |= 255ull << shift;
        }
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int i = tid; i < c; i += TL_THREADS) {  // prefix is now the k-th largest key itself
            const uint64_t key = arr[i];
            if (key >= prefix) {
                const uint32_t slot = atomicAdd(&s_n, 1u);
                if (slot < (uint32_t)k) best[slot] = key;
            }
        }
        __syncthreads();
        sortbuf = best;
        n = min((int)s_n, k);
    }
    int P = 32;
    while (P < n || P < k) P <<= 1;  // P <= 1024 on the compacted path, <= TL_MAX otherwise (c <= k <= cap)
    __syncthreads();
    for (int i = n + tid; i < P; i += TL_THREADS) sortbuf[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc<TL_THREADS>(sortbuf, P);
    for (int i = tid; i < k; i += TL_THREADS) {   // ranked keys, and their decoded form when asked for
        const uint64_t key = sortbuf[i];
        const int64_t at = (int64_t)row * k + i;
        if (out) out[at] = key;
        if (idx_out) idx_out[at] = key ? (int64_t)(0xFFFFFFFFu - (uint32_t)key) : -1;
        if (val_out) val_out[at] = key ? unord_f32((uint32_t)(key >> 32)) : __int_as_float(0xff800000);
    }
}

int topk_of_lists(const uint64_t *lists, int64_t n_rows, int cap, int32_t *cnt, int32_t k, int32_t min_cnt,
                  uint64_t *keys_out, cudaStream_t st, int64_t *idx_out, float *val_out, int32_t *marked) {
    if (n_rows == 0) return B2R_OK;
    B2R_CHECK_ARG(cap >= 1 && cap <= TL_MAX && k >= 1 && k <= cap && k <= B2R_TOPK_MAX_FAST,
                  "top-k of lists: cap=%d / k=%d unsupported", cap, k);
    // the uncompacted path (c <= k) sorts in place: room for the power of two above k (>= 32) as well
    int pk = 32;
    while (pk < k) pk <<= 1;
    const size_t smem = (size_t)(cap > pk ? cap : pk) * 8;
    if (smem > 40 * 1024)
        B2R_CUDA(cudaFuncSetAttribute(topk_of_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_of_lists_kernel<<<(unsigned)n_rows, TL_THREADS, smem, st>>>(lists, cap, cnt, k, min_cnt, keys_out, idx_out, val_out,
                                                                     marked);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

// ------------------------------------------------------------------------------ k-th of group maxima
// One CTA per row: radix select (4 passes of 8 bits over the ordered encoding) of the k-th largest of the
// row's group maxima.  The rows are short (hundreds to a few ten thousand values) and stay in L1/L2.
__global__ void __launch_bounds__(256)
kth_of_maxima_kernel(const float *__restrict__ maxima, int64_t n_groups, int64_t row_stride, int k, int lower,
                     int positive_floor, uint64_t *__restrict__ thr_out, int32_t *__restrict__ zero_a,
                     int32_t *__restrict__ zero_b, int32_t *__restrict__ zero_scalar) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t s_prefix, s_k, s_bin;
    const float *row = maxima + (int64_t)blockIdx.x * row_stride;
    const int tid = threadIdx.x;
    uint32_t prefix = 0, mask = 0, kk = (uint32_t)k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n_groups; i += 256) {
            const uint32_t o = ord_f32(__ldg(row + i));
            if ((o & mask) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
        }
        __syncthreads();
        select_bin_256(hist, kk, wsum, &s_bin, &s_k);
        if (tid == 0) s_prefix = prefix | (s_bin << shift);
        __syncthreads();
        prefix = s_prefix;
        kk = s_k;
        mask |= 255u << shift;
        __syncthreads();
    }
    if (tid == 0) {
        // prefix = ordered encoding of the k-th largest (if fewer than k values exist the scan above ends in
        // bin 0 of every pass: prefix <= the encoding of -inf, which also means "no threshold")
        uint32_t o = prefix;
        if (o <= 0x007fffffu) {
            o = 0;
        } else if (lower) {
            const float t = unord_f32(o);
            const float tl = __fsub_rn(__fsub_rn(t, __fmul_rn(fabsf(t), 9.5367431640625e-07f)), 1e-37f);
            o = ord_f32(tl);
            if (o <= 0x007fffffu) o = 0;
        }
        uint64_t thr = (uint64_t)o << 32;
        // A sample whose k-th best is <= 0 (a query that matches few documents: untouched documents score exactly
        // 0) would admit every document.  Keep only strictly positive scores instead; a query with fewer than k
        // of them is caught by the short-list gate of topk_of_lists and rescored exhaustively.
        if (positive_floor && o <= 0x80000000u) thr = (0x80000000ull << 32) | 0xFFFFFFFFull;
        thr_out[blockIdx.x] = thr;
        // the row's candidate counters start at zero (saves the caller a memset node per step)
        if (zero_a) zero_a[blockIdx.x] = 0;
        if (zero_b) zero_b[blockIdx.x] = 0;
        if (zero_scalar && blockIdx.x == 0) *zero_scalar = 0;
    }
}

int kth_of_maxima(const float *maxima, int64_t n_rows, int64_t n_groups, int64_t row_stride, int32_t k, bool lower,
                  bool positive_floor, uint64_t *thr_out, cudaStream_t st, int32_t *zero_a, int32_t *zero_b,
                  int32_t *zero_scalar) {
    if (n_rows == 0) return B2R_OK;
    kth_of_maxima_kernel<<<(unsigned)n_rows, 256, 0, st>>>(maxima, n_groups, row_stride, k, lower ? 1 : 0,
                                                           positive_floor ? 1 : 0, thr_out, zero_a, zero_b, zero_scalar);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

// ------------------------------------------------------------------------------ threshold-filter top-k of score rows
// The fast path of b2r_topk (fast_topk_selection) for long rows and k <= 128, the scheme of the fused search path
// applied to a dense score matrix: (1) group maxima of every step-th 4096-score chunk, (2) the k-th largest group
// maximum T per row (k groups, hence k scores, reach it), (3) one streaming pass that appends the scores >= T to a
// candidate list (a quarter LDG.128 and one FMNMX per score, no barrier), (4) top-k of the lists.  Rows whose list
// overflows (or comes up short: NaNs) are redone by the gated streaming selector.  Reads (1 + 1/step) x 4 B per score.
constexpr int TF_CHUNK = 4096;  // scores per chunk = 256 threads x 4 x float4

struct TfPlan {
    bool on;
    int step, cap;
    int64_t n_chunks, n_sample, n_groups;
};
static TfPlan tf_plan(int64_t n, int32_t k) {
    TfPlan p = {};
    p.n_chunks = (n + TF_CHUNK - 1) / TF_CHUNK;
    p.on = k >= 1 && k <= 128 && p.n_chunks >= 32;
    if (!p.on) return p;
    p.step = k <= 16 ? 64 : 16;
    p.n_sample = (p.n_chunks + p.step - 1) / p.step;
    p.n_groups = p.n_sample * 256;
    const double r = (double)p.n_chunks / p.n_sample;
    const double want = k * r + 6.0 * sqrt((double)k) * r + k;
    p.cap = 256;
    while (p.cap < want && p.cap < TL_MAX) p.cap <<= 1;
    return p;
}
static size_t tf_bytes(const TfPlan &p, int64_t n_rows) {
    if (!p.on) return 0;
    return align_up((size_t)n_rows * p.n_groups * 4, 256) + align_up((size_t)n_rows * 8, 256) +
           align_up((size_t)n_rows * p.cap * 8, 256) + align_up((size_t)n_rows * 4, 256) + 256;
}

// thread t of CTA (sample chunk j, row): maximum of its 16 scores of chunk j * step
__global__ void __launch_bounds__(256)
row_maxima_kernel(const float *__restrict__ scores, int64_t n, int64_t row_stride, int step, int64_t n_groups,
                  float *__restrict__ maxima) {
    const float *rp = scores + (int64_t)blockIdx.y * row_stride;
    const int64_t base = (int64_t)blockIdx.x * step * TF_CHUNK;
    float m = __int_as_float(0xff800000);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int64_t j = base + ((int64_t)v * 256 + threadIdx.x) * 4;
        if (j + 3 < n) {
            const float4 x = ldg_stream_f4(rp + j);
            m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
        } else {
            for (int64_t e = j; e < n; ++e) m = fmaxf(m, __ldg(rp + e));
        }
    }
    maxima[(int64_t)blockIdx.y * n_groups + (int64_t)blockIdx.x * 256 + threadIdx.x] = m;
}

// CTA (segment, row) walks chunks segment, segment + gridDim.x, ...; scores >= the row's threshold are keyed and appended
__global__ void __launch_bounds__(256)
row_filter_kernel(const float *__restrict__ scores, int64_t n, int64_t row_stride, int64_t n_chunks, uint32_t id_base,
                  const uint64_t *__restrict__ thr_keys, uint64_t *__restrict__ cand, int32_t *__restrict__ cnt, int cap) {
    const int row = blockIdx.y;
    const float *rp = scores + (int64_t)row * row_stride;
    const uint64_t thr = thr_keys[row];
    const uint32_t thr_hi = (uint32_t)(thr >> 32);
    const float thr_f = thr_hi ? unord_f32(thr_hi) : __int_as_float(0xff800000);
    uint64_t *my = cand + (int64_t)row * cap;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t base = c * TF_CHUNK + (int64_t)threadIdx.x * 4;
        float4 x[4];
        float m = __int_as_float(0xff800000);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int64_t j = base + (int64_t)v * 1024;
            if (j + 3 < n) {
                x[v] = ldg_stream_f4(rp + j);
            } else {
                const float ninf = __int_as_float(0xff800000);
                x[v] = make_float4(j < n ? __ldg(rp + j) : ninf, j + 1 < n ? __ldg(rp + j + 1) : ninf,
                                   j + 2 < n ? __ldg(rp + j + 2) : ninf, ninf);
            }
            m = fmaxf(m, fmaxf(fmaxf(x[v].x, x[v].y), fmaxf(x[v].z, x[v].w)));
        }
        if (m >= thr_f) {  // rare
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float f[4] = {x[v].x, x[v].y, x[v].z, x[v].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t j = base + (int64_t)v * 1024 + e;
                    if (f[e] >= thr_f && j < n) {
                        const uint64_t key = make_key(ord_f32(f[e]), id_base + (uint32_t)j);
                        if (key > thr) {
                            const int slot = atomicAdd(cnt + row, 1);
                            if (slot < cap) my[slot] = key;
                        }
                    }
                }
            }
        }
    }
}

// keys_out[row, 0..k): exact top-k of every row through the threshold filter; rows it cannot finish fall back to
// topk_scores_rows (gated).  Needs 16-byte aligned rows (checked by the caller).
static int topk_scores_rows_filtered(const float *scores, int64_t n_rows, int64_t n, int64_t row_stride, int32_t k,
                                     int64_t doc_id_base, uint64_t *keys_out, const TfPlan &p, void *ws,
                                     size_t ws_bytes, cudaStream_t st) {
    char *wp = static_cast<char *>(ws);
    if (tf_bytes(p, n_rows) > ws_bytes) {
        set_error("top-k: workspace too small (%zu > %zu)", tf_bytes(p, n_rows), ws_bytes);
        return B2R_ERR_WORKSPACE;
    }
    auto carve = [&](size_t bytes) -> void * {
        void *q = wp;
        wp += align_up(bytes, 256);
        return q;
    };
    float *maxima = static_cast<float *>(carve((size_t)n_rows * p.n_groups * 4));
    uint64_t *thr = static_cast<uint64_t *>(carve((size_t)n_rows * 8));
    uint64_t *cand = static_cast<uint64_t *>(carve((size_t)n_rows * p.cap * 8));
    int32_t *cnt = static_cast<int32_t *>(carve((size_t)n_rows * 4));
    const size_t left = ws_bytes - (size_t)(wp - static_cast<char *>(ws));
    row_maxima_kernel<<<dim3((unsigned)p.n_sample, (unsigned)n_rows), 256, 0, st>>>(scores, n, row_stride, p.step,
                                                                                  p.n_groups, maxima);
    B2R_LAUNCH_CHECK();
    int rc = kth_of_maxima(maxima, n_rows, p.n_groups, p.n_groups, k, false, false, thr, st);
    if (rc) return rc;
    B2R_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n_rows * 4, st));
    int64_t segs = (148 * 16 + n_rows - 1) / n_rows;
    if (segs < 1) segs = 1;
    if (segs > p.n_chunks) segs = p.n_chunks;
    row_filter_kernel<<<dim3((unsigned)segs, (unsigned)n_rows), 256, 0, st>>>(scores, n, row_stride, p.n_chunks,
                                                                            (uint32_t)doc_id_base, thr, cand, cnt, p.cap);
    B2R_LAUNCH_CHECK();
    rc = topk_of_lists(cand, n_rows, p.cap, cnt, k, k, keys_out, st);
    if (rc) return rc;
    TopkOpts gate;
    gate.gate = cnt;
    gate.gate_cap = p.cap;
    return topk_scores_rows(scores, n_rows, n, row_stride, k, doc_id_base, keys_out, wp, left, st, gate);
}

// ------------------------------------------------------------------------------------------ decode
__global__ void decode_keys_kernel(const uint64_t *__restrict__ keys, int64_t n, int64_t *__restrict__ idx_out,
                                   float *__restrict__ val_out, const float *__restrict__ scores,
                                   int64_t row_stride, int k, int64_t doc_id_base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    int64_t idx = -1;
    float val = __int_as_float(0xff800000);  // -inf
    if (key != 0) {
        uint32_t gid = 0xFFFFFFFFu - (uint32_t)key;
        idx = (int64_t)gid;
        val = scores ? scores[(i / k) * row_stride + (idx - doc_id_base)] : unord_f32((uint32_t)(key >> 32));
    }
    if (idx_out) idx_out[i] = idx;
    if (val_out) val_out[i] = val;
}

int decode_keys(const uint64_t *keys, int64_t n, int64_t *idx_out, float *val_out, const float *scores,
                int64_t row_stride, int32_t k, int64_t doc_id_base, cudaStream_t st) {
    if (n == 0 || (!idx_out && !val_out)) return B2R_OK;
    int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    decode_keys_kernel<<<(unsigned)blocks, threads, 0, st>>>(keys, n, idx_out, val_out, scores, row_stride,
                                                            k > 0 ? k : 1, doc_id_base);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

// ------------------------------------------------------------------------------------------ full sort (k > 1024)
__global__ void make_keys_kernel(const float *__restrict__ scores, int64_t n, int64_t P, uint32_t id_base,
                                 uint64_t *__restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    keys[i] = i < n ? make_key(ord_f32(scores[i]), id_base + (uint32_t)i) : 0ull;
}

__global__ void bitonic_global_step_kernel(uint64_t *__restrict__ keys, int64_t P, int64_t size, int64_t stride) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (P >> 1)) return;
    int64_t pos = 2 * i - (i & (stride - 1));
    uint64_t a = keys[pos], b = keys[pos + stride];
    bool desc = ((pos & size) == 0);
    if (desc ? (a < b) : (a > b)) {
        keys[pos] = b;
        keys[pos + stride] = a;
    }
}

static int64_t pow2_ceil(int64_t x) {
    int64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

static int full_sort_row(const float *scores, int64_t n, int64_t doc_id_base, uint64_t *buf, cudaStream_t st) {
    int64_t P = pow2_ceil(n < 2 ? 2 : n);
    int threads = 256;
    make_keys_kernel<<<(unsigned)((P + threads - 1) / threads), threads, 0, st>>>(scores, n, P, (uint32_t)doc_id_base,
                                                                                 buf);
    B2R_LAUNCH_CHECK();
    unsigned blocks = (unsigned)(((P >> 1) + threads - 1) / threads);
    for (int64_t size = 2; size <= P; size <<= 1)
        for (int64_t stride = size >> 1; stride > 0; stride >>= 1)
            bitonic_global_step_kernel<<<blocks, threads, 0, st>>>(buf, P, size, stride);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

}  // namespace b2r

using namespace b2r;

// B2R_TOPK_FILTER=0 keeps b2r_topk on the streaming selector (A/B measurements only)
static const bool g_topk_filter = [] {
    const char *e = getenv("B2R_TOPK_FILTER");
    return !(e && atoi(e) == 0);
}();

extern "C" int b2r_topk_workspace(int64_t n_rows, int64_t n, int32_t k, size_t *bytes) {
    B2R_CHECK_ARG(bytes && n_rows >= 0 && n >= 0 && k >= 1, "b2r_topk_workspace: bad arguments");
    if (k <= B2R_TOPK_MAX_FAST) {
        *bytes = topk_ws_bytes(n_rows > 0 ? n_rows : 1, n > 0 ? n : 1, k) + (size_t)n_rows * k * 8 + 256 +
                 tf_bytes(tf_plan(n > 0 ? n : 1, k), n_rows > 0 ? n_rows : 1);
    } else {
        *bytes = (size_t)pow2_ceil(n < 2 ? 2 : n) * 8 + 256;
    }
    return B2R_OK;
}

extern "C" int b2r_topk(const float *scores, int64_t n_rows, int64_t n, int64_t row_stride, int32_t k,
                        int64_t doc_id_base, uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace,
                        size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(scores && n_rows >= 0 && n >= 1 && row_stride >= n, "b2r_topk: bad shape");
    B2R_CHECK_ARG(k >= 1 && k <= n, "b2r_topk: need 1 <= k <= n (k=%d, n=%lld)", k, (long long)n);
    B2R_CHECK_ARG(doc_id_base >= 0 && doc_id_base + n < 0xFFFFFFFFll, "b2r_topk: global doc index exceeds 2^32-2");
    if (n_rows == 0) return B2R_OK;
    char *wp = static_cast<char *>(workspace);
    size_t left = workspace_bytes;
    if (k <= B2R_TOPK_MAX_FAST) {
        uint64_t *keys = keys_out;
        if (!keys) {
            size_t need = align_up((size_t)n_rows * k * 8, 256);
            if (need > left) {
                set_error("b2r_topk: workspace too small");
                return B2R_ERR_WORKSPACE;
            }
            keys = reinterpret_cast<uint64_t *>(wp);
            wp += need;
            left -= need;
        }
        const TfPlan fp = tf_plan(n, k);
        const bool aligned = (reinterpret_cast<uintptr_t>(scores) & 15) == 0 && (n_rows == 1 || (row_stride & 3) == 0);
        int rc = (fp.on && aligned && g_topk_filter)
                     ? topk_scores_rows_filtered(scores, n_rows, n, row_stride, k, doc_id_base, keys, fp, wp, left, st)
                     : topk_scores_rows(scores, n_rows, n, row_stride, k, doc_id_base, keys, wp, left, st);
        if (rc) return rc;
        return decode_keys(keys, n_rows * k, idx_out, val_out, scores, row_stride, k, doc_id_base, st);
    }
    // large k: sort the whole row (one row at a time; rare path: top_k > 1024)
    size_t need = (size_t)pow2_ceil(n < 2 ? 2 : n) * 8;
    if (need > left) {
        set_error("b2r_topk: workspace too small for the full-sort path (%zu > %zu)", need, left);
        return B2R_ERR_WORKSPACE;
    }
    uint64_t *buf = reinterpret_cast<uint64_t *>(wp);
    for (int64_t r = 0; r < n_rows; ++r) {
        int rc = full_sort_row(scores + r * row_stride, n, doc_id_base, buf, st);
        if (rc) return rc;
        if (keys_out)
            B2R_CUDA(cudaMemcpyAsync(keys_out + r * k, buf, (size_t)k * 8, cudaMemcpyDeviceToDevice, st));
        rc = decode_keys(buf, k, idx_out ? idx_out + r * k : nullptr, val_out ? val_out + r * k : nullptr,
                         scores + r * row_stride, row_stride, k, doc_id_base, st);
        if (rc) return rc;
    }
    return B2R_OK;
}

extern "C" int b2r_merge_candidates(const uint64_t *gathered, int32_t n_parts, int32_t n_queries, int32_t k,
                                    uint64_t *keys_out, int64_t *idx_out, float *val_out, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(gathered && n_parts >= 1 && n_queries >= 0 && k >= 1 && k <= B2R_TOPK_MAX_FAST,
                  "b2r_merge_candidates: bad arguments");
    if (n_queries == 0) return B2R_OK;
    char *wp = static_cast<char *>(workspace);
    size_t left = workspace_bytes;
    uint64_t *keys = keys_out;
    if (!keys) {
        size_t need = align_up((size_t)n_queries * k * 8, 256);
        if (need > left) {
            set_error("b2r_merge_candidates: workspace too small");
            return B2R_ERR_WORKSPACE;
        }
        keys = reinterpret_cast<uint64_t *>(wp);
        wp += need;
        left -= need;
    }
    int rc = topk_keys_rows(gathered, n_queries, (int64_t)n_parts * k, k, k, (int64_t)n_queries * k, k, keys, wp,
                            left, st);
    if (rc) return rc;
    return decode_keys(keys, (int64_t)n_queries * k, idx_out, val_out, nullptr, 0, k, 0, st);
}

extern "C" int b2r_decode_keys(const uint64_t *keys, int64_t n, int64_t *idx_out, float *val_out, void *stream) {
    B2R_CHECK_ARG(keys || n == 0, "b2r_decode_keys: null keys");
    return decode_keys(keys, n, idx_out, val_out, nullptr, 0, 1, 0, static_cast<cudaStream_t>(stream));
}
