// Candidate exchange + merge for a doc-sharded corpus in ONE kernel over peer memory (NVLink / NVSwitch).
// No reference counterpart (the reference is a single CPU process); SURVEY.md section 8e / section 5.
//
// Every rank holds a RECEIVE buffer that all ranks of the box can write (PyTorch symmetric memory: CUDA VMM
// allocations mapped into every process).  Per batch each rank has k ranked candidate keys per query
// (b2r_search_batch).  A CTA of b2r_exchange_merge owns a chunk of queries and
//   1. PUSHES its rank's keys of the chunk into slot `rank` of every rank's receive buffer (plain 8-byte stores that
//      travel over NVLink), fences, and raises flag[rank][chunk] on every rank to this call's epoch;
//   2. WAITS until all ranks' flags of the chunk in its OWN buffer carry the epoch;
//   3. MERGES the n_parts ranked lists of each query of the chunk: a key's final rank is the number of keys (of all
//      lists) above it, found by one binary search per list -- no sort, every key in parallel.
// The exchange is 80 KB per rank at k = 10: what it costs is latency, and this replaces an NCCL all-gather launch, its
// protocol round trip and a separate merge launch by one launch whose peers synchronise chunk by chunk.
//
// Safety of buffer reuse without a barrier: data and flags are double-buffered on the parity of the epoch.  A rank can
// only finish call n + 1 after every rank pushed n + 1, which every rank does after finishing ITS call n (stream
// order), so nobody writes parity p again before every reader of the previous use of parity p is done.
// The wait is bounded (about 2 s): on expiry the kernel records an error in the buffer header instead of hanging.
#include "common.cuh"

namespace b2r {

constexpr int XC_THREADS = 256;
constexpr int XC_MAX_CHUNKS = 128;   // CTAs per call (all resident at once: they wait for each other's peers)
constexpr int XC_MAX_PARTS = 16;
constexpr size_t XC_HDR_BYTES = 256;

struct XcHeader {
    uint32_t epoch;    // calls completed on this buffer
    uint32_t ticket;   // CTAs of the running call that are done
    uint32_t error;    // sticky: 1 = a wait timed out
    uint32_t pad;
};

__host__ __device__ inline size_t xc_flags_bytes(int n_parts) {
    return (size_t)2 * n_parts * XC_MAX_CHUNKS * sizeof(uint32_t);
}
__host__ __device__ inline size_t xc_data_offset(int n_parts) {
    return XC_HDR_BYTES + ((xc_flags_bytes(n_parts) + 255) / 256) * 256;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(XC_THREADS)
exchange_merge_kernel(const uint64_t *__restrict__ local_keys, void *const *__restrict__ peer_bufs, int rank, int n_parts,
                      int n_queries, int k, int q_per_chunk, uint64_t *__restrict__ keys_out, int64_t *__restrict__ idx_out,
                      float *__restrict__ val_out) {
    extern __shared__ uint64_t xs[];   // [q_per_chunk][n_parts][k] the chunk's lists, part-major per query
    __shared__ uint32_t s_epoch;
    const int tid = threadIdx.x, c = blockIdx.x;
    char *mine = static_cast<char *>(peer_bufs[rank]);
    XcHeader *hdr = reinterpret_cast<XcHeader *>(mine);
    if (tid == 0) s_epoch = *reinterpret_cast<volatile uint32_t *>(&hdr->epoch) + 1u;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const int par = (int)(epoch & 1u);
    const int q0 = c * q_per_chunk, q1 = min(n_queries, q0 + q_per_chunk);
    const int n_keys = (q1 - q0) * k;
    const size_t data_off = xc_data_offset(n_parts);
    const size_t part_stride = (size_t)n_queries * k;              // keys of one rank
    const size_t par_stride = (size_t)n_parts * part_stride;       // keys of one parity
    // ---- 1. push my keys of the chunk to every rank (own buffer included)
    for (int p = 0; p < n_parts; ++p) {
        uint64_t *dst = reinterpret_cast<uint64_t *>(static_cast<char *>(peer_bufs[p]) + data_off) + par * par_stride +
                        (size_t)rank * part_stride + (size_t)q0 * k;
        for (int i = tid; i < n_keys; i += XC_THREADS) dst[i] = local_keys[(size_t)q0 * k + i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < n_parts) {
        uint32_t *flag = reinterpret_cast<uint32_t *>(static_cast<char *>(peer_bufs[tid]) + XC_HDR_BYTES) +
                         ((size_t)par * n_parts + rank) * XC_MAX_CHUNKS + c;
        st_release_sys(flag, epoch);
    }
    // ---- 2. wait for every rank's keys of this chunk
    if (tid < n_parts) {
        const uint32_t *flag = reinterpret_cast<const uint32_t *>(mine + XC_HDR_BYTES) +
                               ((size_t)par * n_parts + tid) * XC_MAX_CHUNKS + c;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != epoch) {
            if (clock64() - t0 > (1ll << 32)) {   // ~2 s at 2 GHz: give up loudly instead of hanging the GPU
                atomicExch(&hdr->error, 1u);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    // ---- 3. merge: stage the chunk's lists, then every key finds its rank
    const uint64_t *src = reinterpret_cast<const uint64_t *>(mine + data_off) + par * par_stride;
    const int per_q = n_parts * k;
    for (int i = tid; i < (q1 - q0) * per_q; i += XC_THREADS) {
        const int ql = i / per_q, r = i % per_q, p = r / k, j = r % k;
        xs[i] = ld_volatile_u64(src + (size_t)p * part_stride + (size_t)(q0 + ql) * k + j);
    }
    for (int i = tid; i < n_keys; i += XC_THREADS) {   // "no candidate" everywhere first
        const size_t at = (size_t)q0 * k + i;
        if (keys_out) keys_out[at] = 0ull;
        if (idx_out) idx_out[at] = -1;
        if (val_out) val_out[at] = __int_as_float(0xff800000);
    }
    __syncthreads();
    for (int i = tid; i < (q1 - q0) * per_q; i += XC_THREADS) {
        const uint64_t key = xs[i];
        if (key == 0ull) continue;   // padding of a short list
        const int ql = i / per_q;
        const uint64_t *lists = xs + (size_t)ql * per_q;
        int above = 0;   // keys of all lists that rank before this one (keys are distinct: they carry the document)
        for (int p = 0; p < n_parts; ++p) {
            const uint64_t *l = lists + p * k;   // descending
            int lo = 0, hi = k;                  // first position whose key is <= mine
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (l[mid] > key) lo = mid + 1;
                else hi = mid;
            }
            above += lo;
        }
        if (above < k) {
            const size_t at = (size_t)(q0 + ql) * k + above;
            if (keys_out) keys_out[at] = key;
            if (idx_out) idx_out[at] = (int64_t)(0xFFFFFFFFu - (uint32_t)key);
            if (val_out) val_out[at] = unord_f32((uint32_t)(key >> 32));
        }
    }
    // ---- the last CTA of the call closes the epoch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&hdr->ticket, 1u) == gridDim.x - 1) {
            hdr->ticket = 0;
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(&hdr->epoch) = epoch;
        }
    }
}

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_exchange_bytes(int32_t n_parts, int32_t n_queries, int32_t k) {
    if (n_parts < 1 || n_queries < 0 || k < 1) return 0;
    return xc_data_offset(n_parts) + (size_t)2 * n_parts * (size_t)n_queries * k * 8;
}

extern "C" int b2r_exchange_merge(const uint64_t *local_keys, void *const *peer_bufs_dev, int32_t rank, int32_t n_parts,
                                  int32_t n_queries, int32_t k, uint64_t *keys_out, int64_t *idx_out, float *val_out,
                                  void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(local_keys && peer_bufs_dev && n_parts >= 1 && n_parts <= XC_MAX_PARTS && rank >= 0 && rank < n_parts,
                  "b2r_exchange_merge: bad rank / n_parts (<= %d)", XC_MAX_PARTS);
    B2R_CHECK_ARG(n_queries >= 0 && k >= 1 && k <= B2R_TOPK_MAX_FAST, "b2r_exchange_merge: bad n_queries / k");
    if (n_queries == 0) return B2R_OK;
    const int q_per_chunk = (n_queries + XC_MAX_CHUNKS - 1) / XC_MAX_CHUNKS;
    // the staged lists of a chunk live in the shared memory of its CTA
    const size_t smem = (size_t)n_parts * k * 8 * q_per_chunk;
    if (smem > 200 * 1024) {
        set_error("b2r_exchange_merge: %d queries x %d parts x k=%d do not fit the %d chunks of a call", n_queries,
                  n_parts, k, XC_MAX_CHUNKS);
        return B2R_ERR_UNSUPPORTED;
    }
    const int n_chunks = (n_queries + q_per_chunk - 1) / q_per_chunk;
    if (smem > 48 * 1024)
        B2R_CUDA(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    exchange_merge_kernel<<<n_chunks, XC_THREADS, smem, st>>>(local_keys, peer_bufs_dev, rank, n_parts, n_queries, k,
                                                             q_per_chunk, keys_out, idx_out, val_out);
    B2R_LAUNCH_CHECK();
    return B2R_OK;
}

extern "C" int b2r_exchange_status(const void *recv_buf, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2R_CHECK_ARG(recv_buf, "b2r_exchange_status: null buffer");
    XcHeader h;
    B2R_CUDA(cudaMemcpyAsync(&h, recv_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    B2R_CUDA(cudaStreamSynchronize(st));
    if (h.error) {
        set_error("b2r_exchange_merge: a rank's candidates never arrived (wait timed out)");
        return B2R_ERR_CUDA;
    }
    return B2R_OK;
}
