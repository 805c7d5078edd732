// Host-side helpers of the on-disk index format (include/b200ret.h, "On-disk form of a b2r_index").
// No device code: the file holds the device buffers verbatim, the host language (Python here) moves the bytes.
#include <string.h>

#include "common.cuh"

namespace b2r {

static int section_sizes(const b2r_index_file_header *h, uint64_t out[B2R_SEC_COUNT]) {
    B2R_CHECK_ARG(h->n_docs >= 1 && h->nnz >= 0 && h->n_vocab >= 1, "index file: bad dimensions");
    b2r_index_sizes sz;
    int rc = b2r_index_sizes_for(h->nnz, h->n_docs, h->n_vocab, h->tile_docs, h->kind, &sz);
    if (rc) return rc;
    const int64_t n_tiles = (h->n_docs + h->tile_docs - 1) / h->tile_docs;
    B2R_CHECK_ARG(h->n_dense_max >= 1, "index file: bad n_dense_max");
    out[B2R_SEC_POST_DOC] = sz.post_doc_bytes;
    out[B2R_SEC_POST_VAL] = sz.post_val_bytes;
    out[B2R_SEC_BLK_PTR] = sz.blk_ptr_bytes;
    out[B2R_SEC_DENSE_ID] = sz.dense_id_bytes;
    // the dense table is sized by the n_dense_max the index was BUILT with (stored in the header)
    out[B2R_SEC_DENSE_PTR] = align_up((size_t)h->n_dense_max * ((size_t)n_tiles * B2R_SUBTILES + 1) * 4, 256);
    out[B2R_SEC_IDF] = (uint64_t)h->n_vocab * 4;
    return B2R_OK;
}

}  // namespace b2r

using namespace b2r;

// Two interleaved multiply-xorshift lanes over 64-bit words (tail bytes zero-padded), folded with the length.
extern "C" uint64_t b2r_checksum64(const void *data, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    uint64_t a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full;
    size_t i = 0;
    for (; i + 16 <= bytes; i += 16) {
        uint64_t x, y;
        memcpy(&x, p + i, 8);
        memcpy(&y, p + i + 8, 8);
        a = (a ^ x) * 0xFF51AFD7ED558CCDull;
        a ^= a >> 29;
        b = (b ^ y) * 0xC4CEB9FE1A85EC53ull;
        b ^= b >> 31;
    }
    if (i < bytes) {
        unsigned char tail[16] = {0};
        memcpy(tail, p + i, bytes - i);
        uint64_t x, y;
        memcpy(&x, tail, 8);
        memcpy(&y, tail + 8, 8);
        a = (a ^ x) * 0xFF51AFD7ED558CCDull;
        a ^= a >> 29;
        b = (b ^ y) * 0xC4CEB9FE1A85EC53ull;
        b ^= b >> 31;
    }
    uint64_t h = a ^ (b * 0x9E3779B97F4A7C15ull) ^ (uint64_t)bytes;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    return h;
}

extern "C" int b2r_index_file_layout(b2r_index_file_header *hdr, uint64_t *file_bytes) {
    B2R_CHECK_ARG(hdr, "b2r_index_file_layout: null header");
    static_assert(sizeof(b2r_index_file_header) <= B2R_FILE_ALIGN, "header must fit its page");
    uint64_t sizes[B2R_SEC_COUNT];
    int rc = section_sizes(hdr, sizes);
    if (rc) return rc;
    memcpy(hdr->magic, B2R_FILE_MAGIC, 8);
    hdr->version = B2R_FILE_VERSION;
    hdr->header_bytes = B2R_FILE_ALIGN;
    hdr->n_tiles = (int32_t)((hdr->n_docs + hdr->tile_docs - 1) / hdr->tile_docs);
    hdr->subtiles = B2R_SUBTILES;
    uint64_t off = B2R_FILE_ALIGN;
    for (int s = 0; s < B2R_SEC_COUNT; ++s) {
        hdr->sections[s].offset = off;
        hdr->sections[s].bytes = sizes[s];
        hdr->sections[s].checksum = 0;
        off = align_up((size_t)(off + sizes[s]), B2R_FILE_ALIGN);
    }
    if (file_bytes) *file_bytes = off;
    return B2R_OK;
}

#define B2R_CHECK_DATA(cond, ...)        \
    do {                                 \
        if (!(cond)) {                   \
            b2r::set_error(__VA_ARGS__); \
            return B2R_ERR_DATA;         \
        }                                \
    } while (0)

extern "C" int b2r_index_file_check(const b2r_index_file_header *hdr, uint64_t file_bytes) {
    B2R_CHECK_ARG(hdr, "b2r_index_file_check: null header");
    B2R_CHECK_DATA(file_bytes >= B2R_FILE_ALIGN, "index file: %llu bytes is shorter than the header page",
                   (unsigned long long)file_bytes);
    B2R_CHECK_DATA(memcmp(hdr->magic, B2R_FILE_MAGIC, 8) == 0, "index file: bad magic (not a b200ret index)");
    B2R_CHECK_DATA(hdr->version == B2R_FILE_VERSION, "index file: version %u, this library reads version %u",
                   hdr->version, B2R_FILE_VERSION);
    B2R_CHECK_DATA(hdr->header_bytes == B2R_FILE_ALIGN && hdr->subtiles == B2R_SUBTILES,
                   "index file: header_bytes/subtiles (%u, %d) do not match this library (%u, %d)", hdr->header_bytes,
                   hdr->subtiles, B2R_FILE_ALIGN, B2R_SUBTILES);
    uint64_t sizes[B2R_SEC_COUNT];
    if (section_sizes(hdr, sizes) != B2R_OK) return B2R_ERR_DATA;  // message set by section_sizes
    B2R_CHECK_DATA(hdr->n_tiles == (hdr->n_docs + hdr->tile_docs - 1) / hdr->tile_docs, "index file: n_tiles mismatch");
    B2R_CHECK_DATA(hdr->doc_id_base >= 0 && hdr->doc_id_base + hdr->n_docs < 0xFFFFFFFFll,
                   "index file: global doc index exceeds 2^32-2");
    uint64_t prev_end = B2R_FILE_ALIGN;
    for (int s = 0; s < B2R_SEC_COUNT; ++s) {
        const b2r_file_section &sec = hdr->sections[s];
        B2R_CHECK_DATA(sec.bytes == sizes[s], "index file: section %d holds %llu bytes, the dimensions need %llu", s,
                       (unsigned long long)sec.bytes, (unsigned long long)sizes[s]);
        B2R_CHECK_DATA(sec.offset % B2R_FILE_ALIGN == 0 && sec.offset >= prev_end && sec.offset + sec.bytes <= file_bytes,
                       "index file: section %d [%llu, +%llu) is misaligned, overlaps or exceeds the file (%llu bytes)", s,
                       (unsigned long long)sec.offset, (unsigned long long)sec.bytes, (unsigned long long)file_bytes);
        prev_end = sec.offset + sec.bytes;
    }
    return B2R_OK;
}
