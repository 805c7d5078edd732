"""Drop-in functions with the reference's kernel signatures, running on the B200.

  simd_bm25_score / simd_bm25_batch_score  <- rag_system/core/retrieval.py:41-76,
                                              rag_system/core/retriever_registry.py:37-72
  fast_topk_selection                      <- retrieval.py:79-92 (int64 indices),
                                              evaluate_rag_pipeline.py:124-159 (int32: index_dtype=)
  simd_tfidf_score                         <- rag_system/pipeline/evaluate_rag_pipeline.py:95-121
  quantized_dot_product_batch              <- rag_system/core/retriever_registry.py:90-117
  optimized_bm25_score / fast_topk         <- README.md:128-203 aliases (semantics of the executable
                                              code above; k1/b/avgdl as keyword arguments)

Inputs and outputs are numpy arrays like the reference's; CUDA tensors are accepted too and then
returned.  The array-level scorers need a term-major index: it is built on the GPU on first use and
cached on the identity of the CSR arrays (pointer, size, k1, b, avgdl), so a loop over queries --
how the reference calls these functions -- pays the O(nnz) build once, not per query.  Mutating the
CSR arrays in place invalidates that assumption: call clear_index_cache().
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Tuple

import numpy as np
import torch

from . import _abi
from .index import TermMajorIndex, _cuda_device, _stream_ptr, _to_device, queries_from_dense

def set_int8_mma(enabled: bool) -> None:
    """Profiling / test hook: False forces the dp4a kernel for every shape (results are identical)."""
    _abi.lib.b2r_set_int8_mma(1 if enabled else 0)


def set_int8_cluster(max_cluster: int) -> None:
    """Profiling / test hook: largest thread-block cluster (TMA multicast group of query-tile CTAs) of the tcgen05
    scan kernel; 1 disables clusters.  Results are identical."""
    _abi.lib.b2r_set_int8_cluster(int(max_cluster))


def set_int8_pair(enabled: bool) -> None:
    """Profiling / test hook: True runs the fused INT8 scan of batches > 128 queries on CTA pairs (cta_group::2).
    Results are identical."""
    _abi.lib.b2r_set_int8_pair(1 if enabled else 0)


def set_bank_schedule(enabled) -> None:
    """Profiling / test hook: False / 0 makes later index builds keep dense segments doc-ascending; 16 selects the
    16-class schedule of round 1 (doc mod 16: the f64 accumulators' slots); True / anything else the default, 32
    classes (doc mod 32: also conflict-free for the f32 accumulators of the search path's pre-filter)."""
    _abi.lib.b2r_set_bank_schedule(int(enabled))


def set_int8_fused(mode) -> None:
    """Profiling / test hook: 0/False = plain chunked dense-tile + select path, otherwise the fused-selection
    path (default).  Results are identical."""
    _abi.lib.b2r_set_int8_fused(int(mode))


__all__ = ["set_int8_mma", "set_int8_fused", "set_int8_cluster", "set_int8_pair", "set_bank_schedule", "simd_bm25_score", "simd_bm25_batch_score", "fast_topk_selection", "simd_tfidf_score",
           "quantized_dot_product_batch", "optimized_bm25_score", "fast_topk", "clear_index_cache",
           "int8_scan_topk", "int8_rerank", "hybrid_search", "dense_topk"]

# key -> (index, the source arrays: holding them keeps their addresses from being reused)
_INDEX_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_INDEX_CACHE_MAX = 2


def clear_index_cache() -> None:
    _INDEX_CACHE.clear()


def _ident(a) -> tuple:
    if isinstance(a, torch.Tensor):
        return ("t", a.data_ptr(), tuple(a.shape), str(a.dtype))
    a = np.asarray(a)
    return ("n", a.__array_interface__["data"][0], a.shape, a.dtype.str)


def _fingerprint(a) -> int:
    """Cheap content fingerprint (64 strided samples + the last element) so that an in-place edit of a cached CSR
    array is noticed in the common cases; identity of the arrays stays the primary key."""
    n = int(a.shape[0]) if len(a.shape) else 0
    if n == 0:
        return 0
    step = max(1, n // 64)
    if isinstance(a, torch.Tensor):
        smp = torch.cat([a[::step][:64].reshape(-1), a[-1:].reshape(-1)]).detach().cpu().numpy()
    else:
        smp = np.concatenate([np.asarray(a)[::step][:64].reshape(-1), np.asarray(a)[-1:].reshape(-1)])
    return hash(smp.tobytes())


def _cached_index(kind, data, indices, indptr, doc_lengths, n_vocab, k1, b, avgdl) -> TermMajorIndex:
    """Index over the caller's CSR, cached on array identity + a sampled content fingerprint.  The reference
    silently skips CSR term ids >= len(query_tf) (`if term_idx < len(query_tf)`, retrieval.py:66): the index is
    therefore built over max(len(query_tf), max term id + 1) terms, and the extra terms can never be queried."""
    key = (kind, _ident(data), _ident(indices), _ident(indptr),
           _ident(doc_lengths) if doc_lengths is not None else None, int(n_vocab), float(k1), float(b), float(avgdl),
           _fingerprint(data), _fingerprint(indices), _fingerprint(indptr))
    hit = _INDEX_CACHE.get(key)
    ix = hit[0] if hit is not None else None
    if ix is None:
        if int(indptr[-1]) > 0:
            top = int(indices.max()) + 1
            n_vocab = max(int(n_vocab), top)
        ones = np.ones(n_vocab, np.float32)   # idf is a per-call input: set below
        ix = TermMajorIndex.from_csr(data, indices, indptr, doc_lengths, n_vocab=n_vocab, idf=ones,
                                     avgdl=avgdl, k1=k1, b=b, kind=kind, prefilter=False)   # dense scores only
        _INDEX_CACHE[key] = (ix, (data, indices, indptr, doc_lengths))
        while len(_INDEX_CACHE) > _INDEX_CACHE_MAX:
            _INDEX_CACHE.popitem(last=False)
    else:
        _INDEX_CACHE.move_to_end(key)
    return ix


def _fit(idf, n: int) -> np.ndarray:
    """idf cut / zero-padded to the index's vocabulary (terms beyond len(query_tf) are never queried)."""
    idf = np.asarray(idf, np.float32)
    return idf[:n] if len(idf) >= n else np.concatenate([idf, np.zeros(n - len(idf), np.float32)])


def _n_vocab_for(query_tf, idf_weights) -> int:
    # the reference guards `term_idx < len(query_tf)` and indexes idf_weights[term_idx]
    return int(np.shape(query_tf)[-1])


def simd_bm25_score(query_tf, doc_tf_data, doc_tf_indices, doc_tf_indptr, doc_lengths, idf_weights,
                    k1: float, b: float, avgdl: float):
    """BM25 scores f32[N] of one query (or f32[Q, N] for a [Q, V] batch of query vectors)."""
    was_tensor = isinstance(query_tf, torch.Tensor)
    q = query_tf.detach().cpu().numpy() if was_tensor else np.asarray(query_tf, dtype=np.float32)
    n_vocab = _n_vocab_for(q, idf_weights)
    ix = _cached_index("bm25", doc_tf_data, doc_tf_indices, doc_tf_indptr, doc_lengths, n_vocab, k1, b, avgdl)
    idf = idf_weights.detach().cpu().numpy() if isinstance(idf_weights, torch.Tensor) else np.asarray(idf_weights)
    ix.set_idf(_fit(idf, ix.n_vocab))
    ptr, terms, w = queries_from_dense(q)
    s = ix.score_dense(ptr, terms, w)
    if q.ndim == 1:
        s = s[0]
    return s if was_tensor else s.cpu().numpy()


simd_bm25_batch_score = simd_bm25_score


def simd_tfidf_score(query_tf, doc_tf_data, doc_tf_indices, doc_tf_indptr, idf_weights):
    """Impact-weighted sparse dot  sum tf * idf * qtf  (f32 products, f64 accumulate)."""
    was_tensor = isinstance(query_tf, torch.Tensor)
    q = query_tf.detach().cpu().numpy() if was_tensor else np.asarray(query_tf, dtype=np.float32)
    n_vocab = _n_vocab_for(q, idf_weights)
    ix = _cached_index("impact", doc_tf_data, doc_tf_indices, doc_tf_indptr, None, n_vocab, 0.0, 0.0, 0.0)
    idf = idf_weights.detach().cpu().numpy() if isinstance(idf_weights, torch.Tensor) else np.asarray(idf_weights)
    ix.set_idf(_fit(idf, ix.n_vocab))
    ptr, terms, w = queries_from_dense(q)
    s = ix.score_dense(ptr, terms, w)
    if q.ndim == 1:
        s = s[0]
    return s if was_tensor else s.cpu().numpy()


def fast_topk_selection(scores, k: int, index_dtype=np.int64) -> Tuple[np.ndarray, np.ndarray]:
    """k largest of scores[n], descending; ties broken by ascending index; k >= n returns all n
    sorted.  Returns (indices, scores[indices]).  Accepts a [rows, n] batch as well."""
    was_tensor = isinstance(scores, torch.Tensor)
    dev = _cuda_device(scores.device if was_tensor and scores.is_cuda else None)
    s = _to_device(scores, torch.float32, dev)
    one_row = s.dim() == 1
    if one_row:
        s = s[None, :]
    rows, n = int(s.shape[0]), int(s.shape[1])
    k = min(int(k), n)
    if k <= 0 or n == 0:
        e_i, e_v = np.zeros((rows, 0), index_dtype), np.zeros((rows, 0), np.float32)
        return (e_i[0], e_v[0]) if one_row else (e_i, e_v)
    nbytes = C.c_size_t(0)
    _abi.check(_abi.lib.b2r_topk_workspace(rows, n, k, C.byref(nbytes)), "top-k workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    idx = torch.empty((rows, k), dtype=torch.int64, device=dev)
    val = torch.empty((rows, k), dtype=torch.float32, device=dev)
    _abi.check(_abi.lib.b2r_topk(s.data_ptr(), rows, n, int(s.stride(0)), k, 0, None, idx.data_ptr(),
                                 val.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "top-k")
    if one_row:
        idx, val = idx[0], val[0]
    if was_tensor:
        return idx, val
    return idx.cpu().numpy().astype(index_dtype, copy=False), val.cpu().numpy()


def quantized_dot_product_batch(queries_int8, corpus_int8, query_scales, corpus_scales):
    """f32[Q, N] = f32((f64(int dot) * f64(qscale[q])) * f64(dscale[n]))."""
    was_tensor = isinstance(corpus_int8, torch.Tensor)
    dev = _cuda_device(corpus_int8.device if was_tensor and corpus_int8.is_cuda else None)
    q8 = _to_device(queries_int8, torch.int8, dev)
    d8 = _to_device(corpus_int8, torch.int8, dev)
    qs = _to_device(np.asarray(query_scales).reshape(-1) if not isinstance(query_scales, torch.Tensor)
                    else query_scales.reshape(-1), torch.float32, dev)
    ds = _to_device(np.asarray(corpus_scales).reshape(-1) if not isinstance(corpus_scales, torch.Tensor)
                    else corpus_scales.reshape(-1), torch.float32, dev)
    if q8.dim() != 2 or d8.dim() != 2 or q8.shape[1] != d8.shape[1]:
        raise ValueError("queries_int8 [Q, dim] and corpus_int8 [N, dim] must share dim")
    if qs.numel() != q8.shape[0] or ds.numel() != d8.shape[0]:
        raise ValueError("one scale per query row and per corpus row is required")
    out = torch.empty((q8.shape[0], d8.shape[0]), dtype=torch.float32, device=dev)
    _abi.check(_abi.lib.b2r_int8_dot_batch(q8.data_ptr(), q8.shape[0], d8.data_ptr(), d8.shape[0], q8.shape[1],
                                           qs.data_ptr(), ds.data_ptr(), out.data_ptr(), _stream_ptr(dev)),
               "int8 dot")
    return out if was_tensor else out.cpu().numpy()


def int8_scan_topk(queries_int8, corpus_int8, query_scales, corpus_scales, k: int, doc_id_base: int = 0):
    """Exhaustive INT8 scan + per-query top-k without materialising [Q, N].
    Returns CUDA tensors (idx i64[Q,k], val f32[Q,k], keys)."""
    dev = _cuda_device(corpus_int8.device if isinstance(corpus_int8, torch.Tensor) and corpus_int8.is_cuda else None)
    q8 = _to_device(queries_int8, torch.int8, dev)
    d8 = _to_device(corpus_int8, torch.int8, dev)
    qs = _to_device(query_scales, torch.float32, dev).reshape(-1)
    ds = _to_device(corpus_scales, torch.float32, dev).reshape(-1)
    nq, n, dim = int(q8.shape[0]), int(d8.shape[0]), int(d8.shape[1])
    k = min(int(k), n)
    nbytes = C.c_size_t(0)
    _abi.check(_abi.lib.b2r_int8_scan_workspace(nq, n, dim, k, C.byref(nbytes)), "int8 scan workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    val = torch.empty((nq, k), dtype=torch.float32, device=dev)
    _abi.check(_abi.lib.b2r_int8_scan_topk(q8.data_ptr(), nq, d8.data_ptr(), n, dim, qs.data_ptr(), ds.data_ptr(), k,
                                           int(doc_id_base), keys.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "int8 scan")
    return idx, val, keys


def dense_topk(embeddings, query_vectors, k: int, *, doc_id_base: int = 0, return_scores: bool = False):
    """fp32 similarities <embeddings[r], query> + top-k on the GPU (reference: search_by_vector, retrieval.py:402-436,
    np.dot through host BLAS).  `embeddings` f32[N, D] (a CUDA tensor stays where it is; anything else is uploaded),
    `query_vectors` f32[D] or f32[Q, D].  Returns CUDA tensors (idx i64[Q, k], val f32[Q, k][, scores f32[Q, N]])."""
    dev = _cuda_device(embeddings.device if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda else None)
    emb = _to_device(embeddings, torch.float32, dev)
    q = _to_device(query_vectors, torch.float32, dev)
    if q.dim() == 1:
        q = q[None, :]
    if emb.dim() != 2 or q.dim() != 2 or q.shape[1] != emb.shape[1]:
        raise ValueError("embeddings [N, D] and query_vectors [Q, D] must share D")
    n, dim, nq = int(emb.shape[0]), int(emb.shape[1]), int(q.shape[0])
    k = min(int(k), n)
    if k < 1:
        raise ValueError("k must be >= 1")
    nbytes = C.c_size_t(0)
    _abi.check(_abi.lib.b2r_f32_dot_topk_workspace(nq, n, k, C.byref(nbytes)), "dense top-k workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    stride = (n + 3) // 4 * 4
    scores = torch.empty((nq, stride), dtype=torch.float32, device=dev) if return_scores else None
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    val = torch.empty((nq, k), dtype=torch.float32, device=dev)
    _abi.check(_abi.lib.b2r_f32_dot_topk(emb.data_ptr(), n, dim, q.data_ptr(), nq, k, int(doc_id_base),
                                         scores.data_ptr() if scores is not None else None, stride, idx.data_ptr(),
                                         val.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "dense top-k")
    return (idx, val, scores[:, :n]) if return_scores else (idx, val)


def int8_rerank(cand_idx, cand_sparse, queries_int8, query_scales, corpus_int8, corpus_scales, k_out: int, *,
                sparse_weight: float = 0.3, dense_weight: float = 0.7, doc_id_base: int = 0, return_dense: bool = False):
    """Hybrid sparse -> dense rerank on candidates only (the reference only names this retriever:
    configs/ms_marco_paper_results.yaml:108-124, sparse_weight 0.3 / dense_weight 0.7).

    cand_idx i64[Q, k_in] are global document indices (e.g. TermMajorIndex.search output; -1 = none), cand_sparse
    their sparse scores f32[Q, k_in] or None (pure dense rerank).  For every pair the INT8 similarity of
    quantized_dot_product_batch is evaluated on the candidate's vector only; score = f32(f64(ws) * f64(sparse)
    + f64(wd) * f64(dense)).  Returns CUDA tensors (idx i64[Q, k_out], score f32[Q, k_out][, dense f32[Q, k_in]])."""
    dev = _cuda_device(corpus_int8.device if isinstance(corpus_int8, torch.Tensor) and corpus_int8.is_cuda else None)
    ci = _to_device(cand_idx, torch.int64, dev)
    cs = _to_device(cand_sparse, torch.float32, dev) if cand_sparse is not None else None
    q8 = _to_device(queries_int8, torch.int8, dev)
    d8 = _to_device(corpus_int8, torch.int8, dev)
    qs = _to_device(query_scales, torch.float32, dev).reshape(-1)
    ds = _to_device(corpus_scales, torch.float32, dev).reshape(-1)
    if ci.dim() != 2 or q8.dim() != 2 or d8.dim() != 2 or q8.shape[1] != d8.shape[1] or ci.shape[0] != q8.shape[0]:
        raise ValueError("cand_idx [Q, k_in], queries_int8 [Q, dim] and corpus_int8 [N, dim] do not fit together")
    if cs is not None and tuple(cs.shape) != tuple(ci.shape):
        raise ValueError("cand_sparse must have the shape of cand_idx")
    if qs.numel() != q8.shape[0] or ds.numel() != d8.shape[0]:
        raise ValueError("one scale per query row and per corpus row is required")
    nq, k_in = int(ci.shape[0]), int(ci.shape[1])
    k_out = min(int(k_out), k_in)
    if k_out < 1:
        raise ValueError("k_out must be >= 1")
    nbytes = C.c_size_t(0)
    _abi.check(_abi.lib.b2r_int8_rerank_workspace(nq, k_in, k_out, C.byref(nbytes)), "rerank workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    idx = torch.empty((nq, k_out), dtype=torch.int64, device=dev)
    val = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
    dense = torch.empty((nq, k_in), dtype=torch.float32, device=dev) if return_dense else None
    _abi.check(_abi.lib.b2r_int8_rerank(ci.data_ptr(), cs.data_ptr() if cs is not None else None, nq, k_in,
                                        q8.data_ptr(), qs.data_ptr(), d8.data_ptr(), ds.data_ptr(), int(d8.shape[0]),
                                        int(d8.shape[1]), int(doc_id_base), float(sparse_weight), float(dense_weight),
                                        k_out, dense.data_ptr() if dense is not None else None, idx.data_ptr(),
                                        val.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "int8 rerank")
    return (idx, val, dense) if return_dense else (idx, val)


def hybrid_search(index: TermMajorIndex, q_ptr, q_terms, q_weights, queries_int8, query_scales, corpus_int8,
                  corpus_scales, k_candidates: int, k_out: int, *, sparse_weight: float = 0.3,
                  dense_weight: float = 0.7):
    """BM25 (or impact) top-k_candidates on the term-major index, then int8_rerank on those candidates: the
    two-stage "hybrid" retriever of the reference's MS MARCO config, every stage on the GPU.  The corpus rows of
    `corpus_int8` are the documents of `index` (same shard, same order)."""
    idx, val = index.search(q_ptr, q_terms, q_weights, k_candidates)
    return int8_rerank(idx, val, queries_int8, query_scales, corpus_int8, corpus_scales, k_out,
                       sparse_weight=sparse_weight, dense_weight=dense_weight, doc_id_base=index.doc_id_base)


# ----------------------------------------------------------------------------- README aliases
def optimized_bm25_score(query_vector, doc_vectors, doc_lengths, idf_weights, *, k1: float = 1.2, b: float = 0.75,
                         avgdl=None):
    """README.md:129-155 alias.  `doc_vectors` is a scipy CSR matrix (or a (data, indices, indptr)
    triple, or a dense [N, V] array); k1 / b / avgdl -- free variables in the README pseudo-code --
    are keyword arguments; avgdl defaults to the reference's float(np.mean(doc_lengths))."""
    if hasattr(doc_vectors, "indptr"):
        data, indices, indptr = doc_vectors.data, doc_vectors.indices, doc_vectors.indptr
    elif isinstance(doc_vectors, (tuple, list)) and len(doc_vectors) == 3:
        data, indices, indptr = doc_vectors
    else:
        dense = np.asarray(doc_vectors, dtype=np.float32)
        rows, cols = np.nonzero(dense)
        data, indices = dense[rows, cols], cols.astype(np.int32)
        indptr = np.zeros(dense.shape[0] + 1, np.int64)
        np.cumsum(np.bincount(rows, minlength=dense.shape[0]), out=indptr[1:])
    if avgdl is None:
        avgdl = float(np.mean(np.asarray(doc_lengths, dtype=np.float32)))
    return simd_bm25_score(query_vector, data, indices, indptr, doc_lengths, idf_weights, k1, b, avgdl)


def fast_topk(scores, k: int):
    """README.md:187-203 alias: indices of the k largest scores, best first."""
    return fast_topk_selection(scores, k)[0]
