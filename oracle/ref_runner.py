"""Runs the reference's OWN kernels (oracle/_ref, see oracle/make_ref.py) on packed queries: the CPU baseline of
kind "reference".  TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference legs and tests).

search_batch() is RetrievalService._score_bm25_query (rag_system/core/retrieval.py:233-284) at array level: a dense
f32 query_tf vector per query, simd_bm25_score over the doc-major CSR (retrieval.py:41-76), then
fast_topk_selection (retrieval.py:79-92) -- the two Numba functions are called unmodified."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_REF_ROOT = os.path.join(HERE, "_ref")
_mod = None


def available() -> bool:
    return os.path.isfile(os.path.join(_REF_ROOT, "rag_system", "core", "retrieval.py"))


def module():
    """The reference's retrieval module (imports numba; first call JIT-compiles on use)."""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref is empty (run oracle/make_ref.py where /root/reference exists)")
        if _REF_ROOT not in sys.path:
            sys.path.insert(0, _REF_ROOT)
        import logging
        logging.getLogger("rag_system").setLevel(logging.WARNING)
        from rag_system.core import retrieval as m
        if not m.NUMBA_AVAILABLE:
            raise RuntimeError("numba is missing: the reference would take its numpy fallback")
        _mod = m
    return _mod


def use_all_host_threads() -> int:
    import numba
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = max(1, min(n, numba.config.NUMBA_NUM_THREADS))
    numba.set_num_threads(n)
    return n


def search_batch(q_ptr, q_terms, q_weights, n_vocab, data, indices, indptr, doc_lengths, idf, k1, b, avgdl, k):
    """(idx i64[Q, k], val f32[Q, k]) in the reference's own order (ties unspecified, retrieval.py:79-92)."""
    m = module()
    nq = len(q_ptr) - 1
    idx = np.full((nq, k), -1, np.int64)
    val = np.zeros((nq, k), np.float32)
    data = np.ascontiguousarray(data, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    indptr = np.ascontiguousarray(indptr)
    doc_lengths = np.ascontiguousarray(doc_lengths, np.float32)
    idf = np.ascontiguousarray(idf, np.float32)
    for q in range(nq):
        qtf = np.zeros(n_vocab, np.float32)
        s, e = int(q_ptr[q]), int(q_ptr[q + 1])
        qtf[q_terms[s:e]] = q_weights[s:e]
        scores = m.simd_bm25_score(qtf, data, indices, indptr, doc_lengths, idf, k1, b, avgdl)
        ti, tv = m.fast_topk_selection(scores, k)
        idx[q, :len(ti)], val[q, :len(tv)] = ti, tv
    return idx, val
