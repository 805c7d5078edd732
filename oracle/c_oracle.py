"""ctypes front-end of oracle/liboracle.so (the C restatement in bm25_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of bm25_oracle.c.  Used by tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bm25_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_topk.restype = C.c_int64
        _LIB.oracle_num_threads.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def use_all_host_threads() -> int:
    """bench.py's CPU legs: use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it
    to 1 for every rank).  Returns the thread count in effect."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:      # pragma: no cover
        n = os.cpu_count() or 1
    lib().oracle_set_num_threads(int(n))
    return num_threads()


def bm25_scores(query_tf, data, indices, indptr, doc_lengths, idf, k1, b, avgdl):
    query_tf = np.ascontiguousarray(query_tf, np.float32)
    data = np.ascontiguousarray(data, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    indptr = np.ascontiguousarray(indptr, np.int64)
    doc_lengths = np.ascontiguousarray(doc_lengths, np.float32)
    idf = np.ascontiguousarray(idf, np.float32)
    n = len(indptr) - 1
    out = np.empty(n, np.float32)
    lib().oracle_bm25_score(_p(query_tf, C.c_float), C.c_int32(len(query_tf)), _p(data, C.c_float),
                            _p(indices, C.c_int32), _p(indptr, C.c_int64), C.c_int64(n),
                            _p(doc_lengths, C.c_float), _p(idf, C.c_float), C.c_double(k1), C.c_double(b),
                            C.c_double(avgdl), _p(out, C.c_float))
    return out


def tfidf_scores(query_tf, data, indices, indptr, idf):
    query_tf = np.ascontiguousarray(query_tf, np.float32)
    data = np.ascontiguousarray(data, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    indptr = np.ascontiguousarray(indptr, np.int64)
    idf = np.ascontiguousarray(idf, np.float32)
    n = len(indptr) - 1
    out = np.empty(n, np.float32)
    lib().oracle_tfidf_score(_p(query_tf, C.c_float), C.c_int32(len(query_tf)), _p(data, C.c_float),
                             _p(indices, C.c_int32), _p(indptr, C.c_int64), C.c_int64(n),
                             _p(idf, C.c_float), _p(out, C.c_float))
    return out


def int8_dot_batch(q8, d8, q_scales, d_scales):
    q8 = np.ascontiguousarray(q8, np.int8)
    d8 = np.ascontiguousarray(d8, np.int8)
    q_scales = np.ascontiguousarray(q_scales, np.float32)
    d_scales = np.ascontiguousarray(d_scales, np.float32)
    out = np.empty((q8.shape[0], d8.shape[0]), np.float32)
    lib().oracle_int8_dot_batch(_p(q8, C.c_int8), C.c_int32(q8.shape[0]), _p(d8, C.c_int8),
                                C.c_int64(d8.shape[0]), C.c_int32(q8.shape[1]), _p(q_scales, C.c_float),
                                _p(d_scales, C.c_float), _p(out, C.c_float))
    return out


def topk(scores, k):
    scores = np.ascontiguousarray(scores, np.float32)
    k = min(int(k), len(scores))
    idx = np.empty(k, np.int64)
    val = np.empty(k, np.float32)
    m = lib().oracle_topk(_p(scores, C.c_float), C.c_int64(len(scores)), C.c_int64(k), _p(idx, C.c_int64),
                          _p(val, C.c_float))
    return idx[:m], val[:m]


def bm25_search_batch(q_ptr, q_terms, q_weights, n_vocab, data, indices, indptr, doc_lengths, idf, k1, b, avgdl, k):
    q_ptr = np.ascontiguousarray(q_ptr, np.int32)
    q_terms = np.ascontiguousarray(q_terms, np.int32)
    q_weights = np.ascontiguousarray(q_weights, np.float32)
    data = np.ascontiguousarray(data, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    indptr = np.ascontiguousarray(indptr, np.int64)
    doc_lengths = np.ascontiguousarray(doc_lengths, np.float32)
    idf = np.ascontiguousarray(idf, np.float32)
    nq = len(q_ptr) - 1
    n = len(indptr) - 1
    idx = np.empty((nq, k), np.int64)
    val = np.empty((nq, k), np.float32)
    lib().oracle_bm25_search_batch(_p(q_ptr, C.c_int32), _p(q_terms, C.c_int32), _p(q_weights, C.c_float),
                                   C.c_int32(nq), C.c_int32(n_vocab), _p(data, C.c_float), _p(indices, C.c_int32),
                                   _p(indptr, C.c_int64), C.c_int64(n), _p(doc_lengths, C.c_float),
                                   _p(idf, C.c_float), C.c_double(k1), C.c_double(b), C.c_double(avgdl),
                                   C.c_int64(k), _p(idx, C.c_int64), _p(val, C.c_float))
    return idx, val
