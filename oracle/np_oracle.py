"""NumPy CPU restatement of the reference's retrieval scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
(or the CPU baseline), never as the thing shipped.

Parity pin: every function here is checked bit-for-bit against outputs of the
reference's own Numba kernels, generated in the build container by
``oracle/gen_golden.py`` (which imports ``/root/reference``) and committed under
``tests/golden/``.  The reference's own tests hold no golden vectors for these
functions, so the pin is "outputs of the reference
itself run here".

Reference lines restated (paths relative to the reference root):

* ``bm25_scores``       <- ``rag_system/core/retrieval.py:41-76``  (simd_bm25_score)
* ``tfidf_scores``      <- ``rag_system/pipeline/evaluate_rag_pipeline.py:95-121`` (simd_tfidf_score)
* ``int8_dot_batch``    <- ``rag_system/core/retriever_registry.py:90-117`` (quantized_dot_product_batch)
* ``quantize_rows``     <- ``rag_system/core/retriever_registry.py:435-447`` (_quantize_embeddings)
* ``topk_canonical``    <- ``rag_system/core/retrieval.py:79-92`` (fast_topk_selection) with the
  stated tie-break (score desc, doc index asc); the reference leaves ties unspecified.
* ``idf_from_csr`` / ``avgdl_from_lengths`` <- ``rag_system/core/retrieval.py:187-190``
* ``tokenize`` / ``build_text_index``       <- ``rag_system/core/retrieval.py:141-184``
"""
from __future__ import annotations

import re
from collections import Counter
from typing import Dict, Iterable, List, Tuple

import numpy as np

__all__ = [
    "bm25_scores", "tfidf_scores", "int8_dot_batch", "quantize_rows", "topk_canonical",
    "idf_from_csr", "avgdl_from_lengths", "tokenize", "build_text_index", "csr_to_csc",
    "dense_query", "search_bm25_text",
]


# --------------------------------------------------------------------------- layout helpers
def csr_to_csc(data: np.ndarray, indices: np.ndarray, indptr: np.ndarray, n_vocab: int):
    """Stable doc-major -> term-major transpose (each term list is doc-ascending)."""
    n_docs = len(indptr) - 1
    rows = np.repeat(np.arange(n_docs, dtype=np.int64), np.diff(indptr).astype(np.int64))
    order = np.argsort(indices, kind="stable")
    t_ptr = np.zeros(n_vocab + 1, dtype=np.int64)
    np.cumsum(np.bincount(indices, minlength=n_vocab), out=t_ptr[1:])
    return data[order], rows[order], t_ptr


def dense_query(terms: Iterable[int], weights: Iterable[float], n_vocab: int) -> np.ndarray:
    q = np.zeros(n_vocab, dtype=np.float32)
    for t, w in zip(terms, weights):
        q[int(t)] = np.float32(w)
    return q


# --------------------------------------------------------------------------- K1
def bm25_scores(query_tf, data, indices, indptr, doc_lengths, idf, k1, b, avgdl, csc=None):
    """f32[N] BM25 scores of one query; f64 accumulate in ascending term id, one f32 rounding.

    Follows retrieval.py:52-74: k1/b/avgdl are Python floats, so every product is f64:
      norm = k1 * (1.0 - b + b * dl / avgdl);  s += idf * ((tf * (k1 + 1.0)) / (tf + norm)) * qtf
    Only terms with query_tf > 0 contribute (retrieval.py:67).
    """
    n_docs = len(indptr) - 1
    k1 = float(k1); b = float(b); avgdl = float(avgdl)
    if csc is None:
        csc = csr_to_csc(data, indices, indptr, len(query_tf))
    c_tf, c_doc, c_ptr = csc
    dl = doc_lengths.astype(np.float64)
    norm = k1 * ((1.0 - b) + (b * dl) / avgdl)
    acc = np.zeros(n_docs, dtype=np.float64)
    for t in np.flatnonzero(query_tf > 0):
        s, e = c_ptr[t], c_ptr[t + 1]
        if s == e:
            continue
        docs = c_doc[s:e]
        tf = c_tf[s:e].astype(np.float64)
        contrib = (np.float64(idf[t]) * ((tf * (k1 + 1.0)) / (tf + norm[docs]))) * np.float64(query_tf[t])
        acc[docs] += contrib          # docs are unique inside one term list
    return acc.astype(np.float32)


# --------------------------------------------------------------------------- K3
def tfidf_scores(query_tf, data, indices, indptr, idf, csc=None):
    """f32[N] impact-weighted sparse dot (evaluate_rag_pipeline.py:102-121).

    All three factors are f32, so ``tf * idf * qtf`` is two f32 roundings; the running sum is
    f64 (``doc_score = 0.0``) and the store is f32.  The kernel is ``@njit(fastmath=True)`` and LLVM
    reassociates the f32 product to ``(tf * qtf) * idf`` (query_tf[term] is already loaded for the
    ``> 0`` test): measured against the reference run in the build container, that order is
    bit-identical on 18,000 / 18,000 fractional-weight scores, the source order is not (57 % match).
    """
    n_docs = len(indptr) - 1
    if csc is None:
        csc = csr_to_csc(data, indices, indptr, len(query_tf))
    c_tf, c_doc, c_ptr = csc
    acc = np.zeros(n_docs, dtype=np.float64)
    for t in np.flatnonzero(query_tf > 0):
        s, e = c_ptr[t], c_ptr[t + 1]
        if s == e:
            continue
        prod = (c_tf[s:e].astype(np.float32) * np.float32(query_tf[t])) * np.float32(idf[t])
        acc[c_doc[s:e]] += prod.astype(np.float64)
    return acc.astype(np.float32)


# --------------------------------------------------------------------------- K4
def int8_dot_batch(q8, d8, q_scales, d_scales):
    """f32[Q,N]: exact integer dot, then (dot * qs) * ds in f64, f32 store
    (retriever_registry.py:101-115; Numba types int64 * f32 as f64)."""
    dots = q8.astype(np.int64) @ d8.astype(np.int64).T
    out = (dots.astype(np.float64) * q_scales.astype(np.float64)[:, None]) * d_scales.astype(np.float64)[None, :]
    return out.astype(np.float32)


def quantize_rows(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Symmetric per-row INT8 quantiser of the corpus side (retriever_registry.py:435-447):
    scale = max(max|x|, 1e-8), q = round(x / scale * 127) as int8 (scale is NOT divided by 127)."""
    scales = np.max(np.abs(x), axis=1, keepdims=True)
    scales = np.maximum(scales, 1e-8)
    q = np.round(x / scales * 127.0).astype(np.int8)
    return q, scales.flatten().astype(np.float32)


# --------------------------------------------------------------------------- K2
def topk_canonical(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """k largest, score descending then index ascending; k >= n returns all n
    (retrieval.py:82-92 with the stated tie-break).  -0.0 ties with +0.0; NaN sorts last."""
    scores = np.asarray(scores, dtype=np.float32)
    n = len(scores)
    k = min(int(k), n)
    neg = -scores
    neg = np.where(neg == 0, np.float32(0.0), neg)      # fold -0.0 / +0.0
    order = np.lexsort((np.arange(n), neg))
    idx = order[:k].astype(np.int64)
    return idx, scores[idx]


# --------------------------------------------------------------------------- host expressions
def idf_from_csr(indices: np.ndarray, n_docs: int, n_vocab: int) -> np.ndarray:
    """retrieval.py:187-189 verbatim: Robertson/Sparck-Jones idf, no +1, cast to f32."""
    df = np.bincount(indices, minlength=n_vocab)
    return np.log((n_docs - df + 0.5) / (df + 0.5)).astype(np.float32)


def avgdl_from_lengths(doc_lengths: np.ndarray) -> float:
    """retrieval.py:190: a float32 mean widened to a Python float."""
    return float(np.mean(doc_lengths.astype(np.float32)))


_TOKEN = re.compile(r"\b\w+\b")


def tokenize(text: str) -> List[str]:
    """retrieval.py:148 / :236."""
    return _TOKEN.findall(text.lower())


def build_text_index(corpus: Dict[str, Dict]):
    """retrieval.py:141-190 restated with arrays instead of scipy: returns a dict with
    data/indices/indptr (rows sorted by term id), doc_lengths, idf, avgdl, vocabulary, doc_ids."""
    doc_ids = list(corpus.keys())
    toks = []
    vocab = set()
    for d in doc_ids:
        doc = corpus[d]
        text = doc.get("text", doc.get("content", doc.get("body", "")))
        t = tokenize(text) if text else []
        toks.append(t)
        vocab.update(t)
    vocabulary = {term: i for i, term in enumerate(sorted(vocab))}
    indptr = np.zeros(len(doc_ids) + 1, dtype=np.int64)
    cols: List[int] = []
    vals: List[float] = []
    doc_lengths = np.zeros(len(doc_ids), dtype=np.float32)
    for i, t in enumerate(toks):
        doc_lengths[i] = len(t)
        row = sorted((vocabulary[w], c) for w, c in Counter(t).items())
        cols.extend(c for c, _ in row)
        vals.extend(float(v) for _, v in row)
        indptr[i + 1] = len(cols)
    indices = np.asarray(cols, dtype=np.int32)
    data = np.asarray(vals, dtype=np.float32)
    return dict(data=data, indices=indices, indptr=indptr, doc_lengths=doc_lengths,
                idf=idf_from_csr(indices, len(doc_ids), len(vocabulary)),
                avgdl=avgdl_from_lengths(doc_lengths), vocabulary=vocabulary, doc_ids=doc_ids)


def search_bm25_text(ix, queries: Dict[str, str], top_k: int = 10, k1: float = 1.2, b: float = 0.75):
    """retrieval.py:203-296 without the cache: blank / out-of-vocab query -> {}, keep score > 0,
    dict ordered by (score desc, doc index asc)."""
    out: Dict[str, Dict[str, float]] = {}
    n_vocab = len(ix["vocabulary"])
    csc = csr_to_csc(ix["data"], ix["indices"], ix["indptr"], n_vocab)
    for qid, text in queries.items():
        if not text or not text.strip():
            out[qid] = {}
            continue
        counts = Counter(tokenize(text))
        q = np.zeros(n_vocab, dtype=np.float32)
        hit = 0
        for term, c in counts.items():
            if term in ix["vocabulary"]:
                q[ix["vocabulary"][term]] = float(c)
                hit += 1
        if hit == 0:
            out[qid] = {}
            continue
        s = bm25_scores(q, ix["data"], ix["indices"], ix["indptr"], ix["doc_lengths"], ix["idf"],
                        k1, b, ix["avgdl"], csc=csc)
        idx, val = topk_canonical(s, top_k)
        out[qid] = {ix["doc_ids"][i]: float(v) for i, v in zip(idx, val) if v > 0}
    return out


def hybrid_rerank(cand_idx, cand_sparse, q8, q_scales, d8, d_scales, sparse_weight, dense_weight, k_out,
                  doc_id_base=0):
    """Candidate-only INT8 rerank + linear score fusion.  The reference NAMES this retriever
    (rag_system/configs/ms_marco_paper_results.yaml:108-124: type "hybrid", sparse_weight 0.3, dense_weight 0.7)
    but ships no implementation, so there is nothing to pin it to -- parity unpinned for the fusion rule; the
    dense part is quantized_dot_product_batch (retriever_registry.py:90-117, pinned by tests/golden/int8.npz)
    restricted to the candidates.  Semantics checked against the CUDA path:
        dense  = f32((f64(int dot) * f64(qs)) * f64(ds))
        score  = f32(f64(ws) * f64(sparse) + f64(wd) * f64(dense))                (dense alone if cand_sparse is None)
    rank by score descending, document index ascending; candidates < 0 or outside the shard are skipped.
    Returns (idx i64[Q, k_out] padded with -1, score f32[Q, k_out] padded with -inf, dense f32[Q, k_in])."""
    cand_idx = np.asarray(cand_idx, np.int64)
    nq, k_in = cand_idx.shape
    ws, wd = np.float64(sparse_weight), np.float64(dense_weight)
    out_i = np.full((nq, k_out), -1, np.int64)
    out_v = np.full((nq, k_out), -np.inf, np.float32)
    dense_all = np.full((nq, k_in), -np.inf, np.float32)
    for q in range(nq):
        loc = cand_idx[q] - doc_id_base
        ok = (cand_idx[q] >= 0) & (loc >= 0) & (loc < len(d8))
        if not ok.any():
            continue
        rows = loc[ok]
        dots = d8[rows].astype(np.int64) @ q8[q].astype(np.int64)
        dense = ((dots.astype(np.float64) * np.float64(q_scales[q])) * d_scales[rows].astype(np.float64)).astype(np.float32)
        dense_all[q, ok] = dense
        if cand_sparse is None:
            score = dense
        else:
            score = (ws * np.asarray(cand_sparse, np.float32)[q, ok].astype(np.float64)
                     + wd * dense.astype(np.float64)).astype(np.float32)
        ids = cand_idx[q, ok]
        order = np.lexsort((ids, -np.where(score == 0, np.float32(0), score)))[:k_out]
        out_i[q, :len(order)] = ids[order]
        out_v[q, :len(order)] = np.where(score[order] == 0, np.float32(0), score[order])
    return out_i, out_v, dense_all
