"""Seeded synthetic workloads of the shapes BASELINE.json names.

Array-level generators (they bypass the text tokenizer: the reference's text builder cannot reach
1M+ documents).  numpy only, so the same arrays feed the GPU path and the CPU oracle.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

__all__ = ["zipf_corpus", "zipf_queries", "impact_corpus", "impact_queries", "clustered_embeddings",
           "quantize_corpus", "quantize_queries", "fiqa_shape_corpus", "fiqa_shape_queries", "FIQA_SHAPE"]


def _zipf_cdf(n: int) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64)
    return np.cumsum(p / p.sum())


def zipf_corpus(n_docs: int, n_vocab: int, mean_len: float = 60.0, seed: int = 20260101, min_len: int = 5,
                max_len: int = 400, chunk_docs: int = 1 << 18):
    """Doc length ~ clip(floor(Gamma(2, mean/2)), min_len, max_len); tokens i.i.d. Zipf(s=1);
    tf = counts.  Returns CSR (data f32, indices i32 sorted per row, indptr i64) + doc_lengths f32."""
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(n_vocab)
    lens = np.clip(np.floor(rng.gamma(2.0, mean_len / 2.0, n_docs)), min_len, max_len).astype(np.int64)
    datas, inds = [], []
    row_nnz = np.zeros(n_docs, np.int64)
    for lo in range(0, n_docs, chunk_docs):
        hi = min(n_docs, lo + chunk_docs)
        l = lens[lo:hi]
        toks = np.minimum(np.searchsorted(cdf, rng.random(int(l.sum()))), n_vocab - 1).astype(np.int64)
        doc = np.repeat(np.arange(hi - lo, dtype=np.int64), l)
        key = doc * n_vocab + toks
        key.sort()
        first = np.empty(len(key), bool)
        first[:1] = True
        np.not_equal(key[1:], key[:-1], out=first[1:])
        starts = np.flatnonzero(first)
        uniq = key[starts]
        cnt = np.diff(np.append(starts, len(key)))
        row_nnz[lo:hi] = np.bincount(uniq // n_vocab, minlength=hi - lo)
        inds.append((uniq % n_vocab).astype(np.int32))
        datas.append(cnt.astype(np.float32))
    indptr = np.zeros(n_docs + 1, np.int64)
    np.cumsum(row_nnz, out=indptr[1:])
    return np.concatenate(datas), np.concatenate(inds), indptr, lens.astype(np.float32)


def zipf_queries(n_queries: int, n_vocab: int, seed: int = 20260102, min_terms: int = 4, max_terms: int = 8,
                 head_fraction: float = 0.25, uniform: bool = False):
    """n_t ~ U{min..max} distinct terms drawn Zipf-weighted from the top `head_fraction` ranks
    (the reference's own query law, tests/bm25_performance.py:262-272), weight 1.0.
    uniform=True draws terms uniformly over the whole vocabulary (tiny lists: the stress set)."""
    rng = np.random.default_rng(seed)
    head = n_vocab if uniform else max(max_terms, int(n_vocab * head_fraction))
    cdf = None if uniform else _zipf_cdf(head)
    ptr = np.zeros(n_queries + 1, np.int32)
    terms = []
    for q in range(n_queries):
        nt = int(rng.integers(min_terms, max_terms + 1))
        got = np.zeros(0, np.int64)
        while len(got) < nt:
            draw = (rng.integers(0, head, nt * 2) if uniform
                    else np.minimum(np.searchsorted(cdf, rng.random(nt * 2)), head - 1))
            _, first = np.unique(np.concatenate([got, draw]), return_index=True)
            got = np.concatenate([got, draw])[np.sort(first)][:nt]
        terms.append(np.sort(got).astype(np.int32))
        ptr[q + 1] = ptr[q] + nt
    q_terms = np.concatenate(terms)
    return ptr, q_terms, np.ones(len(q_terms), np.float32)


def impact_corpus(n_docs: int, n_vocab: int = 30522, nnz_per_doc: int = 120, seed: int = 20260103,
                  chunk_docs: int = 1 << 16):
    """SPLADE-shape: nnz_per_doc distinct Zipf terms per doc (oversample + dedupe), weights Gamma(2, 0.5)."""
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(n_vocab)
    datas, inds = [], []
    row_nnz = np.zeros(n_docs, np.int64)
    over = nnz_per_doc * 3
    for lo in range(0, n_docs, chunk_docs):
        hi = min(n_docs, lo + chunk_docs)
        n = hi - lo
        toks = np.minimum(np.searchsorted(cdf, rng.random(n * over)), n_vocab - 1).astype(np.int64)
        key = np.repeat(np.arange(n, dtype=np.int64), over) * n_vocab + toks
        key = np.unique(key)
        rows = key // n_vocab
        # keep at most nnz_per_doc terms per row (the first ones in term order after a shuffle-free cut)
        start = np.searchsorted(rows, np.arange(n))
        rank = np.arange(len(key)) - start[rows]
        keep = rank < nnz_per_doc
        key, rows = key[keep], rows[keep]
        row_nnz[lo:hi] = np.bincount(rows, minlength=n)
        inds.append((key % n_vocab).astype(np.int32))
        datas.append(rng.gamma(2.0, 0.5, len(key)).astype(np.float32))
    indptr = np.zeros(n_docs + 1, np.int64)
    np.cumsum(row_nnz, out=indptr[1:])
    return np.concatenate(datas), np.concatenate(inds), indptr


def impact_queries(n_queries: int, n_vocab: int = 30522, nnz_per_query: int = 30, seed: int = 20260104):
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(n_vocab)
    ptr = np.zeros(n_queries + 1, np.int32)
    terms, weights = [], []
    for q in range(n_queries):
        got = np.zeros(0, np.int64)
        while len(got) < nnz_per_query:
            draw = np.minimum(np.searchsorted(cdf, rng.random(nnz_per_query * 2)), n_vocab - 1)
            got = np.unique(np.concatenate([got, draw]))
        got = np.sort(rng.permutation(got)[:nnz_per_query])
        terms.append(got.astype(np.int32))
        weights.append(rng.gamma(2.0, 0.5, nnz_per_query).astype(np.float32) + np.float32(1e-3))
        ptr[q + 1] = ptr[q] + nnz_per_query
    return ptr, np.concatenate(terms), np.concatenate(weights)


def clustered_embeddings(n: int, dim: int = 768, n_clusters: int = 50, seed: int = 42) -> np.ndarray:
    """tests/embedding_quantizations.py:183-210 law: cluster centres N(0, 0.5^2), noise N(0, 0.2^2), unit norm."""
    rng = np.random.default_rng(seed)
    centers = rng.normal(0, 0.5, (n_clusters, dim)).astype(np.float32)
    x = centers[rng.integers(0, n_clusters, n)] + rng.normal(0, 0.2, (n, dim)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def quantize_corpus(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Corpus-side quantiser of the reference (retriever_registry.py:435-447): scale = max|x| (floored at
    1e-8, NOT divided by 127), q = round(x / scale * 127)."""
    scales = np.maximum(np.max(np.abs(x), axis=1, keepdims=True), 1e-8)
    return np.round(x / scales * 127.0).astype(np.int8), scales.flatten().astype(np.float32)


def quantize_queries(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Query-side quantiser (retriever_registry.py:482-485): scale = max|x| / 127, q = round(x / scale)."""
    scales = (np.max(np.abs(x), axis=1) / 127.0).astype(np.float32)
    scales = np.where(scales == 0, np.float32(1.0), scales)
    return np.clip(np.round(x / scales[:, None]), -127, 127).astype(np.int8), scales


# ------------------------------------------------------------------------------------------------
# BASELINE config 1: the FiQA-shape TEXT corpus of the reference's own generator
# (tests/core_test.py:203-252, SyntheticDataGenerator: seeds 42 / 43).  The reference draws
# `np.random.choice(vocab, size=L, p=zipf)` per document after `np.random.seed(seed)`; the legacy
# RandomState stream is frozen across numpy versions, and choice-with-p is `cdf.searchsorted(
# random_sample(L), side="right")` on the normalised cumulative sum, so the generator below produces
# the same texts word for word (oracle/gen_golden.py checks the sha256 of both) at a fifth of the cost.
FIQA_SHAPE = dict(num_docs=57_638, avg_doc_length=132, vocab_size=60_000, corpus_seed=42,
                  num_queries=648, avg_query_length=11, query_seed=43)


def _legacy_zipf_cdf(n: int) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1)
    p /= p.sum()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf


def fiqa_shape_corpus(num_docs: int = FIQA_SHAPE["num_docs"], avg_doc_length: int = FIQA_SHAPE["avg_doc_length"],
                      vocab_size: int = FIQA_SHAPE["vocab_size"], seed: int = FIQA_SHAPE["corpus_seed"]):
    """Dict doc_id -> {"text", "title"} identical to SyntheticDataGenerator.generate_corpus (core_test.py:206-228)."""
    rs = np.random.RandomState(seed)
    cdf = _legacy_zipf_cdf(vocab_size)
    corpus = {}
    for d in range(num_docs):
        length = max(10, int(rs.gamma(2, avg_doc_length / 2)))
        ids = cdf.searchsorted(rs.random_sample(length), side="right")
        corpus[f"doc_{d}"] = {"text": " ".join([f"word_{w}" for w in ids]), "title": f"Document {d}"}
    return corpus


def fiqa_shape_queries(num_queries: int = FIQA_SHAPE["num_queries"], avg_query_length: int = FIQA_SHAPE["avg_query_length"],
                       vocab_size: int = FIQA_SHAPE["vocab_size"], seed: int = FIQA_SHAPE["query_seed"]):
    """Dict qid -> text identical to SyntheticDataGenerator.generate_queries (core_test.py:230-252)."""
    rs = np.random.RandomState(seed)
    cdf = _legacy_zipf_cdf(vocab_size // 10)
    queries = {}
    for q in range(num_queries):
        length = max(1, int(rs.gamma(1.5, avg_query_length / 1.5)))
        ids = cdf.searchsorted(rs.random_sample(length), side="right")
        queries[f"query_{q}"] = " ".join([f"word_{w}" for w in ids])
    return queries


def corpus_sha256(corpus) -> str:
    """sha256 over "<doc_id>\t<text>\n" of every document in insertion order (pins a generated text corpus)."""
    import hashlib
    h = hashlib.sha256()
    for doc_id, doc in corpus.items():
        h.update(doc_id.encode())
        h.update(b"\t")
        h.update(doc.get("text", "").encode())
        h.update(b"\n")
    return h.hexdigest()
