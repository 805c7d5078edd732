"""RetrievalService drop-in (reference: rag_system/core/retrieval.py:95-506).

Same constructor, attributes, return shapes and error behaviour as the reference class; what changes
is where the work happens: build_bm25_index lays the corpus out term-major in HBM, and search_bm25
scores ALL non-cached queries of a call in one batched GPU pass (the reference loops over queries
and rescans the whole CSR for each).  Host semantics kept byte-for-byte: tokeniser, sorted vocabulary,
f32 doc lengths, RSJ idf, f32-mean avgdl, query cache keyed "<stripped text>:<top_k>" (<= 1000
entries, never evicted), blank / out-of-vocabulary query -> {}, only score > 0 in the result dict,
dict order = rank.  Tie-break between equal scores: ascending document index (the reference leaves
it unspecified).
"""
from __future__ import annotations

import logging
import re
import threading
import time
from collections import Counter
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
from scipy.sparse import csr_matrix

from .docstore import Document, MemoryIndex
from .index import TermMajorIndex, pack_queries, reference_avgdl, reference_idf
from .kernels import dense_topk, fast_topk_selection

logger = logging.getLogger(__name__)
_WORD = re.compile(r"\b\w+\b")
NUMBA_AVAILABLE = False      # get_stats() key kept for callers; the B200 path does not use Numba


class RetrievalService:
    def __init__(self, index_path: Union[str, Path], embedding_path: Optional[Union[str, Path]] = None,
                 num_workers: int = 4, cache_size: int = 1000):
        self.index = MemoryIndex(index_path)
        self.embedding_path = Path(embedding_path) if embedding_path else None
        self.num_workers = num_workers
        self.cache_size = cache_size
        self.logger = logging.getLogger(__name__)

        self.corpus_tf: Optional[csr_matrix] = None
        self.vocabulary: Dict[str, int] = {}
        self.idf_weights: Optional[np.ndarray] = None
        self.doc_lengths: Optional[np.ndarray] = None
        self.doc_ids: List[str] = []
        self.avgdl: float = 0.0
        self.k1: float = 1.2
        self.b: float = 0.75

        self._cache: Dict[str, Document] = {}
        self.query_cache: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}
        self.cache_lock = threading.RLock()

        self.tile_docs = 4096
        self._gpu_index: Optional[TermMajorIndex] = None
        self._gpu_params: Optional[tuple] = None

        self.embedding_index = None
        self._gpu_embeddings = None
        if self.embedding_path and self.embedding_path.exists():
            self._load_embeddings()

    # ------------------------------------------------------------------ index build
    def build_bm25_index(self, corpus: Dict[str, Dict]) -> None:
        """retrieval.py:129-201.  Host part identical in effect; then the CSR goes to the GPU."""
        t0 = time.perf_counter()
        if not corpus:
            raise ValueError("Empty corpus provided")
        self.doc_ids = list(corpus.keys())
        per_doc: List[Counter] = []
        lengths = np.zeros(len(self.doc_ids), dtype=np.float32)
        vocab = set()
        for i, doc_id in enumerate(self.doc_ids):
            doc = corpus[doc_id]
            text = doc.get("text", doc.get("content", doc.get("body", "")))
            if text:
                tokens = _WORD.findall(text.lower())
                lengths[i] = len(tokens)
                c = Counter(tokens)
                vocab.update(c)
            else:
                c = Counter()
            per_doc.append(c)
        self.vocabulary = {term: idx for idx, term in enumerate(sorted(vocab))}
        n_vocab = len(self.vocabulary)
        self.doc_lengths = lengths

        indptr = np.zeros(len(per_doc) + 1, dtype=np.int64)
        np.cumsum([len(c) for c in per_doc], out=indptr[1:])
        cols = np.empty(indptr[-1], dtype=np.int32)
        vals = np.empty(indptr[-1], dtype=np.float32)
        voc = self.vocabulary
        for i, c in enumerate(per_doc):
            s = indptr[i]
            if c:
                cols[s:s + len(c)] = [voc[t] for t in c]
                vals[s:s + len(c)] = list(c.values())
        self.corpus_tf = csr_matrix((vals, cols, indptr), shape=(len(per_doc), max(n_vocab, 1)), dtype=np.float32)
        self.corpus_tf.sort_indices()
        self.corpus_tf.eliminate_zeros()
        self._finish_build()
        self.logger.info("BM25 index built in %.2fs (%d docs, %d terms, %d postings, %.1f MB in HBM)",
                         time.perf_counter() - t0, len(self.doc_ids), n_vocab, self.corpus_tf.nnz,
                         self._gpu_index.device_bytes() / 2 ** 20)

    def build_from_csr(self, data, indices, indptr, doc_lengths, *, n_vocab: Optional[int] = None, idf=None,
                       avgdl: Optional[float] = None, doc_ids: Optional[List[str]] = None,
                       vocabulary: Optional[Dict[str, int]] = None) -> None:
        """Array-level entry: the reference's text builder cannot reach the
        1M / 8.8M-document configurations; idf / avgdl default to the reference's host expressions
        (retrieval.py:187-190)."""
        n_docs = len(indptr) - 1
        if n_docs <= 0:
            raise ValueError("Empty corpus provided")
        if n_vocab is None:
            n_vocab = int(np.max(indices)) + 1 if len(indices) else 1
        self.corpus_tf = csr_matrix((np.asarray(data, np.float32), np.asarray(indices, np.int32),
                                     np.asarray(indptr)), shape=(n_docs, n_vocab))
        self.doc_lengths = np.asarray(doc_lengths, dtype=np.float32)
        self.doc_ids = list(doc_ids) if doc_ids is not None else [str(i) for i in range(n_docs)]
        self.vocabulary = dict(vocabulary) if vocabulary is not None else {}
        self._finish_build(idf=idf, avgdl=avgdl)

    def _finish_build(self, idf=None, avgdl=None) -> None:
        n_docs, n_vocab = self.corpus_tf.shape
        self.idf_weights = (np.asarray(idf, np.float32) if idf is not None
                            else reference_idf(self.corpus_tf.indices, n_docs, n_vocab))
        self.avgdl = float(avgdl) if avgdl is not None else reference_avgdl(self.doc_lengths)
        with self.cache_lock:
            self.query_cache.clear()
        self._sync_gpu_index(force=True)

    def _sync_gpu_index(self, force: bool = False) -> TermMajorIndex:
        """(Re)build the HBM index when k1 / b / avgdl were changed on the instance after the build
        -- they are plain attributes in the reference (retrieval.py:116-117) and the per-posting BM25
        factor is precomputed from them."""
        params = (float(self.k1), float(self.b), float(self.avgdl))
        if force or self._gpu_index is None or params != self._gpu_params:
            tf = self.corpus_tf
            self._gpu_index = TermMajorIndex.from_csr(
                tf.data, tf.indices, tf.indptr, self.doc_lengths, n_vocab=tf.shape[1], idf=self.idf_weights,
                avgdl=self.avgdl, k1=self.k1, b=self.b, kind="bm25", tile_docs=self.tile_docs)
            self._gpu_params = params
        return self._gpu_index

    # ------------------------------------------------------------------ save / load
    def save_bm25_index(self, path: Union[str, Path]) -> None:
        """Persist the built index: `<path>` holds the term-major HBM layout (TermMajorIndex.save), and
        `<path>.meta.npz` the host attributes under the keys of the reference's own index cache
        (evaluate_rag_pipeline.py:280-293: tf_data, tf_indices, tf_indptr, tf_shape, doc_lengths, idf,
        vocabulary, doc_ids, avgdl) plus k1, b -- so a cache written by the reference loads here too."""
        if self.corpus_tf is None:
            raise ValueError("BM25 index not built. Call build_bm25_index() first.")
        path = Path(path)
        if str(path).endswith(".npz"):
            raise ValueError("save_bm25_index: the path names the binary HBM layout and must not end in .npz "
                             "(the host attributes go to '<path>.meta.npz'; a .npz path is read back as a "
                             "reference-style CSR cache)")
        self._sync_gpu_index().save(path)
        tf = self.corpus_tf
        with open(str(path) + ".meta.npz", "wb") as f:
            np.savez_compressed(
                f, tf_data=tf.data, tf_indices=tf.indices, tf_indptr=tf.indptr, tf_shape=np.asarray(tf.shape),
                doc_lengths=self.doc_lengths, idf=self.idf_weights,
                vocabulary=np.asarray(sorted(self.vocabulary, key=self.vocabulary.get), dtype=np.str_),
                doc_ids=np.asarray([str(d) for d in self.doc_ids], dtype=np.str_), avgdl=self.avgdl, k1=self.k1,
                b=self.b)          # unicode arrays: nothing in the file needs pickle

    def load_bm25_index(self, path: Union[str, Path], verify: bool = True) -> None:
        """Inverse of save_bm25_index.  The HBM layout is read back as is (no build kernels) when `<path>` exists
        and was built with the same k1 / b / avgdl; otherwise (e.g. a cache file of the reference: only the
        .npz) it is rebuilt from the CSR."""
        path = Path(path)
        meta_path = Path(str(path) + ".meta.npz") if not str(path).endswith(".npz") else path
        cached = np.load(meta_path, allow_pickle=False)
        try:
            cached["vocabulary"], cached["doc_ids"]
        except ValueError:
            # a cache written by the REFERENCE stores these two as object arrays (evaluate_rag_pipeline.py:280-293):
            # only such a file is opened with pickle enabled -- load reference caches from trusted paths only
            cached = np.load(meta_path, allow_pickle=True)
        shape = tuple(int(x) for x in cached["tf_shape"])
        self.corpus_tf = csr_matrix((cached["tf_data"], cached["tf_indices"], cached["tf_indptr"]), shape=shape)
        self.doc_lengths = np.asarray(cached["doc_lengths"], dtype=np.float32)
        self.idf_weights = np.asarray(cached["idf"], dtype=np.float32)
        self.vocabulary = {str(t): i for i, t in enumerate(cached["vocabulary"])}
        self.doc_ids = [str(d) for d in cached["doc_ids"]]
        self.avgdl = float(cached["avgdl"])
        if "k1" in cached:
            self.k1, self.b = float(cached["k1"]), float(cached["b"])
        with self.cache_lock:
            self.query_cache.clear()
        self._gpu_index, self._gpu_params = None, None
        if meta_path != path and path.exists():
            ix = TermMajorIndex.load(path, verify=verify)
            params = (float(self.k1), float(self.b), float(self.avgdl))
            if (ix.kind == "bm25" and ix.n_docs == shape[0] and ix.n_vocab == shape[1] and
                    (ix.k1, ix.b, ix.avgdl) == params and np.array_equal(ix.idf_host, self.idf_weights)):
                self.tile_docs = ix.tile_docs
                self._gpu_index, self._gpu_params = ix, params
        self._sync_gpu_index()

    # ------------------------------------------------------------------ search
    def search_bm25(self, queries: Dict[str, str], top_k: int = 10) -> Dict[str, Dict[str, float]]:
        """retrieval.py:203-296, batched.  All queries that miss the cache are scored in one GPU pass."""
        if self.corpus_tf is None:
            raise ValueError("BM25 index not built. Call build_bm25_index() first.")
        results: Dict[str, Optional[Dict[str, float]]] = {}
        pending: Dict[str, List[str]] = {}          # cache key -> qids waiting for it
        packed: List[Tuple[np.ndarray, np.ndarray]] = []
        keys_in_order: List[str] = []
        for qid, text in queries.items():
            if not text or not text.strip():
                results[qid] = {}
                continue
            cache_key = f"{text.strip()}:{top_k}"
            with self.cache_lock:
                hit = self.query_cache.get(cache_key)
            if hit is not None:
                results[qid] = self._to_result(*hit)
                continue
            if cache_key in pending:
                results[qid] = None
                pending[cache_key].append(qid)
                continue
            counts = Counter(_WORD.findall(text.lower()))
            voc = self.vocabulary
            terms = sorted((voc[t], float(c)) for t, c in counts.items() if t in voc)   # unique ids, ascending
            if not terms:
                results[qid] = {}
                continue
            results[qid] = None
            pending[cache_key] = [qid]
            keys_in_order.append(cache_key)
            packed.append(terms)

        if packed:
            ix = self._sync_gpu_index()
            n_docs = len(self.doc_ids)
            k = int(top_k)
            # CSR-style packing of the batch (what pack_queries produces; the term lists are already unique, sorted and
            # positive, so the flat arrays are built in one go)
            q_ptr = np.zeros(len(packed) + 1, np.int32)
            np.cumsum([len(t) for t in packed], out=q_ptr[1:])
            q_terms = np.fromiter((t for q in packed for t, _ in q), np.int32, count=int(q_ptr[-1]))
            q_w = np.fromiter((c for q in packed for _, c in q), np.float32, count=int(q_ptr[-1]))
            if k < 1:
                top = [([], [])] * len(packed)
            elif k <= 1024:
                idx, val = ix.search_host(q_ptr, q_terms, q_w, min(k, n_docs))
                top = list(zip(idx, val))
            else:   # top_k > 1024: dense scores + the full-sort selector (rare)
                dense = ix.score_dense(q_ptr, q_terms, q_w)
                i2, v2 = fast_topk_selection(dense, min(k, n_docs))
                top = [(i2[i].cpu().numpy(), v2[i].cpu().numpy()) for i in range(len(packed))]
            for cache_key, (ti, tv) in zip(keys_in_order, top):
                with self.cache_lock:
                    if len(self.query_cache) < 1000:
                        self.query_cache[cache_key] = (ti, tv)
                res = self._to_result(ti, tv)
                qids = pending[cache_key]
                results[qids[0]] = res
                for qid in qids[1:]:
                    results[qid] = dict(res)
        return {qid: results[qid] for qid in queries}

    def _to_result(self, indices, scores) -> Dict[str, float]:
        ids = self.doc_ids
        il = indices.tolist() if hasattr(indices, "tolist") else list(indices)
        sl = scores.tolist() if hasattr(scores, "tolist") else list(scores)
        return {ids[i]: s for i, s in zip(il, sl) if s > 0 and i >= 0}

    # ------------------------------------------------------------------ dense side (off the BM25 path)
    def _load_embeddings(self):
        try:
            n = len(self.doc_ids)
            if n > 0:
                dim = self.embedding_path.stat().st_size // (n * 4)
                self.embedding_index = np.memmap(self.embedding_path, dtype="float32", mode="r", shape=(n, dim))
                self._gpu_embeddings = None
        except Exception as e:  # pragma: no cover
            self.logger.error("Error loading embeddings: %s", e)
            self.embedding_index = None

    def search_by_vector(self, query_vector: np.ndarray, k: int = 10, min_score: float = 0.0) -> List[Dict]:
        """retrieval.py:402-436: fp32 similarities + top-k + min_score cut-off.  The reference runs np.dot over
        the memmap through host BLAS; here the embedding matrix is uploaded to HBM once and b2r_f32_dot_topk
        does the gemv and the selection (scores agree with BLAS within f32 summation-order tolerance)."""
        if self.embedding_index is None:
            raise ValueError("No embedding index available")
        if self._gpu_embeddings is None:
            import torch
            self._gpu_embeddings = torch.from_numpy(np.array(self.embedding_index, dtype=np.float32, order="C")).cuda()
        idx, val = dense_topk(self._gpu_embeddings, np.asarray(query_vector, dtype=np.float32), k)
        out = []
        for i, s in zip(idx[0].cpu().numpy(), val[0].cpu().numpy()):
            if s < min_score:
                break
            if 0 <= i < len(self.doc_ids):
                out.append({"doc_id": self.doc_ids[int(i)], "score": float(s)})
        return out

    # ------------------------------------------------------------------ document fetch
    def get_document(self, doc_id: str) -> Optional[Document]:
        """retrieval.py:356-365."""
        with self.cache_lock:
            if doc_id in self._cache:
                return self._cache[doc_id]
        doc = self.index.get_document(doc_id)
        if doc:
            self._remember([doc])
        return doc

    def get_documents(self, doc_ids: List[str]) -> List[Optional[Document]]:
        """retrieval.py:367-400: cached documents are served from the cache, every other id of the call goes to the
        store in ONE batched fetch (docstore.MemoryIndex.get_documents: file-order pass + threaded inflation); result
        order = request order, None for an unknown id."""
        with self.cache_lock:
            have = {d: self._cache[d] for d in doc_ids if d in self._cache}
        missing = [d for d in dict.fromkeys(doc_ids) if d not in have]
        if missing:
            fetched = self.index.get_documents(missing, num_workers=self.num_workers)
            self._remember(fetched)
            have.update({doc.id: doc for doc in fetched if doc})
        return [have.get(d) for d in doc_ids]

    def _remember(self, docs) -> None:
        """retrieval.py:438-446 (_cache_documents): bounded cache, oldest entry evicted first."""
        with self.cache_lock:
            for d in docs:
                if d:
                    self._cache[d.id] = d
                    if len(self._cache) > self.cache_size:
                        self._cache.pop(next(iter(self._cache)))

    def get_search_results(self, query_results: List[Dict], include_text: bool = True) -> List[Dict]:
        """retrieval.py:436-462: one batched fetch for all results of the call."""
        out = []
        for doc, r in zip(self.get_documents([r["doc_id"] for r in query_results]), query_results):
            if doc:
                d = {"id": doc.id, "score": r["score"]}
                if include_text:
                    d.update({"text": doc.text, "title": doc.title, "metadata": doc.metadata})
                out.append(d)
        return out

    def fetch_results(self, results: Dict[str, Dict[str, float]], include_text: bool = True) -> Dict[str, List[Dict]]:
        """The fetch that follows a batched search: search_bm25's {qid: {doc_id: score}} for a whole query set ->
        {qid: [result dicts in rank order]} (the dict shape of get_search_results), with ONE store fetch for the union
        of all result lists."""
        ids = list(dict.fromkeys(d for r in results.values() for d in r))
        docs = dict(zip(ids, self.get_documents(ids)))
        out: Dict[str, List[Dict]] = {}
        for qid, r in results.items():
            rows = []
            for d, s in r.items():
                doc = docs.get(d)
                if doc:
                    row = {"id": doc.id, "score": s}
                    if include_text:
                        row.update({"text": doc.text, "title": doc.title, "metadata": doc.metadata})
                    rows.append(row)
            out[qid] = rows
        return out

    # ------------------------------------------------------------------ housekeeping
    def clear_cache(self) -> None:
        with self.cache_lock:
            self._cache.clear()
            self.query_cache.clear()

    def get_stats(self) -> Dict[str, object]:
        stats = {"cache_size": len(self._cache), "query_cache_size": len(self.query_cache),
                 "numba_available": NUMBA_AVAILABLE}
        if self.corpus_tf is not None:
            tf = self.corpus_tf
            stats.update({
                "num_docs": tf.shape[0],
                "vocab_size": len(self.vocabulary),
                "matrix_density": tf.nnz / (tf.shape[0] * tf.shape[1]),
                "bm25_memory_mb": (tf.data.nbytes + tf.indices.nbytes + tf.indptr.nbytes) / (1024 * 1024),
                "avgdl": self.avgdl,
            })
        return stats

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()

    def close(self):
        self.index.close()
        if self.embedding_index is not None and hasattr(self.embedding_index, "_mmap"):
            self.embedding_index._mmap.close()
        self.clear_cache()
