"""Registry plugin + pipeline hook.

Reference interfaces this file stands in for (paths relative to the reference root):

  rag_system/core/retriever_registry.py:120-355   OptimizedBM25Retriever(method, model, k1, b, **kwargs) with
                                                  build_index_from_corpus(corpus) / search(queries, top_k) / clear_cache()
  rag_system/core/retriever_registry.py:562-599   RetrieverRegistry.register(name, cls) / .create(config): a registered
                                                  class is instantiated as cls(**config["params"])
  rag_system/pipeline/evaluate_rag_pipeline.py:162-180, 682-689   the pipeline's own factory builds
                                                  OptimizedRetriever(config, hardware_info)
  rag_system/pipeline/evaluate_rag_pipeline.py:181-207, 280-312   its index cache .rag_cache/<method>_index_<hash>.npz
  rag_system/pipeline/evaluate_rag_pipeline.py:741-780            run_rag_experiment feeds retriever.search batches of
                                                  <= 100 queries

B200BM25Retriever keeps those contracts and moves the work: the index lives term-major in HBM and search() scores every
query of the call in ONE GPU pass.  prefetch(queries, top_k) is the pipeline hook: called once with the WHOLE query set
of an experiment it runs a single batched search, after which the pipeline's per-batch search() calls are answered
from its results (same dicts the per-batch calls would have produced)."""
from __future__ import annotations

import hashlib
import os
import tempfile
from pathlib import Path
from typing import Any, Dict, Optional

import numpy as np

from .docstore import MemoryIndex
from .service import RetrievalService


class B200BM25Retriever:
    """BM25 on the B200 behind the reference's retriever interface.  `method="tfidf"` maps to k1=1000, b=0 exactly
    like RetrieverRegistry.create (retriever_registry.py:593-595)."""

    def __init__(self, method: str = "bm25", model: str = None, k1: float = 1.2, b: float = 0.75, **kwargs):
        self.method = method.lower()
        self.model_name = model
        if self.method == "tfidf":
            k1, b = 1000.0, 0.0
        # retriever_registry.py:130-133: accepted and kept; caching of queries follows the reference's flag
        self.use_cache = kwargs.get("cache_matrices", True)
        self.cache_queries = kwargs.get("cache_queries", True)
        self.cache_dir = kwargs.get("cache_dir")          # e.g. ".rag_cache": share the pipeline's index cache files
        self._tmp = tempfile.TemporaryDirectory(prefix="b200ret_")
        path = os.path.join(self._tmp.name, "docs.idx")
        MemoryIndex(path, create=True).close()
        self.service = RetrievalService(path)
        self.service.k1, self.service.b = float(k1), float(b)
        self._prefetched: Dict[tuple, Dict[str, float]] = {}

    # ------------------------------------------------------------------ the pipeline's constructor contract
    @classmethod
    def from_pipeline_config(cls, config: Dict[str, Any], hardware_info: Optional[Dict[str, Any]] = None):
        """evaluate_rag_pipeline.py:165-180: OptimizedRetriever(config, hardware_info)."""
        params = dict(config.get("params", {}) or {})
        k1, b = params.pop("k1", 1.2), params.pop("b", 0.75)
        params.pop("top_k", None)                          # a search-time parameter in the pipeline's config
        self = cls(method=config.get("type", "bm25"), model=config.get("model"), k1=k1, b=b,
                   cache_dir=params.pop("cache_dir", ".rag_cache"), **params)
        self.config, self.hardware = config, hardware_info
        return self

    # ------------------------------------------------------------------ attribute surface of the reference class
    corpus_tf = property(lambda self: self.service.corpus_tf)
    vocabulary = property(lambda self: self.service.vocabulary)
    idf_weights = property(lambda self: self.service.idf_weights)
    idf = idf_weights                                       # the pipeline's retriever calls it `idf`
    doc_lengths = property(lambda self: self.service.doc_lengths)
    doc_ids = property(lambda self: self.service.doc_ids)
    avgdl = property(lambda self: self.service.avgdl)
    k1 = property(lambda self: self.service.k1)
    b = property(lambda self: self.service.b)
    query_cache = property(lambda self: self.service.query_cache)

    # ------------------------------------------------------------------ build
    def cache_file_for(self, corpus: Dict[str, Dict]) -> Optional[Path]:
        """evaluate_rag_pipeline.py:189-192: .rag_cache/<method>_index_<md5 of the first 1000 sorted ids>.npz."""
        if not self.cache_dir:
            return None
        corpus_hash = hashlib.md5(str(sorted(corpus.keys())[:1000]).encode()).hexdigest()[:8]
        return Path(self.cache_dir) / f"{self.method}_index_{corpus_hash}.npz"

    def build_index_from_corpus(self, corpus: Dict[str, Dict]) -> None:
        """Build (or, with cache_dir set, load the pipeline's cached CSR -- a file written by the reference itself is
        accepted -- and lay it out in HBM).  A freshly built index is cached under the reference's own keys
        (evaluate_rag_pipeline.py:280-293), so the reference can read it back."""
        if not corpus:
            raise ValueError("Empty corpus provided")
        self._prefetched.clear()
        cache_file = self.cache_file_for(corpus) if self.use_cache else None
        if cache_file is not None and cache_file.exists():
            self.load_cached_index(cache_file)
            return
        self.service.build_bm25_index(corpus)
        if cache_file is not None:
            try:
                cache_file.parent.mkdir(exist_ok=True)
                self.save_cached_index(cache_file)
            except OSError:
                pass        # the reference also carries on without its cache (evaluate_rag_pipeline.py:294-295)

    def load_cached_index(self, cache_file) -> None:
        """evaluate_rag_pipeline.py:297-312 (_load_cached_index): a reference-style .npz -> HBM."""
        self.service.load_bm25_index(Path(cache_file))
        self._prefetched.clear()

    def save_cached_index(self, cache_file) -> None:
        """evaluate_rag_pipeline.py:277-293 (_save_cached_index), same keys, no pickled objects."""
        s, tf = self.service, self.service.corpus_tf
        with open(cache_file, "wb") as f:
            np.savez_compressed(f, tf_data=tf.data, tf_indices=tf.indices, tf_indptr=tf.indptr,
                                tf_shape=np.asarray(tf.shape), doc_lengths=s.doc_lengths, idf=s.idf_weights,
                                vocabulary=np.asarray(sorted(s.vocabulary, key=s.vocabulary.get), dtype=np.str_),
                                doc_ids=np.asarray([str(d) for d in s.doc_ids], dtype=np.str_), avgdl=s.avgdl)

    # ------------------------------------------------------------------ search
    def prefetch(self, queries: Dict[str, Any], top_k: int = 10) -> int:
        """Pipeline hook: score the WHOLE query set in one batched GPU call (values may be strings or the pipeline's
        query objects, evaluate_rag_pipeline.py:752-766) and keep the results; the per-batch search() calls that follow
        are served from them.  Returns the number of distinct query texts scored."""
        texts: Dict[str, str] = {}
        for qid, qobj in queries.items():
            text = self._query_text(qobj)
            if text and text.strip():
                texts.setdefault(text.strip(), text.strip())
        if not texts:
            return 0
        res = self.service.search_bm25({t: t for t in texts}, top_k=top_k)
        for t, r in res.items():
            self._prefetched[(t, int(top_k))] = r
        return len(texts)

    @staticmethod
    def _query_text(qobj) -> str:
        if isinstance(qobj, str):
            return qobj
        if isinstance(qobj, dict):
            return (qobj.get("text") or qobj.get("query") or qobj.get("title") or qobj.get("question") or
                    qobj.get("body") or str(qobj.get("id", "")))
        return str(qobj) if qobj else ""

    def search(self, queries: Dict[str, str], top_k: int = 10) -> Dict[str, Dict[str, float]]:
        """retriever_registry.py:228-262: one result dict per query id, rank order, only scores > 0; every query of
        the call that is neither prefetched nor cached is scored in the same GPU pass."""
        if self.service.corpus_tf is None:
            raise ValueError("Index not built. Call build_index_from_corpus() first.")
        out: Dict[str, Optional[Dict[str, float]]] = {}
        todo: Dict[str, str] = {}
        for qid, text in queries.items():
            if not text:
                out[qid] = {}
                continue
            hit = self._prefetched.get((text.strip(), int(top_k))) if self._prefetched else None
            if hit is not None:
                out[qid] = dict(hit)
            else:
                out[qid] = None
                todo[qid] = text
        if todo:
            out.update(self.service.search_bm25(todo, top_k=top_k))
        return {qid: out[qid] for qid in queries}

    def clear_cache(self) -> None:
        self._prefetched.clear()
        with self.service.cache_lock:
            self.service.query_cache.clear()

    def get_stats(self):
        return self.service.get_stats()


def register_with(registry, name: str = "bm25_b200") -> None:
    """registry.register(name, cls) -- retriever_registry.py:567-569."""
    registry.register(name, B200BM25Retriever)


def install(name: str = "bm25_b200", take_over_bm25: bool = False):
    """Register into the reference's own RetrieverRegistry when the reference package is importable
    (rag_system.core.retriever_registry).  take_over_bm25=True additionally routes the built-in names
    ('bm25', 'bm25_retriever', 'bm25_custom', 'tfidf') to this class, which is what a maintainer switching the
    hot path over would do.  Returns the registry class."""
    from rag_system.core import retriever_registry as rr      # the reference (raises ImportError when absent)
    rr.RetrieverRegistry.register(name, B200BM25Retriever)
    if take_over_bm25:
        stock_create = rr.RetrieverRegistry.create.__func__

        def create(cls, config):
            cfg = {"type": config} if isinstance(config, str) else dict(config)
            method = str(cfg.get("type", cfg.get("name", ""))).lower()
            if method in ("bm25", "bm25_retriever", "bm25_custom", "tfidf"):
                return B200BM25Retriever(method=method, model=cfg.get("model"), **cfg.get("params", {}))
            return stock_create(cls, config)
        rr.RetrieverRegistry.create = classmethod(create)
    return rr.RetrieverRegistry
