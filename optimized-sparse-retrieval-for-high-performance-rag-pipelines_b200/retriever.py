"""Registry plugin shape (reference: rag_system/core/retriever_registry.py:120-355, 562-625):
an object with build_index_from_corpus(corpus) and search(queries, top_k), creatable through
RetrieverRegistry.register(name, cls) / .create(config)."""
from __future__ import annotations

import os
import tempfile
from typing import Dict

from .docstore import MemoryIndex
from .service import RetrievalService


class B200BM25Retriever:
    """BM25 on the B200 behind the reference's retriever interface.  `method="tfidf"` maps to
    k1=1000, b=0 exactly like RetrieverRegistry.create (retriever_registry.py:593-595)."""

    def __init__(self, method: str = "bm25", model: str = None, k1: float = 1.2, b: float = 0.75, **kwargs):
        self.method = method.lower()
        self.model_name = model
        if self.method == "tfidf":
            k1, b = 1000.0, 0.0
        self._tmp = tempfile.TemporaryDirectory(prefix="b200ret_")
        path = os.path.join(self._tmp.name, "docs.idx")
        MemoryIndex(path, create=True).close()
        self.service = RetrievalService(path)
        self.service.k1, self.service.b = float(k1), float(b)

    def build_index_from_corpus(self, corpus: Dict[str, Dict]) -> None:
        self.service.build_bm25_index(corpus)

    def search(self, queries: Dict[str, str], top_k: int = 10) -> Dict[str, Dict[str, float]]:
        return self.service.search_bm25(queries, top_k=top_k)

    def get_stats(self):
        return self.service.get_stats()


def register_with(registry, name: str = "bm25_b200") -> None:
    """registry.register(name, cls) -- retriever_registry.py:567-569."""
    registry.register(name, B200BM25Retriever)
