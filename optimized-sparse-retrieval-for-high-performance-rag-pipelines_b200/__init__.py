"""b200ret -- B200-native BM25 / top-k retrieval scoring behind the reference's API.

Drop-in surface (reference: nytdevansh/Optimized-Sparse-Retrieval-for-High-Performance-RAG-Pipelines):
RetrievalService.build_bm25_index / search_bm25, simd_bm25_score, fast_topk_selection,
simd_tfidf_score, quantized_dot_product_batch and the README aliases optimized_bm25_score /
fast_topk.  All scoring and selection runs in libb200ret.so (hand-written sm_100a CUDA) through a
C ABI; importing this package fails loudly if that library is missing.
"""
from . import _abi                                              # noqa: F401  (loads libb200ret.so or raises)
from .docstore import Document, MemoryIndex                     # noqa: F401
from .index import (TermMajorIndex, pack_queries, queries_from_dense, reference_avgdl, reference_idf,  # noqa: F401
                    set_approx_prefilter, set_fused_cap, set_fused_selection)
from .kernels import (clear_index_cache, dense_topk, set_bank_schedule, set_int8_cluster, set_int8_pair, set_int8_mma, set_int8_fused, fast_topk, fast_topk_selection, hybrid_search, int8_rerank, int8_scan_topk,            # noqa: F401
                      optimized_bm25_score, quantized_dot_product_batch, simd_bm25_batch_score,
                      simd_bm25_score, simd_tfidf_score)
from .retriever import B200BM25Retriever, register_with         # noqa: F401
from .service import RetrievalService                           # noqa: F401

__version__ = "0.1.0"
