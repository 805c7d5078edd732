"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is mounted:

    python oracle/gen_golden.py            # writes tests/golden/*.npz, *.json

The fixtures hold the inputs and the reference's own outputs for every kernel on the hot path
(simd_bm25_score, fast_topk_selection, simd_tfidf_score, quantized_dot_product_batch) and for
RetrievalService.build_bm25_index/search_bm25 on a small text corpus.  They are what pins the
oracle (oracle/np_oracle.py, oracle/bm25_oracle.c) and, through it, the CUDA path.  The GPU box
has no /root/reference, so nothing at test/bench time imports it; only this script does.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("B2R_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import np_oracle  # noqa: E402


def zipf_csr(rng, n_docs, n_vocab, mean_len, empty_every=0, weights=False):
    """Zipf(s=1) tokens, gamma doc lengths; returns CSR with sorted rows + doc_lengths (f32)."""
    p = 1.0 / np.arange(1, n_vocab + 1)
    cdf = np.cumsum(p / p.sum())
    lens = np.clip(np.floor(rng.gamma(2.0, mean_len / 2.0, n_docs)), 3, 8 * mean_len).astype(np.int64)
    if empty_every:
        lens[::empty_every] = 0
    toks = np.minimum(np.searchsorted(cdf, rng.random(int(lens.sum()))), n_vocab - 1)
    doc = np.repeat(np.arange(n_docs), lens)
    key = doc * n_vocab + toks
    uniq, cnt = np.unique(key, return_counts=True)
    rows = uniq // n_vocab
    indices = (uniq % n_vocab).astype(np.int32)
    indptr = np.zeros(n_docs + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n_docs), out=indptr[1:])
    data = cnt.astype(np.float32)
    if weights:
        data = rng.gamma(2.0, 0.5, len(data)).astype(np.float32)
    return data, indices, indptr, lens.astype(np.float32)


def make_queries(rng, n_q, n_vocab, head, lo=1, hi=8, weighted=False):
    p = 1.0 / np.arange(1, head + 1)
    p /= p.sum()
    ptr = [0]
    terms, w = [], []
    for _ in range(n_q):
        nt = int(rng.integers(lo, hi + 1))
        t = np.unique(rng.choice(head, size=nt, p=p))
        terms.extend(t.tolist())
        if weighted:
            w.extend(rng.gamma(2.0, 0.5, len(t)).astype(np.float32).tolist())
        else:
            w.extend(rng.integers(1, 4, len(t)).astype(np.float32).tolist())
        ptr.append(len(terms))
    return np.asarray(ptr, np.int32), np.asarray(terms, np.int32), np.asarray(w, np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    from rag_system.core import retrieval as ref_ret
    from rag_system.core import retriever_registry as ref_reg
    from rag_system.pipeline import evaluate_rag_pipeline as ref_pipe
    import numba, scipy

    versions = dict(numpy=np.__version__, numba=numba.__version__, scipy=scipy.__version__,
                    threads=int(numba.get_num_threads()))

    # ---------------------------------------------------------------- K1 + K2 on arrays
    rng = np.random.default_rng(20260101)
    n_docs, n_vocab = 2500, 1200
    data, indices, indptr, dl = zipf_csr(rng, n_docs, n_vocab, 30, empty_every=97)
    idf = np_oracle.idf_from_csr(indices, n_docs, n_vocab)
    avgdl = np_oracle.avgdl_from_lengths(dl)
    assert (idf < 0).sum() >= 2, "fixture must contain negative-idf head terms"
    q_ptr, q_terms, q_w = make_queries(rng, 16, n_vocab, n_vocab // 4)
    k1, b = 1.2, 0.75
    scores = np.zeros((len(q_ptr) - 1, n_docs), np.float32)
    top_idx = np.zeros((len(q_ptr) - 1, 10), np.int64)
    top_val = np.zeros((len(q_ptr) - 1, 10), np.float32)
    indptr32 = indptr.astype(np.int32)
    for q in range(len(q_ptr) - 1):
        qtf = np_oracle.dense_query(q_terms[q_ptr[q]:q_ptr[q + 1]], q_w[q_ptr[q]:q_ptr[q + 1]], n_vocab)
        scores[q] = ref_ret.simd_bm25_score(qtf, data, indices, indptr32, dl, idf, k1, b, avgdl)
        i, v = ref_ret.fast_topk_selection(scores[q], 10)
        top_idx[q], top_val[q] = i, v
    # a second parameterisation (the registry maps "tfidf" to k1=1000, b=0: retriever_registry.py:593-595)
    scores_k1000 = np.zeros((4, n_docs), np.float32)
    for q in range(4):
        qtf = np_oracle.dense_query(q_terms[q_ptr[q]:q_ptr[q + 1]], q_w[q_ptr[q]:q_ptr[q + 1]], n_vocab)
        scores_k1000[q] = ref_reg.simd_bm25_batch_score(qtf, data, indices, indptr32, dl, idf, 1000.0, 0.0, avgdl)
    np.savez_compressed(os.path.join(OUT, "bm25_arrays.npz"), data=data, indices=indices, indptr=indptr,
                        doc_lengths=dl, idf=idf, avgdl=np.float64(avgdl), k1=k1, b=b, q_ptr=q_ptr,
                        q_terms=q_terms, q_weights=q_w, ref_scores=scores, ref_top_idx=top_idx,
                        ref_top_val=top_val, ref_scores_k1000=scores_k1000)

    # fractional tf / doc lengths / weights / idf and other k1, b: pins the f64 evaluation order
    rng = np.random.default_rng(7)
    nf, vf = 1200, 500
    dataf, indf, ptrf, dlf = zipf_csr(rng, nf, vf, 25, weights=True)
    dlf = (dlf + rng.random(nf).astype(np.float32)).astype(np.float32)
    idff = (np_oracle.idf_from_csr(indf, nf, vf) * np.float32(1.2345)).astype(np.float32)
    avgdlf = np_oracle.avgdl_from_lengths(dlf)
    qpf, qtf_, qwf = make_queries(rng, 8, vf, vf // 2, lo=2, hi=12, weighted=True)
    sf = np.zeros((8, nf), np.float32)
    for q in range(8):
        qv = np_oracle.dense_query(qtf_[qpf[q]:qpf[q + 1]], qwf[qpf[q]:qpf[q + 1]], vf)
        sf[q] = ref_ret.simd_bm25_score(qv, dataf, indf, ptrf.astype(np.int32), dlf, idff, 0.9, 0.4, avgdlf)
    np.savez_compressed(os.path.join(OUT, "bm25_frac.npz"), data=dataf, indices=indf, indptr=ptrf,
                        doc_lengths=dlf, idf=idff, avgdl=np.float64(avgdlf), k1=0.9, b=0.4, q_ptr=qpf,
                        q_terms=qtf_, q_weights=qwf, ref_scores=sf)

    # ---------------------------------------------------------------- K3
    rng = np.random.default_rng(20260103)
    n_docs3, n_vocab3 = 1500, 900
    data3, indices3, indptr3, _ = zipf_csr(rng, n_docs3, n_vocab3, 40, weights=True)
    idf3 = np_oracle.idf_from_csr(indices3, n_docs3, n_vocab3)
    q_ptr3, q_terms3, q_w3 = make_queries(rng, 12, n_vocab3, n_vocab3 // 2, lo=5, hi=30, weighted=True)
    s3 = np.zeros((len(q_ptr3) - 1, n_docs3), np.float32)
    s3_ones = np.zeros_like(s3)
    ones = np.ones(n_vocab3, np.float32)
    for q in range(len(q_ptr3) - 1):
        qtf = np_oracle.dense_query(q_terms3[q_ptr3[q]:q_ptr3[q + 1]], q_w3[q_ptr3[q]:q_ptr3[q + 1]], n_vocab3)
        s3[q] = ref_pipe.simd_tfidf_score(qtf, data3, indices3, indptr3.astype(np.int32), idf3)
        s3_ones[q] = ref_pipe.simd_tfidf_score(qtf, data3, indices3, indptr3.astype(np.int32), ones)
    np.savez_compressed(os.path.join(OUT, "tfidf_arrays.npz"), data=data3, indices=indices3, indptr=indptr3,
                        idf=idf3, q_ptr=q_ptr3, q_terms=q_terms3, q_weights=q_w3, ref_scores=s3,
                        ref_scores_idf1=s3_ones)

    # ---------------------------------------------------------------- K4
    rng = np.random.default_rng(42)
    dim, n4, nq4 = 768, 300, 8
    centers = rng.normal(0, 0.5, (10, dim))
    emb = centers[rng.integers(0, 10, n4)] + rng.normal(0, 0.2, (n4, dim))
    emb = (emb / np.linalg.norm(emb, axis=1, keepdims=True)).astype(np.float32)
    qemb = centers[rng.integers(0, 10, nq4)] + rng.normal(0, 0.2, (nq4, dim))
    qemb = (qemb / np.linalg.norm(qemb, axis=1, keepdims=True)).astype(np.float32)
    retr = ref_reg.QuantizedEmbeddingRetriever.__new__(ref_reg.QuantizedEmbeddingRetriever)
    retr.quantization_method = 'symmetric'
    d8, dscale = ref_reg.QuantizedEmbeddingRetriever._quantize_embeddings(retr, emb)
    dscale = np.asarray(dscale, np.float32).reshape(-1)
    # query side of QuantizedEmbeddingRetriever.search (retriever_registry.py:482-485): scale = max|x| / 127
    qscale = (np.max(np.abs(qemb), axis=1) / 127.0).astype(np.float32)
    q8 = np.clip(np.round(qemb / qscale[:, None]), -127, 127).astype(np.int8)
    sims = ref_reg.quantized_dot_product_batch(q8, np.ascontiguousarray(d8), qscale, dscale)
    np.savez_compressed(os.path.join(OUT, "int8.npz"), emb=emb[:40], q8=q8, d8=d8, qscale=qscale, dscale=dscale,
                        ref_sims=sims)

    # ---------------------------------------------------------------- K2 distributions (tests/topk_selection.py:274-307)
    rng = np.random.default_rng(42)
    cases = {}
    for name, n, k, arr in [
        ("normal", 100, 10, rng.normal(0, 1, 100)),
        ("uniform", 1000, 50, rng.uniform(0, 1, 1000)),
        ("zipfian", 500, 5, 1.0 / rng.permutation(np.arange(1, 501))),
        ("bimodal", 200, 100, np.concatenate([rng.normal(-2, 0.5, 100), rng.normal(2, 0.5, 100)])),
        ("k_ge_n", 37, 64, rng.normal(0, 1, 37)),
        ("big", 20000, 100, rng.normal(0, 1, 20000)),
    ]:
        arr = arr.astype(np.float32)
        i, v = ref_ret.fast_topk_selection(arr, k)
        cases[f"{name}_scores"] = arr
        cases[f"{name}_k"] = np.int64(k)
        cases[f"{name}_ref_idx"] = np.asarray(i, np.int64)
        cases[f"{name}_ref_val"] = np.asarray(v, np.float32)
    np.savez_compressed(os.path.join(OUT, "topk_cases.npz"), **cases)

    # ---------------------------------------------------------------- service level (text API)
    sys.path.insert(0, os.path.join(REF, "tests"))
    from core_test import SyntheticDataGenerator
    from rag_system.core.memory_index import MemoryIndex
    corpus = {}
    for d in SyntheticDataGenerator.generate_corpus(400, avg_doc_length=40, vocab_size=600, seed=42):
        corpus[d["_id"]] = {"text": d["text"], "title": d["title"]}
    corpus["doc_empty"] = {"text": ""}
    corpus["doc_content"] = {"content": "word_1 word_2 fallback Field CONTENT word_2"}
    corpus["doc_body"] = {"body": "word_3 body-field, punctuation! word_3 word_3"}
    queries = {}
    for q in SyntheticDataGenerator.generate_queries(40, avg_query_length=5, vocab_size=600, seed=43):
        queries[q["qid"]] = q["text"]
    queries["blank"] = "   "
    queries["oov"] = "zzzz qqqq"
    queries["repeat"] = "word_3 word_3 word_3 fallback"
    queries["punct"] = "Body-field, WORD_1!"
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "docs.idx")
        MemoryIndex(path, create=True)
        svc = ref_ret.RetrievalService(path)
        svc.build_bm25_index(corpus)
        res10 = svc.search_bm25(queries, top_k=10)
        svc.clear_cache()
        res500 = svc.search_bm25({k: queries[k] for k in list(queries)[:6]}, top_k=500)   # top_k > n_docs branch
        meta = dict(vocab_size=len(svc.vocabulary), avgdl=svc.avgdl, n_docs=len(svc.doc_ids),
                    nnz=int(svc.corpus_tf.nnz), idf_sum=float(np.sum(svc.idf_weights.astype(np.float64))),
                    vocab_head=sorted(svc.vocabulary, key=svc.vocabulary.get)[:8],
                    stats_keys=sorted(svc.get_stats().keys()))
        svc.close()
    with open(os.path.join(OUT, "service_text.json"), "w") as f:
        json.dump(dict(corpus=corpus, queries=queries, ref_top10=res10, ref_top500=res500, meta=meta,
                       versions=versions), f)
    fiqa_shape(ref_ret, versions)
    print("golden fixtures written to", OUT, versions)


def fiqa_shape(ref_ret, versions):
    """BASELINE config 1: the reference's RetrievalService on the FiQA-shape synthetic text corpus of its own
    generator (tests/core_test.py:203-252; 57,638 docs, 648 queries, top-10).  The corpus (68 MB of text) is not
    committed: the fixture pins it by sha256 and the product-side generator (b200ret/synthetic.py, loaded here
    without the CUDA library) is checked to reproduce the reference generator's output exactly.  Stored per query:
    what search_bm25 returned, and the canonical top-10 (score desc, doc index asc) of the reference's OWN score
    vector (simd_bm25_score called exactly as _score_bm25_query does), which is what a deterministic tie rule must
    reproduce."""
    import importlib.util
    import re as _re
    from collections import Counter
    sys.path.insert(0, os.path.join(REF, "tests"))
    from core_test import SyntheticDataGenerator
    from rag_system.core.memory_index import MemoryIndex
    spec = importlib.util.spec_from_file_location(
        "b2r_synthetic", os.path.join(os.path.dirname(HERE), "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200",
                                      "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    cfg = syn.FIQA_SHAPE
    corpus = {}
    for d in SyntheticDataGenerator.generate_corpus(cfg["num_docs"], cfg["avg_doc_length"], cfg["vocab_size"],
                                                    seed=cfg["corpus_seed"]):
        corpus[d["_id"]] = {"text": d["text"], "title": d["title"]}
    queries = {q["qid"]: q["text"] for q in SyntheticDataGenerator.generate_queries(
        cfg["num_queries"], cfg["avg_query_length"], cfg["vocab_size"], seed=cfg["query_seed"])}
    sha = syn.corpus_sha256(corpus)
    mine = syn.fiqa_shape_corpus()
    assert syn.corpus_sha256(mine) == sha and list(mine) == list(corpus), "restated corpus generator diverges"
    assert syn.fiqa_shape_queries() == queries, "restated query generator diverges"
    del mine
    k = 10
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "docs.idx")
        MemoryIndex(path, create=True)
        svc = ref_ret.RetrievalService(path)
        svc.build_bm25_index(corpus)
        res = svc.search_bm25(queries, top_k=k)
        nq = len(queries)
        ref_idx = np.full((nq, k), -1, np.int64)
        ref_val = np.zeros((nq, k), np.float32)
        canon_idx = np.full((nq, k), -1, np.int64)
        canon_val = np.zeros((nq, k), np.float32)
        pos = {d: i for i, d in enumerate(svc.doc_ids)}
        for qi, (qid, text) in enumerate(queries.items()):
            for j, (doc_id, sc) in enumerate(res[qid].items()):
                ref_idx[qi, j], ref_val[qi, j] = pos[doc_id], np.float32(sc)
            counts = Counter(_re.findall(r"\b\w+\b", text.lower()))
            qtf = np.zeros(len(svc.vocabulary), np.float32)
            hit = 0
            for t, c in counts.items():
                if t in svc.vocabulary:
                    qtf[svc.vocabulary[t]] = float(c)
                    hit += 1
            if not hit:
                continue
            s = ref_ret.simd_bm25_score(qtf, svc.corpus_tf.data, svc.corpus_tf.indices, svc.corpus_tf.indptr,
                                        svc.doc_lengths, svc.idf_weights, svc.k1, svc.b, svc.avgdl)
            ci, cv = np_oracle.topk_canonical(s, k)
            canon_idx[qi, :len(ci)], canon_val[qi, :len(ci)] = ci, cv
        meta = dict(corpus_sha256=sha, n_docs=len(svc.doc_ids), vocab_size=len(svc.vocabulary),
                    nnz=int(svc.corpus_tf.nnz), avgdl=svc.avgdl,
                    idf_sum=float(np.sum(svc.idf_weights.astype(np.float64))),
                    n_empty_results=int(sum(1 for r in res.values() if not r)), versions=versions, config=cfg)
        svc.close()
    np.savez_compressed(os.path.join(OUT, "fiqa_shape.npz"), ref_idx=ref_idx, ref_val=ref_val, canon_idx=canon_idx,
                        canon_val=canon_val, meta=np.asarray(json.dumps(meta)))
    print("fiqa-shape fixture:", meta)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fiqa":      # only the (slow) config-1 fixture
        import numba, scipy
        from rag_system.core import retrieval as _ref_ret
        fiqa_shape(_ref_ret, dict(numpy=np.__version__, numba=numba.__version__, scipy=scipy.__version__,
                                  threads=int(numba.get_num_threads())))
    else:
        main()
