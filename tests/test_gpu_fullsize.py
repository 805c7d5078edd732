"""Parity at the sizes BASELINE.json names, against the CPU oracle.

Every test here runs the CUDA path through the C ABI on the full-size configuration and compares ids AND
scores bit for bit with the oracle (oracle/bm25_oracle.c, pinned to the reference by tests/golden/) on a
query sample the oracle finishes in seconds, plus size-independent properties on all queries.
Corpora of 8.8M documents are generated on the GPU (tools/gpu_synth.py: data generation only) and copied
to the host for the oracle."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import c_oracle, np_oracle  # noqa: E402  (the checker)


def shard_range(n, world, rank):
    from b200ret.dist import shard_range as f
    return f(n, world, rank)


@pytest.fixture(scope="module")
def b2r():
    import b200ret
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return b200ret


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _fold(v):
    """-0.0 -> +0.0: the ranking key does not distinguish them and decode returns +0.0."""
    return np.where(v == 0, np.float32(0), v)


def _merge(b2r, parts, nq, k):
    g = torch.stack(parts).contiguous()
    mi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    mv = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ws = torch.empty(nq * k * 8 + (1 << 22), dtype=torch.uint8, device="cuda")
    b2r._abi.check(b2r._abi.lib.b2r_merge_candidates(g.data_ptr(), len(parts), nq, k, None, mi.data_ptr(), mv.data_ptr(),
                                                     ws.data_ptr(), ws.numel(),
                                                     int(torch.cuda.current_stream().cuda_stream)))
    return mi, mv


# ----------------------------------------------------------------------------------- C1
def test_config1_fiqa_shape_search_bm25_vs_reference(b2r, golden_dir, tmp_path):
    """BASELINE config 1: 57,638-document FiQA-shape text corpus of the reference's own generator, 648 queries,
    top-10 through RetrievalService.build_bm25_index / search_bm25.  Expected = the canonical top-10 (score desc,
    doc index asc) of the REFERENCE's own score vectors, recorded by oracle/gen_golden.py (fiqa_shape.npz)."""
    from b200ret import synthetic as S
    z = np.load(os.path.join(golden_dir, "fiqa_shape.npz"))
    meta = json.loads(str(z["meta"]))
    corpus = S.fiqa_shape_corpus()
    assert S.corpus_sha256(corpus) == meta["corpus_sha256"]
    queries = S.fiqa_shape_queries()
    path = tmp_path / "docs.idx"
    b2r.MemoryIndex(path, create=True).close()
    with b2r.RetrievalService(path) as svc:
        svc.build_bm25_index(corpus)
        assert len(svc.vocabulary) == meta["vocab_size"] and svc.corpus_tf.nnz == meta["nnz"]
        assert svc.avgdl == meta["avgdl"]
        assert float(np.sum(svc.idf_weights.astype(np.float64))) == meta["idf_sum"]
        got = svc.search_bm25(queries, top_k=10)
        assert list(got) == list(queries)
        n_empty = 0
        for qi, qid in enumerate(queries):
            keep = z["canon_val"][qi] > 0
            want_ids = [f"doc_{i}" for i in z["canon_idx"][qi][keep]]
            want_val = [float(v) for v in z["canon_val"][qi][keep]]
            assert list(got[qid]) == want_ids, qid
            assert list(got[qid].values()) == want_val, qid
            # and the reference's own return value: same scores, in order
            n_ref = int((z["ref_idx"][qi] >= 0).sum())
            assert [float(v) for v in z["ref_val"][qi, :n_ref]] == want_val, qid
            n_empty += not got[qid]
        assert n_empty == meta["n_empty_results"] == 22
        assert svc.search_bm25(queries, top_k=10) == got            # served from the query cache


# ----------------------------------------------------------------------------------- C2
def test_config2_all_1024_queries_bit_exact(b2r):
    """BASELINE config 2 (1M docs x 100K vocab, 1024 queries of 4-8 terms, top-10): EVERY query against the oracle,
    single index and 2-/3-shard merge."""
    from b200ret import synthetic as S
    n_docs, n_vocab, k, nq = 1_000_000, 100_000, 10, 1024
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 60)
    idf = b2r.reference_idf(indices, n_docs, n_vocab)
    avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(nq, n_vocab)
    c_oracle.use_all_host_threads()
    wi, wv = c_oracle.bm25_search_batch(q_ptr, q_terms, q_w, n_vocab, data, indices, indptr, dl, idf, 1.2, 0.75, avgdl, k)
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    idx, val = ix.search(q_ptr, q_terms, q_w, k)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(_bits(val.cpu().numpy()), _bits(_fold(wv)))
    hi_, hv_ = ix.search_host(q_ptr, q_terms, q_w, k)              # the host-buffer C-ABI call
    assert np.array_equal(hi_, wi) and np.array_equal(_bits(hv_), _bits(_fold(wv)))
    del ix
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard_range(n_docs, world, r)
            s, e = indptr[lo], indptr[hi]
            sh = b2r.TermMajorIndex.from_csr(data[s:e], indices[s:e], indptr[lo:hi + 1] - s, dl[lo:hi], n_vocab=n_vocab,
                                             idf=idf, avgdl=avgdl, doc_id_base=lo)
            parts.append(sh.search(q_ptr, q_terms, q_w, k, return_keys=True)[2])
            del sh
        mi, mv = _merge(b2r, parts, nq, k)
        assert np.array_equal(mi.cpu().numpy(), wi) and np.array_equal(_bits(mv.cpu().numpy()), _bits(_fold(wv))), world


# ----------------------------------------------------------------------------------- C3
def test_config3_8p8m_docs_top100_vs_oracle_and_sharded(b2r):
    """BASELINE config 3 shape (8.8M docs x 100K vocab, top-100): 64 queries against the oracle on one index and
    through the 2- and 3-shard merge; all 1024 queries: sharded == unsharded, keys strictly descending."""
    from b200ret import synthetic as S
    from gpu_synth import global_bm25_stats, zipf_csr_torch
    dev = torch.device("cuda")
    n_docs, n_vocab, k, nq, n_check = 8_800_000, 100_000, 100, 1024, 64
    data, ind, ptr, dl = zipf_csr_torch(n_docs, n_vocab, 60.0, 20260101, dev)
    _, idf, avgdl = global_bm25_stats(ind, dl, n_docs, n_vocab)
    q_ptr, q_terms, q_w = S.zipf_queries(nq, n_vocab)
    ix = b2r.TermMajorIndex.from_csr(data, ind, ptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    idx, val, keys = ix.search(q_ptr, q_terms, q_w, k, return_keys=True)
    torch.cuda.synchronize()
    ku = keys.cpu().numpy().view(np.uint64)
    assert bool((ku[:, :-1] > ku[:, 1:]).all())
    h_data, h_ind, h_ptr, h_dl = (t.cpu().numpy() for t in (data, ind, ptr, dl))
    c_oracle.use_all_host_threads()
    e = int(q_ptr[n_check])
    wi, wv = c_oracle.bm25_search_batch(q_ptr[:n_check + 1], q_terms[:e], q_w[:e], n_vocab, h_data, h_ind, h_ptr, h_dl,
                                        idf, 1.2, 0.75, avgdl, k)
    assert np.array_equal(idx[:n_check].cpu().numpy(), wi)
    assert np.array_equal(_bits(val[:n_check].cpu().numpy()), _bits(_fold(wv)))
    # the workspace of the search does not grow with the corpus: all 1024 queries in ONE pass in well under 2 GB
    # (a [Q, N] f32 score buffer would be 36 GB here), also when EVERY query overflows its candidate list and is
    # rescored by the exhaustive fallback
    import ctypes as C
    mn, full = C.c_size_t(0), C.c_size_t(0)
    b2r._abi.check(b2r._abi.lib.b2r_search_workspace(C.byref(ix._desc), nq, k, C.byref(mn), C.byref(full)))
    assert full.value < 2 << 30 and ix._ws.numel() >= full.value
    b2r.set_fused_cap(1)
    try:
        xi, xv = ix.search(q_ptr, q_terms, q_w, k)
        torch.cuda.synchronize()
    finally:
        b2r.set_fused_cap(0)
    assert torch.equal(xi, idx) and torch.equal(xv, val)
    del h_data, h_ind, h_dl, ix
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard_range(n_docs, world, r)
            s, e2 = int(h_ptr[lo]), int(h_ptr[hi])
            sh = b2r.TermMajorIndex.from_csr(data[s:e2], ind[s:e2], ptr[lo:hi + 1] - s, dl[lo:hi], n_vocab=n_vocab,
                                             idf=idf, avgdl=avgdl, doc_id_base=lo)
            parts.append(sh.search(q_ptr, q_terms, q_w, k, return_keys=True)[2])
            del sh
        mi, mv = _merge(b2r, parts, nq, k)
        assert torch.equal(mi, idx) and torch.equal(mv, val), world


# ----------------------------------------------------------------------------------- C4
def test_config4_splade_shape_8p8m_docs_vs_oracle(b2r):
    """BASELINE config 4 at its named size: 8.8M docs x 30,522 vocab, 120 nnz/doc (1.056e9 postings), 30-term
    weighted queries, impact dot top-10 (simd_tfidf_score + top-k with idf == 1): 8 queries against the oracle, all
    256: fused selection == keys strictly descending and values == dense scores at the returned ids (8 queries)."""
    from b200ret import synthetic as S
    from gpu_synth import zipf_csr_torch
    dev = torch.device("cuda")
    n_docs, n_vocab, k, nq, n_check = 8_800_000, 30522, 10, 256, 8
    data, ind, ptr, _ = zipf_csr_torch(n_docs, n_vocab, 0, 20260103, dev, distinct_per_doc=120, chunk=1 << 18)
    assert int(ptr[-1]) > 1_000_000_000
    idf = np.ones(n_vocab, np.float32)
    q_ptr, q_terms, q_w = S.impact_queries(nq, n_vocab, 30)
    ix = b2r.TermMajorIndex.from_csr(data, ind, ptr, None, n_vocab=n_vocab, idf=idf, kind="impact")
    idx, val, keys = ix.search(q_ptr, q_terms, q_w, k, return_keys=True)
    torch.cuda.synchronize()
    ku = keys.cpu().numpy().view(np.uint64)
    assert bool((ku[:, :-1] > ku[:, 1:]).all())
    e = int(q_ptr[n_check])
    dense = ix.score_dense(q_ptr[:n_check + 1], q_terms[:e], q_w[:e])
    assert torch.equal(torch.gather(dense, 1, idx[:n_check]), val[:n_check])
    del ix
    h = [t.cpu().numpy() for t in (data, ind, ptr)]
    del data, ind
    c_oracle.use_all_host_threads()
    for q in range(n_check):
        qtf = np.zeros(n_vocab, np.float32)
        qtf[q_terms[q_ptr[q]:q_ptr[q + 1]]] = q_w[q_ptr[q]:q_ptr[q + 1]]
        s = c_oracle.tfidf_scores(qtf, h[0], h[1], h[2], idf)
        assert np.array_equal(_bits(dense[q].cpu().numpy()), _bits(s)), q          # all 8.8M scores of the query
        wi, wv = c_oracle.topk(s, k)
        assert np.array_equal(idx[q].cpu().numpy(), wi), q
        assert np.array_equal(_bits(val[q].cpu().numpy()), _bits(_fold(wv))), q


# ----------------------------------------------------------------------------------- C5
def test_config5_int8_10m_vectors_ids_and_values(b2r):
    """BASELINE config 5 at its named size on one GPU: 10M x 768 INT8 vectors, 1024 queries, top-100 (CTA-pair
    tcgen05 kernel): ids AND values of 8 queries against the oracle's exact-integer evaluation on the host
    (quantized_dot_product_batch semantics, chunked), and against a 4-shard scan + merge for all queries."""
    from gpu_synth import random_int8_corpus, random_int8_queries
    dev = torch.device("cuda")
    n, dim, k, nq, n_check = 10_000_000, 768, 100, 1024, 8
    d8, ds = random_int8_corpus(n, dim, 42, dev)
    q8, qs = random_int8_queries(nq, dim, 43, dev)
    idx, val, keys = b2r.int8_scan_topk(q8, d8, qs, ds, k)
    torch.cuda.synchronize()
    ku = keys.cpu().numpy().view(np.uint64)
    assert bool((ku[:, :-1] > ku[:, 1:]).all())
    parts = []
    for r in range(4):
        lo, hi = shard_range(n, 4, r)
        parts.append(b2r.int8_scan_topk(q8, d8[lo:hi], qs, ds[lo:hi], k, doc_id_base=lo)[2])
    mi, mv = _merge(b2r, parts, nq, k)
    assert torch.equal(mi, idx) and torch.equal(mv, val)
    # host: exact integer dots + f64 scale chain, 1M vectors at a time
    c_oracle.use_all_host_threads()
    hq8, hqs = q8[:n_check].cpu().numpy(), qs[:n_check].cpu().numpy()
    scores = np.empty((n_check, n), np.float32)
    step = 1_000_000
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        scores[:, lo:hi] = c_oracle.int8_dot_batch(hq8, d8[lo:hi].cpu().numpy(), hqs, ds[lo:hi].cpu().numpy())
    for q in range(n_check):
        wi, wv = np_oracle.topk_canonical(scores[q], k)
        assert np.array_equal(idx[q].cpu().numpy(), wi), q
        assert np.array_equal(_bits(val[q].cpu().numpy()), _bits(wv)), q
