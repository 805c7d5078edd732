"""Pins the oracle (numpy + C restatements) bit-for-bit against outputs of the reference's own
Numba kernels, recorded in tests/golden/ by oracle/gen_golden.py."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, np_oracle


def _queries(z):
    ptr = z["q_ptr"]
    return [(z["q_terms"][ptr[i]:ptr[i + 1]], z["q_weights"][ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


@pytest.fixture(scope="module")
def bm25(golden_dir):
    return np.load(os.path.join(golden_dir, "bm25_arrays.npz"))


def test_bm25_numpy_bit_exact(bm25):
    z = bm25
    n_vocab = len(z["idf"])
    csc = np_oracle.csr_to_csc(z["data"], z["indices"], z["indptr"], n_vocab)
    for q, (t, w) in enumerate(_queries(z)):
        qtf = np_oracle.dense_query(t, w, n_vocab)
        s = np_oracle.bm25_scores(qtf, z["data"], z["indices"], z["indptr"], z["doc_lengths"], z["idf"],
                                  float(z["k1"]), float(z["b"]), float(z["avgdl"]), csc=csc)
        assert np.array_equal(s.view(np.uint32), z["ref_scores"][q].view(np.uint32)), f"query {q}"


def test_bm25_c_bit_exact(bm25):
    z = bm25
    n_vocab = len(z["idf"])
    for q, (t, w) in enumerate(_queries(z)):
        qtf = np_oracle.dense_query(t, w, n_vocab)
        s = c_oracle.bm25_scores(qtf, z["data"], z["indices"], z["indptr"], z["doc_lengths"], z["idf"],
                                 float(z["k1"]), float(z["b"]), float(z["avgdl"]))
        assert np.array_equal(s.view(np.uint32), z["ref_scores"][q].view(np.uint32)), f"query {q}"


def test_bm25_fractional_inputs_bit_exact(golden_dir):
    """fractional tf / doc lengths / idf / query weights, k1=0.9, b=0.4: pins the f64 evaluation order."""
    z = np.load(os.path.join(golden_dir, "bm25_frac.npz"))
    n_vocab = len(z["idf"])
    for q, (t, w) in enumerate(_queries(z)):
        qtf = np_oracle.dense_query(t, w, n_vocab)
        for fn in (np_oracle.bm25_scores, c_oracle.bm25_scores):
            s = fn(qtf, z["data"], z["indices"], z["indptr"], z["doc_lengths"], z["idf"], float(z["k1"]),
                   float(z["b"]), float(z["avgdl"]))
            assert np.array_equal(s.view(np.uint32), z["ref_scores"][q].view(np.uint32))


def test_bm25_k1_1000_b0(bm25):
    """registry 'tfidf' parameterisation (retriever_registry.py:593-595) through the same kernel."""
    z = bm25
    n_vocab = len(z["idf"])
    for q, (t, w) in enumerate(_queries(z)[:4]):
        qtf = np_oracle.dense_query(t, w, n_vocab)
        for fn in (np_oracle.bm25_scores, c_oracle.bm25_scores):
            s = fn(qtf, z["data"], z["indices"], z["indptr"], z["doc_lengths"], z["idf"], 1000.0, 0.0,
                   float(z["avgdl"]))
            assert np.array_equal(s.view(np.uint32), z["ref_scores_k1000"][q].view(np.uint32))


def test_bm25_fixture_has_the_hard_cases(bm25):
    z = bm25
    assert (z["idf"] < 0).sum() >= 2                       # negative idf (df > N/2)
    assert (np.diff(z["indptr"]) == 0).sum() >= 10         # empty docs keep a row
    assert (z["ref_scores"] < 0).any() and (z["ref_scores"] == 0).any()


def test_topk_values_match_reference(bm25):
    """ids are only comparable where the reference's arbitrary tie order cannot matter: compare the
    sorted score VALUES everywhere and the ids wherever the k+1 best scores are distinct."""
    z = bm25
    for q in range(z["ref_scores"].shape[0]):
        s = z["ref_scores"][q]
        for fn in (np_oracle.topk_canonical, c_oracle.topk):
            idx, val = fn(s, 10)
            assert np.array_equal(val, z["ref_top_val"][q])
            assert np.array_equal(val, s[idx])
            top11 = np.sort(s)[::-1][:11]
            if len(np.unique(top11)) == 11:
                assert np.array_equal(idx, z["ref_top_idx"][q])


def test_topk_reference_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "topk_cases.npz"))
    for name in ("normal", "uniform", "zipfian", "bimodal", "k_ge_n", "big"):
        s, k = z[f"{name}_scores"], int(z[f"{name}_k"])
        for fn in (np_oracle.topk_canonical, c_oracle.topk):
            idx, val = fn(s, k)
            assert len(idx) == min(k, len(s))
            assert np.array_equal(val, z[f"{name}_ref_val"]), name
            if len(np.unique(s)) == len(s):                 # no ties at all -> ids are pinned too
                assert np.array_equal(idx, z[f"{name}_ref_idx"]), name


def test_topk_tie_rule():
    s = np.array([1.0, 2.0, 2.0, -0.0, 0.0, np.nan, 2.0, 1.0], np.float32)
    for fn in (np_oracle.topk_canonical, c_oracle.topk):
        idx, val = fn(s, 8)
        assert idx.tolist() == [1, 2, 6, 0, 7, 3, 4, 5]
        idx, _ = fn(s, 2)
        assert idx.tolist() == [1, 2]


def test_tfidf_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "tfidf_arrays.npz"))
    n_vocab = len(z["idf"])
    ones = np.ones(n_vocab, np.float32)
    for q, (t, w) in enumerate(_queries(z)):
        qtf = np_oracle.dense_query(t, w, n_vocab)
        for fn in (np_oracle.tfidf_scores, c_oracle.tfidf_scores):
            s = fn(qtf, z["data"], z["indices"], z["indptr"], z["idf"])
            assert np.array_equal(s.view(np.uint32), z["ref_scores"][q].view(np.uint32))
            s = fn(qtf, z["data"], z["indices"], z["indptr"], ones)
            assert np.array_equal(s.view(np.uint32), z["ref_scores_idf1"][q].view(np.uint32))


def test_int8_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "int8.npz"))
    for fn in (np_oracle.int8_dot_batch, c_oracle.int8_dot_batch):
        s = fn(z["q8"], z["d8"], z["qscale"], z["dscale"])
        assert np.array_equal(s.view(np.uint32), z["ref_sims"].view(np.uint32))
    q, sc = np_oracle.quantize_rows(z["emb"])
    assert np.array_equal(q, z["d8"][:len(q)]) and np.array_equal(sc, z["dscale"][:len(q)])


def test_service_text_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "service_text.json")) as f:
        g = json.load(f)
    ix = np_oracle.build_text_index(g["corpus"])
    assert len(ix["vocabulary"]) == g["meta"]["vocab_size"]
    assert ix["avgdl"] == g["meta"]["avgdl"]
    assert len(ix["data"]) == g["meta"]["nnz"]
    assert sorted(ix["vocabulary"], key=ix["vocabulary"].get)[:8] == g["meta"]["vocab_head"]
    assert float(np.sum(ix["idf"].astype(np.float64))) == g["meta"]["idf_sum"]
    got = np_oracle.search_bm25_text(ix, g["queries"], top_k=10)
    for qid, ref in g["ref_top10"].items():
        mine = got[qid]
        # score multiset is pinned exactly; ids wherever the reference's tie order is not in play
        assert sorted(mine.values(), reverse=True) == sorted(ref.values(), reverse=True), qid
        if len(set(ref.values())) == len(ref) and len(ref) < 10:
            assert list(mine) == list(ref), qid
    assert g["ref_top10"]["blank"] == {} and g["ref_top10"]["oov"] == {}
    got500 = np_oracle.search_bm25_text(ix, {k: g["queries"][k] for k in g["ref_top500"]}, top_k=500)
    for qid, ref in g["ref_top500"].items():
        assert got500[qid].keys() == ref.keys()
        for d, v in ref.items():
            assert got500[qid][d] == v


def test_hybrid_rerank_restatement_is_consistent_with_the_pinned_pieces():
    """np_oracle.hybrid_rerank (the fusion rule has no reference implementation to pin it to) must at least agree
    with the pinned pieces it is made of: dense part == int8_dot_batch restricted to the candidates, selection ==
    topk_canonical on the fused scores, padding and shard windows as documented."""
    rng = np.random.default_rng(3)
    nq, n, dim, k_in, k_out, base = 5, 300, 48, 20, 6, 1000
    q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
    d8 = rng.integers(-127, 128, (n, dim)).astype(np.int8)
    qs = (rng.random(nq).astype(np.float32) + 0.01) / 127
    ds = rng.random(n).astype(np.float32) + 0.01
    cand = np.stack([rng.choice(n, k_in, replace=False) for _ in range(nq)]).astype(np.int64) + base
    cand[0, 3] = -1
    cand[1, 5] = base + n            # outside the shard
    cand[2, :] = -1                  # nothing to rerank
    sparse = (rng.random((nq, k_in)) * 10).astype(np.float32)
    full = np_oracle.int8_dot_batch(q8, d8, qs, ds)
    for sp in (sparse, None):
        idx, val, dense = np_oracle.hybrid_rerank(cand, sp, q8, qs, d8, ds, 0.3, 0.7, k_out, doc_id_base=base)
        assert idx.shape == (nq, k_out) and val.shape == (nq, k_out) and dense.shape == (nq, k_in)
        assert (idx[2] == -1).all() and np.isneginf(val[2]).all() and np.isneginf(dense[2]).all()
        for q in (0, 1, 3, 4):
            ok = (cand[q] >= base) & (cand[q] < base + n)
            assert np.array_equal(dense[q, ok], full[q, cand[q, ok] - base]) and np.isneginf(dense[q, ~ok]).all()
            fused = dense[q, ok] if sp is None else (
                np.float64(0.3) * sp[q, ok].astype(np.float64)
                + np.float64(0.7) * dense[q, ok].astype(np.float64)).astype(np.float32)
            order = np.lexsort((cand[q, ok], -fused))[:k_out]
            assert np.array_equal(idx[q, :len(order)], cand[q, ok][order])
            assert np.array_equal(val[q, :len(order)], fused[order])


# ----------------------------------------------------------------------------------- BASELINE config 1
def _load_synthetic_module():
    """b200ret/synthetic.py is numpy-only: load it without the package (no CUDA library needed)."""
    import importlib.util
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "b2r_synthetic", os.path.join(here, "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200",
                                      "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fiqa_shape_config1_oracle_vs_reference(golden_dir):
    """BASELINE config 1 at its full size (57,638 docs, 648 queries, top-10): the restated corpus generator
    reproduces the reference generator's corpus (sha256 recorded by oracle/gen_golden.py), and the oracle's
    text-level restatement gives the canonical top-10 of the reference's own score vectors bit for bit."""
    z = np.load(os.path.join(golden_dir, "fiqa_shape.npz"))
    meta = json.loads(str(z["meta"]))
    syn = _load_synthetic_module()
    corpus = syn.fiqa_shape_corpus()
    assert syn.corpus_sha256(corpus) == meta["corpus_sha256"]
    queries = syn.fiqa_shape_queries()
    ix = np_oracle.build_text_index(corpus)
    assert len(ix["doc_ids"]) == meta["n_docs"] and len(ix["vocabulary"]) == meta["vocab_size"]
    assert len(ix["data"]) == meta["nnz"] and ix["avgdl"] == meta["avgdl"]
    assert float(np.sum(ix["idf"].astype(np.float64))) == meta["idf_sum"]
    # pack the queries like RetrievalService._score_bm25_query does (retrieval.py:236-252)
    from collections import Counter
    ptr, terms, weights, live = [0], [], [], []
    for qi, text in enumerate(queries.values()):
        c = Counter(np_oracle.tokenize(text))
        tw = sorted((ix["vocabulary"][t], float(n)) for t, n in c.items() if t in ix["vocabulary"])
        if tw:
            live.append(qi)
            terms += [t for t, _ in tw]
            weights += [w for _, w in tw]
            ptr.append(len(terms))
    k = z["canon_idx"].shape[1]
    gi, gv = c_oracle.bm25_search_batch(np.asarray(ptr, np.int32), np.asarray(terms, np.int32),
                                        np.asarray(weights, np.float32), len(ix["vocabulary"]), ix["data"],
                                        ix["indices"], ix["indptr"], ix["doc_lengths"], ix["idf"], 1.2, 0.75,
                                        ix["avgdl"], k)
    assert np.array_equal(gi, z["canon_idx"][live])
    assert np.array_equal(gv.view(np.uint32), z["canon_val"][live].view(np.uint32))
    dead = np.setdiff1d(np.arange(len(queries)), live)
    assert (z["canon_idx"][dead] == -1).all() and (z["ref_idx"][dead] == -1).all()
    # what search_bm25 itself returned: same scores as the canonical list (positive part); same ids wherever the
    # reference's unspecified tie order is not in play
    n_empty = 0
    for qi in range(len(queries)):
        pos = z["canon_val"][qi] > 0
        n_ref = int((z["ref_idx"][qi] >= 0).sum())
        n_empty += n_ref == 0
        assert n_ref == int(pos.sum()), qi
        assert np.array_equal(z["ref_val"][qi, :n_ref], z["canon_val"][qi, :n_ref]), qi
        untied = np.ones(n_ref, bool)
        v = z["canon_val"][qi, :n_ref]
        untied[1:] &= v[1:] != v[:-1]
        untied[:-1] &= v[:-1] != v[1:]
        if n_ref == k:
            untied &= v != v[-1]            # the boundary value may tie with documents outside the list
        assert np.array_equal(z["ref_idx"][qi, :n_ref][untied], z["canon_idx"][qi, :n_ref][untied]), qi
    assert n_empty == meta["n_empty_results"]
