"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden vectors recorded
from the reference and against the CPU oracle on seeded inputs.  Bar: bit-exact f32 scores (tolerance
stated in the north star: 1e-5 relative -- we hold 0 ulp) and bit-exact top-k ids under the rule
(score desc, doc index asc)."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import c_oracle, np_oracle  # noqa: E402  (the checker)


@pytest.fixture(scope="module")
def b2r():
    import b200ret
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    assert os.path.exists(b200ret._abi.LIB_PATH)
    return b200ret


PAIR_DEFAULT = True      # default of b2r_set_int8_pair in the library (restored after tests that flip it)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _queries(z):
    p = z["q_ptr"]
    return [(z["q_terms"][p[i]:p[i + 1]], z["q_weights"][p[i]:p[i + 1]]) for i in range(len(p) - 1)]


# ----------------------------------------------------------------------------------- golden: K1
@pytest.mark.parametrize("tile_docs", [256, 1024, 4096])
def test_bm25_scores_bit_exact_vs_reference(b2r, golden_dir, tile_docs):
    z = np.load(os.path.join(golden_dir, "bm25_arrays.npz"))
    ix = b2r.TermMajorIndex.from_csr(z["data"], z["indices"], z["indptr"], z["doc_lengths"], n_vocab=len(z["idf"]),
                                     idf=z["idf"], avgdl=float(z["avgdl"]), k1=float(z["k1"]), b=float(z["b"]),
                                     tile_docs=tile_docs)
    s = ix.score_dense(z["q_ptr"], z["q_terms"], z["q_weights"]).cpu().numpy()
    assert np.array_equal(_bits(s), _bits(z["ref_scores"]))
    idx, val = ix.search(z["q_ptr"], z["q_terms"], z["q_weights"], 10)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for q in range(s.shape[0]):
        wi, wv = np_oracle.topk_canonical(z["ref_scores"][q], 10)
        assert np.array_equal(idx[q], wi), q
        assert np.array_equal(_bits(val[q]), _bits(np.where(wv == 0, np.float32(0), wv))), q
        assert np.array_equal(val[q], z["ref_top_val"][q])          # the reference's own top-k values


def test_bm25_function_signature_drop_in(b2r, golden_dir):
    """simd_bm25_score / simd_bm25_batch_score with the reference's positional signature, one query at a time."""
    z = np.load(os.path.join(golden_dir, "bm25_arrays.npz"))
    nv = len(z["idf"])
    indptr32 = z["indptr"].astype(np.int32)
    for q, (t, w) in enumerate(_queries(z)[:6]):
        qtf = np_oracle.dense_query(t, w, nv)
        s = b2r.simd_bm25_score(qtf, z["data"], z["indices"], indptr32, z["doc_lengths"], z["idf"],
                                float(z["k1"]), float(z["b"]), float(z["avgdl"]))
        assert s.dtype == np.float32 and s.shape == (len(indptr32) - 1,)
        assert np.array_equal(_bits(s), _bits(z["ref_scores"][q]))
        ti, tv = b2r.fast_topk_selection(s, 10)
        assert ti.dtype == np.int64 and np.array_equal(tv, z["ref_top_val"][q])
    for q, (t, w) in enumerate(_queries(z)[:4]):   # registry "tfidf" parameterisation k1=1000, b=0
        qtf = np_oracle.dense_query(t, w, nv)
        s = b2r.simd_bm25_batch_score(qtf, z["data"], z["indices"], indptr32, z["doc_lengths"], z["idf"], 1000.0,
                                      0.0, float(z["avgdl"]))
        assert np.array_equal(_bits(s), _bits(z["ref_scores_k1000"][q]))
    s = b2r.optimized_bm25_score(np_oracle.dense_query(*_queries(z)[0], nv), (z["data"], z["indices"], z["indptr"]),
                                 z["doc_lengths"], z["idf"], k1=float(z["k1"]), b=float(z["b"]))
    assert np.array_equal(_bits(s), _bits(z["ref_scores"][0]))
    b2r.clear_index_cache()


def test_bm25_fractional_inputs(b2r, golden_dir):
    z = np.load(os.path.join(golden_dir, "bm25_frac.npz"))
    ix = b2r.TermMajorIndex.from_csr(z["data"], z["indices"], z["indptr"], z["doc_lengths"], n_vocab=len(z["idf"]),
                                     idf=z["idf"], avgdl=float(z["avgdl"]), k1=float(z["k1"]), b=float(z["b"]),
                                     tile_docs=512)
    s = ix.score_dense(z["q_ptr"], z["q_terms"], z["q_weights"]).cpu().numpy()
    assert np.array_equal(_bits(s), _bits(z["ref_scores"]))


# ----------------------------------------------------------------------------------- golden: K3, K4
def test_tfidf_bit_exact_vs_reference(b2r, golden_dir):
    z = np.load(os.path.join(golden_dir, "tfidf_arrays.npz"))
    nv = len(z["idf"])
    ix = b2r.TermMajorIndex.from_csr(z["data"], z["indices"], z["indptr"], n_vocab=nv, idf=z["idf"], kind="impact",
                                     tile_docs=256)
    s = ix.score_dense(z["q_ptr"], z["q_terms"], z["q_weights"]).cpu().numpy()
    assert np.array_equal(_bits(s), _bits(z["ref_scores"]))
    ix.set_idf(np.ones(nv, np.float32))
    s = ix.score_dense(z["q_ptr"], z["q_terms"], z["q_weights"]).cpu().numpy()
    assert np.array_equal(_bits(s), _bits(z["ref_scores_idf1"]))
    t, w = _queries(z)[3]
    s1 = b2r.simd_tfidf_score(np_oracle.dense_query(t, w, nv), z["data"], z["indices"], z["indptr"].astype(np.int32),
                              z["idf"])
    assert np.array_equal(_bits(s1), _bits(z["ref_scores"][3]))
    b2r.clear_index_cache()


def test_int8_bit_exact_vs_reference(b2r, golden_dir):
    z = np.load(os.path.join(golden_dir, "int8.npz"))
    s = b2r.quantized_dot_product_batch(z["q8"], z["d8"], z["qscale"], z["dscale"])
    assert np.array_equal(_bits(s), _bits(z["ref_sims"]))
    idx, val, _ = b2r.int8_scan_topk(z["q8"], z["d8"], z["qscale"], z["dscale"], 20)
    for q in range(len(z["q8"])):
        wi, wv = np_oracle.topk_canonical(z["ref_sims"][q], 20)
        assert np.array_equal(idx[q].cpu().numpy(), wi) and np.array_equal(_bits(val[q].cpu().numpy()), _bits(wv))


def test_int8_odd_shapes(b2r):
    rng = np.random.default_rng(5)
    for nq, n, dim in [(1, 1, 16), (3, 130, 48), (33, 257, 768), (5, 64, 100)]:
        q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
        d8 = rng.integers(-127, 128, (n, dim)).astype(np.int8)
        qs = rng.random(nq).astype(np.float32) + 0.01
        ds = rng.random(n).astype(np.float32) + 0.01
        got = b2r.quantized_dot_product_batch(q8, d8, qs, ds)
        assert np.array_equal(_bits(got), _bits(np_oracle.int8_dot_batch(q8, d8, qs, ds))), (nq, n, dim)


@pytest.mark.parametrize("nq,n,dim", [(300, 10_000, 768), (128, 4096, 128), (129, 1000, 512), (5, 130, 256)])
def test_int8_tensor_core_path_equals_dp4a_and_oracle(b2r, nq, n, dim):
    """tcgen05 kind::i8 kernel (dim % 128 == 0, dim <= 768) vs the dp4a kernel vs the oracle: bit-exact."""
    rng = np.random.default_rng(nq + n + dim)
    q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
    d8 = rng.integers(-127, 128, (n, dim)).astype(np.int8)
    qs = (rng.random(nq).astype(np.float32) + 0.01) / 127
    ds = rng.random(n).astype(np.float32) + 0.01
    b2r.set_int8_mma(True)
    a = b2r.quantized_dot_product_batch(q8, d8, qs, ds)
    b2r.set_int8_mma(False)
    c = b2r.quantized_dot_product_batch(q8, d8, qs, ds)
    b2r.set_int8_mma(True)
    assert np.array_equal(_bits(a), _bits(c))
    want = np_oracle.int8_dot_batch(q8, d8, qs, ds)
    assert np.array_equal(_bits(a), _bits(want))
    idx, val, _ = b2r.int8_scan_topk(q8, d8, qs, ds, 50)
    for q in (0, nq // 2, nq - 1):
        wi, wv = np_oracle.topk_canonical(want[q], 50)
        assert np.array_equal(idx[q].cpu().numpy(), wi) and np.array_equal(_bits(val[q].cpu().numpy()), _bits(wv))


@pytest.mark.parametrize("k", [10, 100])
def test_int8_fused_scan_equals_plain_and_survives_overflow(b2r, k):
    """Fused scan (threshold from every 32nd/16th 128-doc tile, candidate lists) vs the plain chunked path vs exact
    integer math; zeroing the sample tiles makes the threshold useless and forces the gated exhaustive fallback."""
    rng = np.random.default_rng(77 + k)
    nq, n, dim = 130, 70_000 + 13, 768
    q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
    d8 = rng.integers(-127, 128, (n, dim)).astype(np.int8)
    qs = (rng.random(nq).astype(np.float32) + 0.01) / 127
    ds = rng.random(n).astype(np.float32) + 0.01
    d8_adv = d8.copy()
    d8_adv[((np.arange(n) // 128) % 16) == 0] = 0           # every sample tile scores exactly 0
    # scales outside the range the epilogue's f32 pre-filter is proven for (it must then take the exact path),
    # zero and negative scales (negative / zero thresholds)
    ds_odd = rng.choice(np.array([0.0, 1e-20, 3e-16, 1.0, 0.37, -0.5, 1e16, 1e20], np.float32), n)
    qs_odd = rng.choice(np.array([1 / 127, 0.0, 1e-20, -1e-3, 1e10, 0.02], np.float32), nq)
    for corpus, qs, ds in ((d8, qs, ds), (d8_adv, qs, ds), (d8, qs_odd, ds_odd), (d8, qs, -ds)):
        b2r.set_int8_fused(1)
        fi, fv, _ = b2r.int8_scan_topk(q8, corpus, qs, ds, k, doc_id_base=1000)
        b2r.set_int8_fused(0)
        pi, pv, _ = b2r.int8_scan_topk(q8, corpus, qs, ds, k, doc_id_base=1000)
        b2r.set_int8_fused(1)
        assert torch.equal(fi, pi) and torch.equal(fv, pv)
        for q in (0, 1, 2, 3, 64, nq - 1):
            want = np_oracle.int8_dot_batch(q8[q:q + 1], corpus, qs[q:q + 1], ds)[0]
            wi, wv = np_oracle.topk_canonical(want, k)
            assert np.array_equal(fi[q].cpu().numpy(), wi + 1000)
            assert np.array_equal(_bits(fv[q].cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv)))


@pytest.mark.parametrize("dim", [768, 100])
def test_int8_rerank_on_candidates_vs_oracle(b2r, dim):
    """f3: candidate-only INT8 rerank + score fusion against the numpy restatement (bit-exact scores, exact ids),
    with missing candidates (-1), candidates of another shard, duplicates of scores (ties -> lower doc id first),
    pure dense rerank (no sparse scores) and the dense similarities equal to quantized_dot_product_batch."""
    rng = np.random.default_rng(61 + dim)
    nq, n, k_in, k_out, base = 37, 5000, 100, 10, 20_000
    q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
    d8 = rng.integers(-127, 128, (n, dim)).astype(np.int8)
    d8[100:140] = d8[100]                                             # identical vectors: score ties
    qs = (rng.random(nq).astype(np.float32) + 0.01) / 127
    ds = rng.random(n).astype(np.float32) + 0.01
    ds[100:140] = ds[100]
    cand = np.stack([rng.choice(n, k_in, replace=False) for _ in range(nq)]).astype(np.int64) + base
    cand[0, :40] = np.arange(100, 140) + base                         # the tied block, all with one sparse score
    cand[1, 5:9] = -1
    cand[2, 50:] = -1
    cand[3, 10] = base + n + 5                                        # beyond this shard
    cand[3, 11] = base - 1
    sparse = (rng.random((nq, k_in)).astype(np.float32) * 20).astype(np.float32)
    sparse[0, :40] = 3.25
    for sp, (w_s, w_d) in ((sparse, (0.3, 0.7)), (sparse, (1.0, 0.0)), (None, (0.3, 0.7))):
        gi, gv, gd = b2r.int8_rerank(cand, sp, q8, qs, d8, ds, k_out, sparse_weight=w_s, dense_weight=w_d,
                                     doc_id_base=base, return_dense=True)
        wi, wv, wd = np_oracle.hybrid_rerank(cand, sp, q8, qs, d8, ds, w_s, w_d, k_out, doc_id_base=base)
        assert np.array_equal(gi.cpu().numpy(), wi)
        assert np.array_equal(_bits(gv.cpu().numpy()), _bits(wv))
        assert np.array_equal(_bits(gd.cpu().numpy()), _bits(wd))
    full = b2r.quantized_dot_product_batch(q8, d8, qs, ds)            # the reference kernel's drop-in, whole corpus
    ok = (cand >= base) & (cand < base + n)
    for q in (0, 1, 3, nq - 1):
        assert np.array_equal(_bits(gd.cpu().numpy()[q][ok[q]]), _bits(full[q][cand[q][ok[q]] - base]))


def test_hybrid_search_two_stage(b2r):
    """BM25 top-100 -> INT8 rerank -> top-10 through hybrid_search equals oracle BM25 top-100 + oracle rerank."""
    from b200ret import synthetic as S
    n_docs, n_vocab, dim = 30_000, 3000, 256
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 40, seed=71)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(24, n_vocab, seed=72)
    rng = np.random.default_rng(73)
    x = rng.standard_normal((n_docs, dim)).astype(np.float32)
    d8, ds = np_oracle.quantize_rows(x)
    xq = rng.standard_normal((24, dim)).astype(np.float32)
    q8, qs = np_oracle.quantize_rows(xq); qs = (qs / 127).astype(np.float32)
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, tile_docs=1024)
    gi, gv = b2r.hybrid_search(ix, q_ptr, q_terms, q_w, q8, qs, d8, ds, 100, 10, sparse_weight=0.3, dense_weight=0.7)
    ci, cv = _oracle_topk((data, indices, indptr, dl, idf, 1.2, 0.75, avgdl), q_ptr, q_terms, q_w, 100)
    cv = np.where(cv == 0, np.float32(0), cv)
    wi, wv, _ = np_oracle.hybrid_rerank(ci, cv, q8, qs, d8, ds, 0.3, 0.7, 10)
    assert np.array_equal(gi.cpu().numpy(), wi) and np.array_equal(_bits(gv.cpu().numpy()), _bits(wv))


@pytest.mark.parametrize("n,dim,nq", [(20_000, 768, 1), (5003, 100, 11), (300, 7, 3)])
def test_dense_topk_vs_numpy(b2r, n, dim, nq):
    """a8 (search_by_vector): fp32 gemv + top-k.  BLAS fixes no summation order, so the scores are held to the
    dot-product error scale (1e-5 * sum |a_i b_i|, in practice a few ulp) against an f64 evaluation; the selection
    must be exact on the scores the kernel produced."""
    rng = np.random.default_rng(81 + dim)
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    k = 10
    idx, val, sc = b2r.dense_topk(emb, q if nq > 1 else q[0], k, return_scores=True)
    sc, idx, val = sc.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy()
    exact = emb.astype(np.float64) @ q.astype(np.float64).T
    scale = np.abs(emb).astype(np.float64) @ np.abs(q).astype(np.float64).T
    assert np.all(np.abs(sc.T - exact) <= 1e-5 * scale + 1e-30)
    assert np.all(np.abs(sc.T - (emb @ q.T)) <= 2e-5 * scale + 1e-30)          # the reference's np.dot (f32 BLAS)
    for i in range(nq):
        wi, wv = np_oracle.topk_canonical(sc[i], k)
        assert np.array_equal(idx[i], wi) and np.array_equal(_bits(val[i]), _bits(wv))


def test_service_search_by_vector(b2r, tmp_path):
    rng = np.random.default_rng(91)
    n, dim = 500, 64
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    emb.tofile(tmp_path / "emb.f32")
    store = tmp_path / "docs.idx"
    b2r.MemoryIndex(store, create=True).close()
    svc = b2r.RetrievalService(store, embedding_path=tmp_path / "emb.f32")
    with pytest.raises(ValueError, match="No embedding index"):
        svc.search_by_vector(emb[0])                       # like the reference: nothing is mapped before doc_ids exist
    svc.doc_ids = [f"d{i}" for i in range(n)]
    svc._load_embeddings()
    got = svc.search_by_vector(emb[17], k=5)
    sims = emb @ emb[17]
    assert got[0]["doc_id"] == "d17" and len(got) == 5
    assert [g["doc_id"] for g in got] == [f"d{i}" for i in np.argsort(-sims)[:5]]
    assert np.allclose([g["score"] for g in got], np.sort(sims)[::-1][:5], rtol=1e-5)
    cut = float(np.sort(sims)[::-1][2])
    assert len(svc.search_by_vector(emb[17], k=5, min_score=cut - 1e-4)) == 3


@pytest.mark.parametrize("nq,n,dim", [(130, 70_000 + 13, 768), (512, 40_000, 768), (257, 33_000 + 129, 768),
                                      (300, 36_000, 128), (200, 34_000 + 7, 384)])
def test_int8_pair_scan_equals_single_cta_and_oracle(b2r, nq, n, dim):
    """CTA-pair fused scan (tcgen05.mma cta_group::2, 256 x 256 pair tiles) vs the single-CTA fused scan vs exact
    integer math: odd query counts (zero-padded operand halves), an odd number of 128-document tiles (the second
    CTA of the last pair sees only padding), overflowing candidate lists."""
    rng = np.random.default_rng(101 + nq)
    k = 100
    q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
    d8 = rng.integers(-127, 128, (n, dim)).astype(np.int8)
    qs = (rng.random(nq).astype(np.float32) + 0.01) / 127
    ds = rng.random(n).astype(np.float32) + 0.01
    d8_adv = d8.copy()
    d8_adv[((np.arange(n) // 128) % 16) == 0] = 0           # useless threshold: candidate lists overflow
    for corpus in (d8, d8_adv):
        try:
            b2r.set_int8_pair(True)
            pi, pv, _ = b2r.int8_scan_topk(q8, corpus, qs, ds, k, doc_id_base=7)
            b2r.set_int8_pair(False)
            si, sv, _ = b2r.int8_scan_topk(q8, corpus, qs, ds, k, doc_id_base=7)
        finally:
            b2r.set_int8_pair(PAIR_DEFAULT)
        assert torch.equal(pi, si) and torch.equal(pv, sv)
        for q in (0, 127, 128, nq - 1):
            want = np_oracle.int8_dot_batch(q8[q:q + 1], corpus, qs[q:q + 1], ds)[0]
            wi, wv = np_oracle.topk_canonical(want, k)
            assert np.array_equal(pi[q].cpu().numpy(), wi + 7)
            assert np.array_equal(_bits(pv[q].cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv)))


@pytest.mark.parametrize("rows,n,k", [(64, 300_000, 10), (33, 262_144 + 4, 100), (8, 1_000_000, 128), (5, 140_000, 1)])
def test_topk_threshold_filter_path(b2r, rows, n, k):
    """Long aligned rows with k <= 128 take the threshold-filter path of b2r_topk (group maxima -> k-th largest ->
    one filtering pass -> top-k of the candidate lists); rows it cannot finish (massive ties at the threshold, NaN
    rows, all-equal rows) must come out of the gated streaming selector with the same canonical result."""
    rng = np.random.default_rng(111 + k)
    s = rng.standard_normal((rows, n)).astype(np.float32)
    s[1] = np.round(s[1] * 2) / 2                         # heavy ties everywhere
    s[2, : n // 2] = np.nan                               # NaNs rank last
    s[3] = 1.25                                           # all equal: lowest indices win
    if rows > 4:
        s[4, :] = -np.abs(s[4])
        s[4, 100:100 + k] = 0.0                           # threshold exactly 0, zeros tie with -0.0
        s[4, 7] = -0.0
    idx, val = b2r.fast_topk_selection(torch.from_numpy(s).cuda(), k)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for r in range(rows):
        wi, wv = np_oracle.topk_canonical(s[r], k)
        assert np.array_equal(idx[r], wi), r
        assert np.array_equal(_bits(val[r]), _bits(s[r][wi])), r


def test_int8_dense_epilogue_exact_on_ties_and_at_scale(b2r):
    """quantized_dot_product_batch must return f32((f64(dot) * f64(qs)) * f64(ds)) bit for bit whatever the
    epilogue does internally (a conversion-free f32 evaluation with an f64 fallback near ties was measured in round 1
    and passed this test, but was slower than the plain f64 chain and is not in the library).  (1) crafted exact ties:
    odd dots in [2^23/1.5, 2^24/1.5) with qs = 1, ds = 1.5 put dot * 1.5 exactly half-way between two f32 numbers
    (round-to-even must win), plus powers of two, zero dots with signed scales, extreme scales; (2) 25 M random
    outputs against an exact f64 evaluation."""
    dim = 768
    q = np.zeros((4, dim), np.int8); q[:, :600] = 127; q[:, 600:664] = 1
    q[1] = -q[1]; q[2] = 0; q[3, :] = 0; q[3, 0] = 64
    n = 400
    d = np.zeros((n, dim), np.int8); d[:, :600] = 127                    # 127 * 127 * 600 = 9,677,400 (even)
    rng = np.random.default_rng(5)
    d[:, 600:664] = rng.integers(-127, 128, (n, 64))                      # dot = 9,677,400 + sum: odd half of the time
    d[7] = 0; d[8] = 0; d[8, 0] = 64                                      # dot 0 and 4096 (a power of two) with q[3]
    qs = np.array([1.0, 1.0, 0.25, 1.0], np.float32)
    ds = np.full(n, 1.5, np.float32)
    ds[1::4] = -1.5; ds[2::8] = 3e-13; ds[3::8] = 7e13; ds[5] = 0.0; ds[6] = 0.75
    got = b2r.quantized_dot_product_batch(q, d, qs, ds)
    want = np_oracle.int8_dot_batch(q, d, qs, ds)
    assert np.array_equal(_bits(got), _bits(want))
    dots = q.astype(np.int64) @ d.astype(np.int64).T
    assert (np.abs(dots[0]) % 2 == 1).sum() > 50                          # the exact ties are really there
    rng = np.random.default_rng(6)
    nq, n2 = 256, 100_000
    q8 = rng.integers(-127, 128, (nq, dim)).astype(np.int8)
    d8 = rng.integers(-127, 128, (n2, dim)).astype(np.int8)
    qs2 = ((rng.random(nq) + 0.01) / 127).astype(np.float32)
    ds2 = (rng.random(n2) + 0.01).astype(np.float32)
    got = b2r.quantized_dot_product_batch(q8, d8, qs2, ds2)
    dd = q8.astype(np.float64) @ d8.astype(np.float64).T                  # exact: |dot| < 2^24
    want = ((dd * qs2.astype(np.float64)[:, None]) * ds2.astype(np.float64)[None, :]).astype(np.float32)
    assert np.array_equal(_bits(got), _bits(want))


# ----------------------------------------------------------------------------------- golden + edge: K2
def test_topk_reference_cases(b2r, golden_dir):
    z = np.load(os.path.join(golden_dir, "topk_cases.npz"))
    for name in ("normal", "uniform", "zipfian", "bimodal", "k_ge_n", "big"):
        s, k = z[f"{name}_scores"], int(z[f"{name}_k"])
        idx, val = b2r.fast_topk_selection(s, k)
        wi, wv = np_oracle.topk_canonical(s, k)
        assert np.array_equal(idx, wi) and np.array_equal(_bits(val), _bits(wv)), name
        assert np.array_equal(val, z[f"{name}_ref_val"]), name      # the reference's values
        assert np.array_equal(b2r.fast_topk(s, k), wi)


def test_topk_edge_cases(b2r):
    s = np.array([1.0, 2.0, 2.0, -0.0, 0.0, np.nan, 2.0, 1.0], np.float32)
    idx, val = b2r.fast_topk_selection(s, 8)
    assert idx.tolist() == [1, 2, 6, 0, 7, 3, 4, 5]
    assert np.array_equal(_bits(val), _bits(s[idx]))                # -0.0 and NaN come back bit-for-bit
    assert b2r.fast_topk_selection(s, 2)[0].tolist() == [1, 2]
    assert b2r.fast_topk_selection(s, 100)[0].tolist() == [1, 2, 6, 0, 7, 3, 4, 5]      # k >= n: all, sorted
    assert b2r.fast_topk_selection(np.zeros(1, np.float32), 1)[0].tolist() == [0]
    i32, _ = b2r.fast_topk_selection(s, 3, index_dtype=np.int32)    # pipeline variant returns int32
    assert i32.dtype == np.int32
    z = np.zeros(100_000, np.float32)                               # all ties: lowest indices win
    assert b2r.fast_topk_selection(z, 17)[0].tolist() == list(range(17))
    asc = np.arange(300_000, dtype=np.float32)                      # ascending: every element beats the threshold
    assert b2r.fast_topk_selection(asc, 5)[0].tolist() == [299999, 299998, 299997, 299996, 299995]
    neg = -np.arange(70_000, dtype=np.float32) - 1                  # all negative
    assert b2r.fast_topk_selection(neg, 3)[0].tolist() == [0, 1, 2]


@pytest.mark.parametrize("n,k", [(5000, 1), (4097, 10), (100_000, 100), (1_000_003, 10), (1_000_003, 1000),
                                 (50_000, 1024), (3000, 2000), (70_000, 5000)])
def test_topk_random_with_ties(b2r, n, k):
    rng = np.random.default_rng(n + k)
    s = np.round(rng.gamma(2.0, 2.0, n), 1).astype(np.float32)      # heavy ties
    s[rng.integers(0, n, n // 50)] *= -1
    idx, val = b2r.fast_topk_selection(s, k)
    wi, wv = np_oracle.topk_canonical(s, k)
    assert np.array_equal(idx, wi)
    assert np.array_equal(_bits(val), _bits(wv))


def test_topk_batched_rows(b2r):
    rng = np.random.default_rng(11)
    s = rng.normal(0, 1, (37, 20_011)).astype(np.float32)
    idx, val = b2r.fast_topk_selection(torch.from_numpy(s).cuda(), 50)
    for r in range(s.shape[0]):
        wi, wv = np_oracle.topk_canonical(s[r], 50)
        assert np.array_equal(idx[r].cpu().numpy(), wi) and np.array_equal(val[r].cpu().numpy(), wv)


# ----------------------------------------------------------------------------------- oracle on seeded inputs
def _oracle_topk(ix_args, q_ptr, q_terms, q_w, k):
    data, indices, indptr, dl, idf, k1, b, avgdl = ix_args
    return c_oracle.bm25_search_batch(q_ptr, q_terms, q_w, len(idf), data, indices, indptr, dl, idf, k1, b, avgdl, k)


@pytest.mark.parametrize("tile_docs,k", [(256, 10), (4096, 10), (4096, 100), (16384, 10)])
def test_bm25_search_vs_oracle_medium(b2r, tile_docs, k):
    from b200ret import synthetic as S
    n_docs, n_vocab = 50_000 + 37, 20_000
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 60, seed=1)
    dl[::1000] = 0; indptr = indptr.copy()                          # (lengths only: rows keep their postings)
    idf = b2r.reference_idf(indices, n_docs, n_vocab)
    avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(96, n_vocab, seed=2)
    u_ptr, u_terms, u_w = S.zipf_queries(32, n_vocab, seed=3, uniform=True)       # rare-term stress set
    q_ptr = np.concatenate([q_ptr, u_ptr[1:] + q_ptr[-1]]).astype(np.int32)
    q_terms = np.concatenate([q_terms, u_terms]); q_w = np.concatenate([q_w, u_w * 2])
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl,
                                     tile_docs=tile_docs)
    idx, val = ix.search(q_ptr, q_terms, q_w, k)
    wi, wv = _oracle_topk((data, indices, indptr, dl, idf, 1.2, 0.75, avgdl), q_ptr, q_terms, q_w, k)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(_bits(val.cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv)))
    hi, hv = ix.search_host(q_ptr, q_terms, q_w, k)                 # host-buffer C-ABI call
    assert np.array_equal(hi, wi) and np.array_equal(_bits(hv), _bits(val.cpu().numpy()))


def test_bank_schedule_is_a_permutation_inside_segments_and_changes_no_result(b2r):
    """The bank schedule of the index builder (pass 6) may only permute postings INSIDE a dense (term, sub-tile)
    segment.  Property checks on the device layout: same multiset of (doc, value) per segment as the
    doc-ascending build; every full row hits different accumulator slots -- 32 postings with 32 different residues
    doc mod 32 whose two halves are each distinct mod 16 (the default schedule: f32 and f64 accumulators), or 16
    postings distinct mod 16 (the round-1 schedule, set_bank_schedule(16)); scores / top-k are bit-identical with
    and without it (BM25 and impact kinds)."""
    from b200ret import synthetic as S
    n_docs, n_vocab, tile = 30_000 + 5, 2000, 1024
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 50, seed=31)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(64, n_vocab, seed=32)
    for kind in ("bm25", "impact"):
        built = {}
        for on in (False, True, 16):
            b2r.set_bank_schedule(on)
            try:
                ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl,
                                                 tile_docs=tile, kind=kind)
            finally:
                b2r.set_bank_schedule(True)
            nnz = len(indices)
            doc = ix._bufs["post_doc"].view(torch.int32)[:nnz].cpu().numpy().astype(np.int64)
            vdt = torch.float64 if kind == "bm25" else torch.float32
            val = ix._bufs["post_val"].view(vdt)[:nnz].cpu().numpy()
            did = ix._bufs["dense_id"].view(torch.int32)[:n_vocab].cpu().numpy()
            n_seg = ix.n_tiles * 8
            dptr = ix._bufs["dense_ptr"].view(torch.int32).cpu().numpy().astype(np.int64)
            scores = ix.score_dense(q_ptr, q_terms, q_w).cpu().numpy()
            top = ix.search(q_ptr, q_terms, q_w, 10)
            built[on] = (doc, val, did, dptr, n_seg, scores, top[0].cpu().numpy(), top[1].cpu().numpy())
        for classes in (32, 16):
          (d0, v0, did0, p0, n_seg, s0, i0, tv0), (d1, v1, did1, p1, _, s1, i1, tv1) = built[False], built[True if classes == 32 else 16]
          assert np.array_equal(did0 >= 0, did1 >= 0)            # (row numbers are handed out by an atomic counter)
          assert np.array_equal(_bits(s0), _bits(s1)) and np.array_equal(i0, i1) and np.array_equal(_bits(tv0), _bits(tv1))
          dense_terms = np.nonzero(did0 >= 0)[0]
          assert len(dense_terms) > 10
          sub, moved, full_rows = tile // 8, 0, 0
          for t in dense_terms[:40]:
              row = p0[did0[t] * (n_seg + 1):(did0[t] + 1) * (n_seg + 1)]
              assert np.array_equal(row, p1[did1[t] * (n_seg + 1):(did1[t] + 1) * (n_seg + 1)])
              for sgm in range(n_seg):
                  lo, hi = row[sgm], row[sgm + 1]
                  a, b_ = d0[lo:hi], d1[lo:hi]
                  assert np.all(np.diff(a) > 0) and (hi == lo or (a[0] >= sgm * sub and a[-1] < (sgm + 1) * sub))
                  o = np.argsort(b_, kind="stable")
                  assert np.array_equal(b_[o], a) and np.array_equal(v1[lo:hi][o], v0[lo:hi])   # same (doc, value) pairs
                  moved += int(not np.array_equal(a, b_))
                  cnt = np.bincount(b_ % classes, minlength=classes)
                  for g in range(int(cnt.min())):                                  # rows no class has run out of
                      r_ = b_[classes * g:classes * g + classes]
                      assert len(set((r_ % classes).tolist())) == classes
                      assert len(set((r_[:16] % 16).tolist())) == 16 and len(set((r_[-16:] % 16).tolist())) == 16
                      full_rows += 1
          assert moved > 0 and full_rows > 0


def test_fused_selection_equals_plain_path_and_survives_overflow(b2r):
    """The fused path (threshold from every 16th tile + candidate lists) must equal the plain path; an
    adversarial corpus whose sample tiles hold no matching document forces the candidate lists to overflow
    and exercises the device-gated exhaustive fallback."""
    from b200ret import synthetic as S
    n_docs, n_vocab, tile = 40_000 + 11, 3000, 256          # 157 tiles -> fused path is active
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 30, seed=21)
    rows = np.repeat(np.arange(n_docs), np.diff(indptr))
    keep = (rows // tile) % 16 != 0                          # empty every sample tile
    data2, indices2 = data[keep], indices[keep]
    indptr2 = np.zeros(n_docs + 1, np.int64); np.cumsum(np.bincount(rows[keep], minlength=n_docs), out=indptr2[1:])
    q_ptr, q_terms, q_w = S.zipf_queries(48, n_vocab, seed=22)
    for (d_, i_, p_) in ((data, indices, indptr), (data2, indices2, indptr2)):
        idf = b2r.reference_idf(i_, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
        ix = b2r.TermMajorIndex.from_csr(d_, i_, p_, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, tile_docs=tile)
        for k in (10, 100):
            b2r.set_fused_selection(True)
            fi, fv = ix.search(q_ptr, q_terms, q_w, k)
            b2r.set_fused_selection(False)
            pi, pv = ix.search(q_ptr, q_terms, q_w, k)
            b2r.set_fused_selection(True)
            assert torch.equal(fi, pi) and torch.equal(fv, pv), k
            # every candidate list overflows (capacity = k): ALL queries take the exhaustive fallback (streaming
            # top-k in shared memory, part lists merged by the last CTA of a query)
            b2r.set_fused_cap(1)
            try:
                xi, xv, xk = ix.search(q_ptr, q_terms, q_w, k, return_keys=True)
                hi_, hv_ = ix.search_host(q_ptr, q_terms, q_w, k)
            finally:
                b2r.set_fused_cap(0)
            assert torch.equal(xi, pi) and torch.equal(xv, pv), k
            assert np.array_equal(hi_, pi.cpu().numpy()) and np.array_equal(_bits(hv_), _bits(pv.cpu().numpy()))
            ku = xk.cpu().numpy().view(np.uint64)
            assert bool((ku[:, :-1] > ku[:, 1:]).all())
            wi, wv = _oracle_topk((d_, i_, p_, dl, idf, 1.2, 0.75, avgdl), q_ptr, q_terms, q_w, k)
            assert np.array_equal(fi.cpu().numpy(), wi), k
            assert np.array_equal(_bits(fv.cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv))), k


def test_bm25_edge_queries(b2r):
    from b200ret import synthetic as S
    n_docs, n_vocab = 3000, 500
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 20, seed=4)
    keep = indices != 7                                             # term 7 never occurs: empty posting list
    rows = np.repeat(np.arange(n_docs), np.diff(indptr))[keep]
    data, indices = data[keep], indices[keep]
    indptr = np.zeros(n_docs + 1, np.int64); np.cumsum(np.bincount(rows, minlength=n_docs), out=indptr[1:])
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    long_q = np.arange(0, 100, dtype=np.int32)                      # > 32 terms: several staging passes
    q_ptr, q_terms, q_w = b2r.pack_queries([([], []), ([7], [1.0]), (long_q, np.ones(100)), ([0], [3.0]),
                                            ([499, 7, 3], [1, 1, 2])])
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, tile_docs=256)
    s = ix.score_dense(q_ptr, q_terms, q_w).cpu().numpy()
    for q in range(len(q_ptr) - 1):
        qtf = np_oracle.dense_query(q_terms[q_ptr[q]:q_ptr[q + 1]], q_w[q_ptr[q]:q_ptr[q + 1]], n_vocab)
        want = c_oracle.bm25_scores(qtf, data, indices, indptr, dl, idf, 1.2, 0.75, avgdl)
        assert np.array_equal(_bits(s[q]), _bits(want)), q
    assert not s[0].any() and not s[1].any()
    idx, val = ix.search(q_ptr, q_terms, q_w, 5)
    assert idx[0].tolist() == [0, 1, 2, 3, 4] and val[0].tolist() == [0.0] * 5       # all-zero scores: lowest ids
    with pytest.raises(ValueError, match="outside"):
        b2r.TermMajorIndex.from_csr(data, indices + 1000, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    dup = indices.copy(); dup[indptr[5] + 1] = dup[indptr[5]]       # row 5 lists one term twice
    with pytest.raises(ValueError, match="twice"):
        b2r.TermMajorIndex.from_csr(data, dup, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, tile_docs=256)


def test_long_queries_walk_several_tiles_per_cta(b2r):
    """Queries of more than 32 terms are re-staged per tile while a CTA walks several doc tiles (the <= 32-term
    queries of the other tests keep their terms in registers across tiles): both against the oracle."""
    from b200ret import synthetic as S
    n_docs, n_vocab, k = 61_000, 4000, 10
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 40, seed=41)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    rng = np.random.default_rng(42)
    qs = [(rng.choice(n_vocab, size=int(rng.integers(33, 70)), replace=False), rng.integers(1, 4, 70)[:0])
          for _ in range(72)]
    qs = [(t_, np.ones(len(t_), np.float32) * (1 + i % 3)) for i, (t_, _) in enumerate(qs)]
    qs[5] = (np.arange(0, 32), np.ones(32, np.float32))              # exactly 32 terms: the staged path
    q_ptr, q_terms, q_w = b2r.pack_queries(qs)
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, tile_docs=256)
    wi, wv = _oracle_topk((data, indices, indptr, dl, idf, 1.2, 0.75, avgdl), q_ptr, q_terms, q_w, k)
    for fused in (True, False):
        b2r.set_fused_selection(fused)
        try:
            idx, val = ix.search(q_ptr, q_terms, q_w, k)
        finally:
            b2r.set_fused_selection(True)
        assert np.array_equal(idx.cpu().numpy(), wi), fused
        assert np.array_equal(_bits(val.cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv))), fused


def test_sharded_merge_equals_single_index(b2r):
    """3 doc shards with global idf/avgdl on one GPU + b2r_merge_candidates == one index over everything."""
    from b200ret import synthetic as S
    from b200ret.dist import shard_range
    n_docs, n_vocab, k = 30_011, 8000, 100
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 40, seed=8)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(64, n_vocab, seed=9)
    full = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, tile_docs=1024)
    fi, fv = full.search(q_ptr, q_terms, q_w, k)
    parts = []
    for r in range(3):
        lo, hi = shard_range(n_docs, 3, r)
        s, e = indptr[lo], indptr[hi]
        sh = b2r.TermMajorIndex.from_csr(data[s:e], indices[s:e], indptr[lo:hi + 1] - s, dl[lo:hi], n_vocab=n_vocab,
                                         idf=idf, avgdl=avgdl, doc_id_base=lo, tile_docs=1024)
        parts.append(sh.search(q_ptr, q_terms, q_w, k, return_keys=True)[2])
    gathered = torch.stack(parts).contiguous()
    nq = gathered.shape[1]
    mi = torch.empty((nq, k), dtype=torch.int64, device="cuda"); mv = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ws = torch.empty(nq * k * 8 + (1 << 20), dtype=torch.uint8, device="cuda")
    lib = b2r._abi.lib
    b2r._abi.check(lib.b2r_merge_candidates(gathered.data_ptr(), 3, nq, k, None, mi.data_ptr(), mv.data_ptr(),
                                            ws.data_ptr(), ws.numel(), int(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(mi, fi) and torch.equal(mv, fv)
    wi, _ = _oracle_topk((data, indices, indptr, dl, idf, 1.2, 0.75, avgdl), q_ptr, q_terms, q_w, k)
    assert np.array_equal(mi.cpu().numpy(), wi)


# ----------------------------------------------------------------------------------- service level
def test_service_text_api_vs_reference(b2r, golden_dir, tmp_path):
    with open(os.path.join(golden_dir, "service_text.json")) as f:
        g = json.load(f)
    path = tmp_path / "docs.idx"
    with pytest.raises(FileNotFoundError):
        b2r.RetrievalService(path)
    b2r.MemoryIndex(path, create=True).close()
    with b2r.RetrievalService(path) as svc:
        with pytest.raises(ValueError, match="not built"):
            svc.search_bm25({"q": "x"})
        with pytest.raises(ValueError, match="Empty corpus"):
            svc.build_bm25_index({})
        svc.build_bm25_index(g["corpus"])
        m = g["meta"]
        assert len(svc.vocabulary) == m["vocab_size"] and svc.avgdl == m["avgdl"] and svc.corpus_tf.nnz == m["nnz"]
        assert float(np.sum(svc.idf_weights.astype(np.float64))) == m["idf_sum"]
        assert sorted(svc.get_stats().keys()) == m["stats_keys"]
        got = svc.search_bm25(g["queries"], top_k=10)
        ix = np_oracle.build_text_index(g["corpus"])
        want = np_oracle.search_bm25_text(ix, g["queries"], top_k=10)
        assert list(got) == list(g["queries"])
        for qid in g["queries"]:
            assert list(got[qid].items()) == list(want[qid].items()), qid            # ids + scores + order
            ref = g["ref_top10"][qid]
            assert sorted(got[qid].values(), reverse=True) == sorted(ref.values(), reverse=True), qid
        assert got["blank"] == {} and got["oov"] == {}
        n_cached = len(svc.query_cache)
        again = svc.search_bm25(g["queries"], top_k=10)                              # served from the cache
        assert again == got and len(svc.query_cache) == n_cached
        got500 = svc.search_bm25({k: g["queries"][k] for k in g["ref_top500"]}, top_k=500)   # top_k > n_docs
        for qid, ref in g["ref_top500"].items():
            assert got500[qid] == ref, qid
        svc.k1 = 0.9                                                                  # attribute change -> re-layout
        changed = svc.search_bm25({"q": g["queries"]["query_0"] + " "}, top_k=5)
        want2 = np_oracle.search_bm25_text(ix, {"q": g["queries"]["query_0"]}, top_k=5, k1=0.9)
        assert list(changed["q"].items()) == list(want2["q"].items())
        svc.clear_cache()
        assert svc.get_stats()["query_cache_size"] == 0


def test_index_file_roundtrip_and_corruption(b2r, tmp_path):
    """save() -> load(): every device buffer comes back bit for bit, searches agree, a re-based shard keeps global
    ids, and a flipped byte / truncated file / foreign file is refused."""
    from b200ret import synthetic as S
    n_docs, n_vocab, k = 20_000 + 7, 3000, 10
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 40, seed=51)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(40, n_vocab, seed=52)
    for kind in ("bm25", "impact"):
        ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl,
                                         tile_docs=512, kind=kind, k1=1.4, b=0.6, doc_id_base=5000)
        path = tmp_path / f"shard_{kind}.b2r"
        size = ix.save(path)
        assert size == os.path.getsize(path) and size % 4096 == 0
        ld = b2r.TermMajorIndex.load(path)
        assert (ld.kind, ld.n_docs, ld.n_vocab, ld.nnz, ld.tile_docs, ld.n_tiles, ld.doc_id_base) == \
               (ix.kind, ix.n_docs, ix.n_vocab, ix.nnz, ix.tile_docs, ix.n_tiles, 5000)
        assert (ld.k1, ld.b, ld.avgdl) == (1.4, 0.6, ix.avgdl)
        for name in ("post_doc", "post_val", "blk_ptr", "dense_id", "dense_ptr", "idf"):
            assert torch.equal(ix._bufs[name].view(torch.uint8).reshape(-1), ld._bufs[name].view(torch.uint8).reshape(-1)), name
        i0, v0 = ix.search(q_ptr, q_terms, q_w, k)
        i1, v1 = ld.search(q_ptr, q_terms, q_w, k)
        assert torch.equal(i0, i1) and torch.equal(v0, v1)
        assert np.array_equal(ld.search_host(q_ptr, q_terms, q_w, k)[0], i0.cpu().numpy())
        rb = b2r.TermMajorIndex.load(path, doc_id_base=0)
        i2, v2 = rb.search(q_ptr, q_terms, q_w, k)
        assert torch.equal(i2 + 5000, i0) and torch.equal(v2, v0)
    raw = bytearray(path.read_bytes())
    bad = tmp_path / "flipped.b2r"
    raw[4096 + 100] ^= 0x40
    bad.write_bytes(raw)
    with pytest.raises(ValueError, match="checksum"):
        b2r.TermMajorIndex.load(bad)
    assert b2r.TermMajorIndex.load(bad, verify=False).n_docs == n_docs          # (checks can be skipped explicitly)
    (tmp_path / "short.b2r").write_bytes(bytes(raw[:len(raw) // 2]))
    with pytest.raises(ValueError, match="exceeds the file"):
        b2r.TermMajorIndex.load(tmp_path / "short.b2r")
    (tmp_path / "foreign.b2r").write_bytes(b"PK\x03\x04" + bytes(8000))
    with pytest.raises(ValueError, match="magic"):
        b2r.TermMajorIndex.load(tmp_path / "foreign.b2r")


def test_service_save_load_and_reference_cache_file(b2r, golden_dir, tmp_path):
    """RetrievalService.save_bm25_index / load_bm25_index: the loaded service answers like the built one without
    running the build kernels; a cache file in the reference's own .npz layout (no .b2r next to it) also loads."""
    g = json.load(open(os.path.join(golden_dir, "service_text.json")))
    store = tmp_path / "docs.idx"
    b2r.MemoryIndex(store, create=True).close()
    svc = b2r.RetrievalService(store)
    svc.build_bm25_index(g["corpus"])
    want = svc.search_bm25(g["queries"], top_k=10)
    svc.save_bm25_index(tmp_path / "bm25.b2r")
    launches = b2r._abi.lib.b2r_launch_count()
    svc2 = b2r.RetrievalService(store)
    svc2.load_bm25_index(tmp_path / "bm25.b2r")
    # no build kernel ran: the only launch is the one that derives the packed copy of the f32 pre-filter
    assert b2r._abi.lib.b2r_launch_count() <= launches + 1
    assert svc2.search_bm25(g["queries"], top_k=10) == want
    assert svc2.vocabulary == svc.vocabulary and svc2.doc_ids == svc.doc_ids and svc2.avgdl == svc.avgdl
    tf = svc.corpus_tf                                                  # reference layout: evaluate_rag_pipeline.py:280-293
    np.savez_compressed(tmp_path / "ref_cache.npz", tf_data=tf.data, tf_indices=tf.indices, tf_indptr=tf.indptr,
                        tf_shape=tf.shape, doc_lengths=svc.doc_lengths, idf=svc.idf_weights,
                        vocabulary=list(svc.vocabulary.keys()), doc_ids=svc.doc_ids, avgdl=svc.avgdl)
    svc3 = b2r.RetrievalService(store)
    svc3.load_bm25_index(tmp_path / "ref_cache.npz")
    assert svc3.search_bm25(g["queries"], top_k=10) == want


def _reference_root():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for root in ("/root/reference", os.path.join(here, "oracle", "_ref")):
        if os.path.isfile(os.path.join(root, "rag_system", "core", "retriever_registry.py")):
            return root
    return None


def test_registry_plugin_shape(b2r, golden_dir):
    with open(os.path.join(golden_dir, "service_text.json")) as f:
        g = json.load(f)

    class Registry:                                   # stand-in with the reference's register/create contract
        _r = {}
        @classmethod
        def register(cls, name, klass): cls._r[name] = klass
        @classmethod
        def create(cls, cfg): return cls._r[cfg["type"]](method=cfg.get("method", "bm25"))
    b2r.register_with(Registry)
    r = Registry.create({"type": "bm25_b200"})
    r.build_index_from_corpus(g["corpus"])
    out = r.search({"a": g["queries"]["query_1"]}, top_k=10)
    assert sorted(out["a"].values(), reverse=True) == sorted(g["ref_top10"]["query_1"].values(), reverse=True)


@pytest.mark.skipif(_reference_root() is None, reason="reference package not available (no /root/reference, no oracle/_ref)")
def test_plugin_in_the_reference_registry_vs_the_reference_retriever(b2r, tmp_path):
    """f2 end to end: both retrievers come out of the REFERENCE's RetrieverRegistry.create -- its own
    OptimizedBM25Retriever (Numba, CPU) and the registered B200 plugin -- are built on the same text corpus and
    answer the same query set.  Same scores in the same order for every query (ids wherever scores are not tied).
    Then the pipeline hook: prefetch() scores the whole query set in ONE GPU call and the pipeline-style loop of
    <= 100-query search() calls that follows launches no kernel; and the reference's own index cache file
    (evaluate_rag_pipeline.py:277-293) loads into the plugin."""
    import sys
    from b200ret import synthetic as S
    root = _reference_root()
    sys.path.insert(0, root)
    try:
        from b200ret import retriever as plug
        from rag_system.core import retriever_registry as rr
        reg = plug.install()
        corpus = S.fiqa_shape_corpus(num_docs=6000, avg_doc_length=60, vocab_size=8000, seed=7)
        queries = S.fiqa_shape_queries(num_queries=260, avg_query_length=6, vocab_size=8000, seed=8)
        queries["blank"] = ""
        queries["oov"] = "zzzz qqqq"
        ref = reg.create({"type": "bm25", "params": {"k1": 1.4, "b": 0.6}})
        mine = reg.create({"type": "bm25_b200", "params": {"k1": 1.4, "b": 0.6}})
        assert type(ref).__name__ == "OptimizedBM25Retriever" and isinstance(mine, b2r.B200BM25Retriever)
        ref.build_index_from_corpus(corpus)
        mine.build_index_from_corpus(corpus)
        assert mine.vocabulary == ref.vocabulary and mine.doc_ids == ref.doc_ids and mine.avgdl == ref.avgdl
        assert np.array_equal(mine.idf_weights, ref.idf_weights)
        want = ref.search(queries, top_k=20)
        got = mine.search(queries, top_k=20)
        assert list(got) == list(want)
        n_nonempty = 0
        for qid in queries:
            wv, gv = list(want[qid].values()), list(got[qid].values())
            assert gv == wv, qid
            if len(set(wv)) == len(wv) and len(wv) < 20:       # no ties, and no tie at the cut either
                assert list(got[qid]) == list(want[qid]), qid
            n_nonempty += bool(wv)
        assert n_nonempty > 200 and got["blank"] == {} and got["oov"] == {}
        # the pipeline hook
        mine.clear_cache()
        lib = b2r._abi.lib
        assert mine.prefetch({q: {"text": t} for q, t in queries.items()}, top_k=20) >= 200
        n0 = lib.b2r_launch_count()
        items = list(queries.items())
        merged = {}
        for i in range(0, len(items), 100):                    # evaluate_rag_pipeline.py:741-780
            merged.update(mine.search(dict(items[i:i + 100]), top_k=20))
        assert lib.b2r_launch_count() == n0 and merged == got
        # a cache file with the reference's keys, object arrays included (what np.savez makes of its Python lists)
        cache = tmp_path / "bm25_index_cafe.npz"
        tf = ref.corpus_tf
        np.savez_compressed(cache, tf_data=tf.data, tf_indices=tf.indices, tf_indptr=tf.indptr, tf_shape=tf.shape,
                            doc_lengths=ref.doc_lengths, idf=ref.idf_weights, vocabulary=list(ref.vocabulary.keys()),
                            doc_ids=ref.doc_ids, avgdl=ref.avgdl)
        again = reg.create({"type": "bm25_b200", "params": {"k1": 1.4, "b": 0.6}})
        again.load_cached_index(cache)
        assert again.search(queries, top_k=20) == got
        rr.RetrieverRegistry._retrievers.pop("bm25_b200", None)
    finally:
        sys.path.remove(root)


# ----------------------------------------------------------------------------------- full size (C2) properties
def test_full_size_1m_docs_properties(b2r):
    """BASELINE config 2 shape (1M docs x 100K vocab, 1024 queries, top-10): spot-check 6 queries against
    the oracle, and check size-independent properties on all 1024: descending order, values equal the dense
    score at the returned index, sharded == unsharded."""
    from b200ret import synthetic as S
    n_docs, n_vocab, k = 1_000_000, 100_000, 10
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 60)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(1024, n_vocab)
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    idx, val, keys = ix.search(q_ptr, q_terms, q_w, k, return_keys=True)
    torch.cuda.synchronize()
    ku = keys.cpu().numpy().view(np.uint64)
    assert bool((val[:, :-1] >= val[:, 1:]).all()) and bool((ku[:, :-1] > ku[:, 1:]).all())
    sub = slice(0, 6)
    wi, wv = _oracle_topk((data, indices, indptr, dl, idf, 1.2, 0.75, avgdl), q_ptr[:7], q_terms[:q_ptr[6]],
                          q_w[:q_ptr[6]], k)
    assert np.array_equal(idx[sub].cpu().numpy(), wi) and np.array_equal(_bits(val[sub].cpu().numpy()), _bits(wv))
    dense = ix.score_dense(q_ptr[:9], q_terms[:q_ptr[8]], q_w[:q_ptr[8]])
    assert torch.equal(torch.gather(dense, 1, idx[:8]), val[:8])
    # two shards + merge == whole
    half = n_docs // 2
    parts = []
    for lo, hi in ((0, half), (half, n_docs)):
        s, e = indptr[lo], indptr[hi]
        sh = b2r.TermMajorIndex.from_csr(data[s:e], indices[s:e], indptr[lo:hi + 1] - s, dl[lo:hi], n_vocab=n_vocab,
                                         idf=idf, avgdl=avgdl, doc_id_base=lo)
        parts.append(sh.search(q_ptr, q_terms, q_w, k, return_keys=True)[2])
        del sh
    g = torch.stack(parts).contiguous()
    mi = torch.empty_like(idx); mv = torch.empty_like(val)
    ws = torch.empty(1 << 22, dtype=torch.uint8, device="cuda")
    b2r._abi.check(b2r._abi.lib.b2r_merge_candidates(g.data_ptr(), 2, 1024, k, None, mi.data_ptr(), mv.data_ptr(),
                                                     ws.data_ptr(), ws.numel(),
                                                     int(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(mi, idx) and torch.equal(mv, val)


# ----------------------------------------------------------------------------------- peer-memory exchange + merge
@pytest.mark.parametrize("world,nq,k", [(3, 300, 10), (2, 1024, 100), (4, 7, 1)])
def test_exchange_merge_kernel_emulated_ranks_on_one_gpu(b2r, world, nq, k):
    """b2r_exchange_merge (one launch per rank: push my ranked keys into every rank's receive buffer, wait chunk by
    chunk, merge by rank counting).  The ranks are emulated on ONE GPU: `world` receive buffers in plain device
    memory, one stream per rank so that the kernels run side by side and can wait for each other; repeated calls
    exercise the epoch / parity double buffering.  Expected = top-k of the union of the ranks' lists."""
    lib = b2r._abi.lib
    dev = torch.device("cuda")
    rng = np.random.default_rng(100 + world)
    nbytes = int(lib.b2r_exchange_bytes(world, nq, k))
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    ptrs = torch.tensor([b_.data_ptr() for b_ in bufs], dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for rep in range(4):
        # distinct keys (score bits << 32 | ~doc); every rank's list ranked descending, some lists short (0-padded)
        lists = np.zeros((world, nq, k), np.uint64)
        for r in range(world):
            sc = rng.integers(1, 1 << 20, (nq, k)).astype(np.uint64)
            doc = (rng.permutation(nq * k).reshape(nq, k) + r * nq * k).astype(np.uint64)
            key = (sc << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - doc)
            key = -np.sort(-key.astype(np.int64), axis=1)
            key = key.astype(np.uint64)
            short = rng.random(nq) < 0.2
            cut = rng.integers(0, k + 1, nq)
            for q in np.nonzero(short)[0]:
                key[q, cut[q]:] = 0
            lists[r] = key
        local = [torch.from_numpy(lists[r].view(np.int64)).to(dev) for r in range(world)]
        outs = []
        torch.cuda.synchronize()
        for r in range(world):
            ko = torch.empty((nq, k), dtype=torch.int64, device=dev)
            io = torch.empty((nq, k), dtype=torch.int64, device=dev)
            vo = torch.empty((nq, k), dtype=torch.float32, device=dev)
            with torch.cuda.stream(streams[r]):
                b2r._abi.check(lib.b2r_exchange_merge(local[r].data_ptr(), ptrs.data_ptr(), r, world, nq, k, ko.data_ptr(),
                                                      io.data_ptr(), vo.data_ptr(), int(streams[r].cuda_stream)))
            outs.append((ko, io, vo))
        torch.cuda.synchronize()
        for r in range(world):
            b2r._abi.check(lib.b2r_exchange_status(bufs[r].data_ptr(), None), f"rank {r}")
        union = np.concatenate([lists[r] for r in range(world)], axis=1)
        want = -np.sort(-union.view(np.int64), axis=1)[:, :k]
        want_u = want.view(np.uint64)
        for r in range(world):
            ko, io, vo = (x.cpu().numpy() for x in outs[r])
            assert np.array_equal(ko.view(np.uint64), want_u), (rep, r)
            wi = np.where(want_u != 0, (np.uint64(0xFFFFFFFF) - (want_u & np.uint64(0xFFFFFFFF))).astype(np.int64), -1)
            assert np.array_equal(io, wi), (rep, r)
            assert np.all(np.isneginf(vo[want_u == 0]))


# ----------------------------------------------------------------------------------- batches in flight / threads
def test_batch_pipeline_and_concurrent_host_searches(b2r):
    """(1) dist.BatchPipeline: two batches in flight on two streams inside one captured graph give, lane by lane,
    what a plain search gives, replay after replay, also after new queries were written into the captured inputs.
    (2) RetrievalService-level thread safety (the reference's search_bm25 may be called from several threads): four
    threads issue host-buffer searches on ONE index at the same time; every result equals the serial one."""
    import threading
    from b200ret import synthetic as S
    from b200ret.dist import BatchPipeline
    n_docs, n_vocab, k = 90_000, 9000, 10
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 50, seed=71)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    sets = [S.zipf_queries(200, n_vocab, seed=72 + i) for i in range(4)]
    want = [tuple(t.clone() for t in ix.search(*q, k)) for q in sets]
    # (1) same shapes are needed for the captured inputs: pad every set's term arrays to the longest
    n_t = max(len(q[1]) for q in sets)
    dev = torch.device("cuda")
    d_ptr = torch.from_numpy(sets[0][0]).to(dev)
    d_terms = torch.zeros(n_t, dtype=torch.int32, device=dev)
    d_w = torch.zeros(n_t, dtype=torch.float32, device=dev)
    d_terms[:len(sets[0][1])] = torch.from_numpy(sets[0][1]).to(dev)
    d_w[:len(sets[0][2])] = torch.from_numpy(sets[0][2]).to(dev)
    pipe = BatchPipeline(ix, depth=2)
    outs = pipe.capture(d_ptr, d_terms, d_w, k)
    for rep, qi in enumerate((0, 0, 1, 2, 1)):
        q = sets[qi]
        d_ptr.copy_(torch.from_numpy(q[0]))
        d_terms.zero_(); d_w.zero_()
        d_terms[:len(q[1])] = torch.from_numpy(q[1]).to(dev)
        d_w[:len(q[2])] = torch.from_numpy(q[2]).to(dev)
        pipe.replay()
        torch.cuda.synchronize()
        for li, lv in outs:
            assert torch.equal(li, want[qi][0]) and torch.equal(lv, want[qi][1]), (rep, qi)
    pipe.close()
    # (2)
    res, errs = {}, []

    def worker(i):
        try:
            for r in range(6):
                q = sets[(i + r) % 4]
                hi, hv = ix.search_host(q[0], q[1], q[2], k)
                res[(i, r)] = (hi, hv, (i + r) % 4)
        except Exception as ex:          # pragma: no cover
            errs.append(ex)
    th = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t_ in th: t_.start()
    for t_ in th: t_.join()
    assert not errs and len(res) == 24
    for (i, r), (hi, hv, qi) in res.items():
        assert np.array_equal(hi, want[qi][0].cpu().numpy()) and np.array_equal(_bits(hv), _bits(want[qi][1].cpu().numpy())), (i, r)


def test_drop_in_kernels_accept_a_query_vector_shorter_than_the_vocabulary(b2r):
    """The reference skips CSR term ids >= len(query_tf) (`if term_idx < len(query_tf)`, retrieval.py:66): a query
    vector shorter than the vocabulary must work on the drop-in simd_bm25_score / simd_tfidf_score too (the index is
    built over every term id of the CSR; the tail terms can simply never be queried)."""
    from b200ret import synthetic as S
    n_docs, n_vocab = 5000, 900
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 30, seed=81)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    rng = np.random.default_rng(82)
    for n_short in (300, 1):
        qtf = np.zeros(n_short, np.float32)
        qtf[rng.integers(0, n_short, min(5, n_short))] = 1.0
        qtf[0] = 2.0
        want = c_oracle.bm25_scores(qtf, data, indices, indptr, dl, idf, 1.2, 0.75, avgdl)
        got = b2r.simd_bm25_score(qtf, data, indices, indptr.astype(np.int32), dl, idf, 1.2, 0.75, avgdl)
        assert np.array_equal(_bits(got), _bits(want)), n_short
        want3 = c_oracle.tfidf_scores(qtf, data, indices, indptr, idf)
        got3 = b2r.simd_tfidf_score(qtf, data, indices, indptr.astype(np.int32), idf)
        assert np.array_equal(_bits(got3), _bits(want3)), n_short
    b2r.clear_index_cache()


# ----------------------------------------------------------------------------------- f32 pre-filter of the search path
def test_approx_prefilter_equals_f64_path_and_oracle(b2r, tmp_path):
    """The search path's f32 pre-filter (packed 4-byte postings, f32 accumulators, exact f64 rescoring of the survivors;
    csrc/score_approx.cu) must return bit for bit what the f64-only path and the oracle return: Zipfian queries (normal
    mode, negative idf on the head terms), rare-term queries (positive mode: few matches), long queries (> 32 terms),
    an empty query, absurd weights (no finite error bound -> exhaustive fallback), fractional tf / lengths, several k,
    a shard with a document-id base, forced list overflow, and an index that went through save() / load()."""
    from b200ret import synthetic as S
    n_docs, n_vocab = 120_000 + 123, 6000                   # 30 tiles of 4096 -> the fused path is active
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 50, seed=71)
    rng = np.random.default_rng(72)
    data = (data * rng.uniform(0.25, 1.75, len(data))).astype(np.float32)       # fractional term frequencies
    dl = (dl * rng.uniform(0.5, 1.5, len(dl))).astype(np.float32)
    idf = b2r.reference_idf(indices, n_docs, n_vocab); avgdl = b2r.reference_avgdl(dl)
    assert (idf[:5] < 0).all()                               # head terms occur in most documents: negative idf
    z_ptr, z_terms, z_w = S.zipf_queries(150, n_vocab, seed=73)
    u_ptr, u_terms, u_w = S.zipf_queries(40, n_vocab, seed=74, uniform=True)
    qs = [(z_terms[z_ptr[i]:z_ptr[i + 1]], z_w[z_ptr[i]:z_ptr[i + 1]] * (1 + i % 3)) for i in range(150)]
    qs += [(u_terms[u_ptr[i]:u_ptr[i + 1]], u_w[u_ptr[i]:u_ptr[i + 1]] * 0.5) for i in range(40)]
    qs += [(rng.choice(n_vocab, size=int(rng.integers(33, 80)), replace=False), np.ones(80, np.float32)[:0])
           for _ in range(6)]
    qs = [(t_, w_ if len(w_) == len(t_) else np.ones(len(t_), np.float32)) for t_, w_ in qs]
    qs += [([], []),                                          # no terms: all-zero scores, lowest ids
           ([0, 1, 2], [1.0, 1.0, 1.0]),                      # only negative-idf terms: every touched score <= 0
           ([3, 900, 4100], [1e30, 1e30, 1e30]),              # no finite bound: exact fallback
           ([5, 1200, 5000], [1e-30, 1e-30, 1e-30]),          # tiny weights
           ([n_vocab - 1], [2.0])]                            # one rare term
    q_ptr, q_terms, q_w = b2r.pack_queries(qs)
    args = (data, indices, indptr, dl, idf, 1.2, 0.75, avgdl)
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    assert "post_pk" in ix._bufs and ix.prefilter_u_max is not None and 0 < ix.prefilter_u_max < 2.3
    path = str(tmp_path / "ix.b2r")
    ix.save(path)
    loaded = b2r.TermMajorIndex.load(path)
    assert "post_pk" in loaded._bufs                          # derived data: rebuilt on load, never stored
    for k in (1, 10, 100):
        wi, wv = _oracle_topk(args, q_ptr, q_terms, q_w, k)
        wv = np.where(wv == 0, np.float32(0), wv)
        ai, av, ak = ix.search(q_ptr, q_terms, q_w, k, return_keys=True)
        b2r.set_approx_prefilter(False)
        try:
            fi, fv, fk = ix.search(q_ptr, q_terms, q_w, k, return_keys=True)
        finally:
            b2r.set_approx_prefilter(True)
        bad = np.flatnonzero((ai != fi).any(1).cpu().numpy())
        assert torch.equal(ai, fi) and torch.equal(av, fv) and torch.equal(ak, fk), (k, bad[:10])
        assert np.array_equal(ai.cpu().numpy(), wi), k
        assert np.array_equal(_bits(av.cpu().numpy()), _bits(wv)), k
        li, lv = loaded.search(q_ptr, q_terms, q_w, k)
        assert torch.equal(li, ai) and torch.equal(lv, av), k
        hi_, hv_ = ix.search_host(q_ptr, q_terms, q_w, k)
        assert np.array_equal(hi_, wi) and np.array_equal(_bits(hv_), _bits(wv)), k
    # forced overflow: every query goes through the exhaustive f64 fallback behind the pre-filter
    k = 10
    wi, wv = _oracle_topk(args, q_ptr, q_terms, q_w, k)
    b2r.set_fused_cap(1)
    try:
        xi, xv = ix.search(q_ptr, q_terms, q_w, k)
    finally:
        b2r.set_fused_cap(0)
    assert np.array_equal(xi.cpu().numpy(), wi) and np.array_equal(_bits(xv.cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv)))
    # a shard with a document-id base: keys carry global ids
    lo, hi = 40_000, n_docs
    s, e = indptr[lo], indptr[hi]
    sh = b2r.TermMajorIndex.from_csr(data[s:e], indices[s:e], indptr[lo:hi + 1] - s, dl[lo:hi], n_vocab=n_vocab,
                                     idf=idf, avgdl=avgdl, doc_id_base=lo)
    si, sv = sh.search(q_ptr, q_terms, q_w, k)
    b2r.set_approx_prefilter(False)
    try:
        ti, tv = sh.search(q_ptr, q_terms, q_w, k)
    finally:
        b2r.set_approx_prefilter(True)
    assert torch.equal(si, ti) and torch.equal(sv, tv) and int(si.min()) >= lo


def test_approx_prefilter_mass_ties_and_non_finite_values(b2r):
    """Identical documents (thousands of equal scores around the k-th best: more survivors than the rescoring stage
    holds -> exhaustive fallback, ties broken by document index) and an index whose saturation factors are not finite
    (doc length NaN: the packed copy is refused and the f64 path serves the index)."""
    n_docs, n_vocab, k = 60_000, 50, 10
    indptr = np.arange(0, 2 * n_docs + 1, 2, dtype=np.int64)
    indices = np.tile(np.array([3, 7], np.int32), n_docs)
    data = np.ones(2 * n_docs, np.float32)
    data[2 * 31_000] = 2.0                                   # one document stands out
    dl = np.full(n_docs, 2.0, np.float32)
    idf = np.linspace(0.5, 2.0, n_vocab).astype(np.float32); avgdl = 2.0
    q_ptr, q_terms, q_w = b2r.pack_queries([([3, 7], [1.0, 1.0]), ([3], [1.0]), ([9], [1.0])])
    ix = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    assert "post_pk" in ix._bufs
    idx, val = ix.search(q_ptr, q_terms, q_w, k)
    wi, wv = _oracle_topk((data, indices, indptr, dl, idf, 1.2, 0.75, avgdl), q_ptr, q_terms, q_w, k)
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(_bits(val.cpu().numpy()), _bits(np.where(wv == 0, np.float32(0), wv)))
    assert idx[0, 0].item() == 31_000 and idx[0, 1:].tolist() == list(range(9))
    dl2 = dl.copy(); dl2[17] = np.nan
    ix2 = b2r.TermMajorIndex.from_csr(data, indices, indptr, dl2, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
    assert "post_pk" not in ix2._bufs and not ix2._desc.post_pk
    i2, v2 = ix2.search(q_ptr, q_terms, q_w, k)
    b2r.set_fused_selection(False)
    try:
        p2, pv2 = ix2.search(q_ptr, q_terms, q_w, k)
    finally:
        b2r.set_fused_selection(True)
    assert torch.equal(i2, p2) and np.array_equal(_bits(v2.cpu().numpy()), _bits(pv2.cpu().numpy()))


def test_approx_prefilter_long_candidate_lists():
    """Candidate lists of more than 4096 keys (k = 100 with a 1/64 threshold sample: the selection kernels then need
    dynamic shared memory above the default limit): pre-filter on == f64 path.  B2R_SAMPLE_STEP is read when the
    library is loaded, so the check runs in its own process (tools/check_bigcap.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "check_bigcap.py")], capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert p.returncode == 0 and "prefilter == f64 path: True" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    assert " 8192 " in p.stdout          # the plan really used a long list
