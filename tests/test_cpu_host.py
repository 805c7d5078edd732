"""CPU-side checks: the C-ABI library loads and exports every declared symbol (no compute calls),
host query packing follows the reference's dense query_tf semantics, the document-store shim
round-trips the reference's file format, and the doc-sharding plumbing works under gloo (world 2)."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import b200ret
    header = open(os.path.join(ROOT, "include", "b200ret.h")).read()
    declared = set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", header))
    declared -= {"b2r_status", "b2r_kind", "b2r_index", "b2r_index_sizes"}
    assert len(declared) >= 15
    lib = b200ret._abi.lib
    for name in sorted(declared):
        assert hasattr(lib, name), f"libb200ret.so does not export {name}"
        assert name in b200ret._abi.SIGNATURES, f"_abi.py does not bind {name}"
    assert lib.b2r_version() == 100


def test_sizes_and_argument_errors_without_a_gpu():
    import ctypes as C
    import b200ret
    lib = b200ret._abi.lib
    sz = b200ret._abi.B2RIndexSizes()
    assert lib.b2r_index_sizes_for(1000, 100, 50, 4096, 0, C.byref(sz)) == 0
    assert sz.post_doc_bytes >= 4000 and sz.post_val_bytes >= 8000 and sz.blk_ptr_bytes >= 51 * 4
    assert lib.b2r_index_sizes_for(1000, 100, 50, 1000, 0, C.byref(sz)) == -1      # tile not a power of two
    assert b"tile_docs" in lib.b2r_last_error()
    with pytest.raises(ValueError):
        b200ret._abi.check(lib.b2r_index_sizes_for(-1, 100, 50, 4096, 0, C.byref(sz)), "sizes")
    n = C.c_size_t(0)
    assert lib.b2r_topk_workspace(4, 100000, 10, C.byref(n)) == 0 and n.value > 0


def test_no_cpu_fallback_when_cuda_is_absent():
    import b200ret
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ret.fast_topk_selection(np.arange(10, dtype=np.float32), 3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


def test_index_file_header_layout_and_validation():
    """Host-only part of the on-disk index format (include/b200ret.h): layout, checksum, every rejection."""
    import ctypes as C
    import b200ret
    from b200ret import _abi
    lib = _abi.lib
    assert C.sizeof(_abi.B2RIndexFileHeader) <= _abi.FILE_ALIGN

    def fresh():
        h = _abi.B2RIndexFileHeader()
        h.n_docs, h.nnz, h.n_vocab, h.tile_docs, h.kind, h.n_dense_max = 10_000, 123_456, 777, 1024, 0, 5
        h.k1, h.b, h.avgdl = 1.2, 0.75, 12.5
        total = C.c_uint64(0)
        assert lib.b2r_index_file_layout(C.byref(h), C.byref(total)) == 0
        return h, total.value

    h, total = fresh()
    assert h.magic == b"B2RIDX01" and h.version == 1 and h.header_bytes == 4096 and h.n_tiles == 10 and h.subtiles == 8
    sizes = _abi.B2RIndexSizes()
    assert lib.b2r_index_sizes_for(123_456, 10_000, 777, 1024, 0, C.byref(sizes)) == 0
    want = [sizes.post_doc_bytes, sizes.post_val_bytes, sizes.blk_ptr_bytes, sizes.dense_id_bytes, None, 777 * 4]
    end = 4096
    for i, sec in enumerate(h.sections):
        assert sec.offset % 4096 == 0 and sec.offset >= end and (want[i] is None or sec.bytes == want[i])
        end = sec.offset + sec.bytes
    assert end <= total and total % 4096 == 0
    assert lib.b2r_index_file_check(C.byref(h), total) == 0

    def rejected(mutate, size=None):
        g, tot = fresh()
        mutate(g)
        rc = lib.b2r_index_file_check(C.byref(g), tot if size is None else size)
        return rc == -5 and len(lib.b2r_last_error()) > 0

    assert rejected(lambda g: setattr(g, "magic", b"NOTANIDX"))
    assert rejected(lambda g: setattr(g, "version", 2))
    assert rejected(lambda g: setattr(g, "n_tiles", 11))
    assert rejected(lambda g: setattr(g, "tile_docs", 1000))                 # not a power of two
    assert rejected(lambda g: setattr(g.sections[1], "bytes", g.sections[1].bytes - 8))
    assert rejected(lambda g: setattr(g.sections[2], "offset", g.sections[2].offset + 16))
    assert rejected(lambda g: None, size=total - 4096)                       # truncated file
    assert rejected(lambda g: setattr(g, "doc_id_base", 2 ** 32))

    a = np.arange(4099, dtype=np.uint8)
    c0 = lib.b2r_checksum64(a.ctypes.data, a.size)
    assert c0 == lib.b2r_checksum64(a.copy().ctypes.data, a.size)
    assert c0 != lib.b2r_checksum64(a.ctypes.data, a.size - 1)
    b = a.copy(); b[4000] ^= 1
    assert c0 != lib.b2r_checksum64(b.ctypes.data, b.size)
    assert lib.b2r_checksum64(None, 0) == lib.b2r_checksum64(a.ctypes.data, 0)


def test_pack_queries_matches_dense_query_tf():
    import b200ret
    ptr, terms, w = b200ret.pack_queries([([5, 2, 5, 9], [1.0, 2.0, 3.0, 0.0]), ([], []), ([7], [-1.0]), ([1, 0], [4, 5])])
    assert ptr.tolist() == [0, 2, 2, 2, 4]
    assert terms.tolist() == [2, 5, 0, 1] and w.tolist() == [2.0, 3.0, 5.0, 4.0]   # last write wins, w>0, ascending
    q = np.zeros((2, 6), np.float32)
    q[0, 3] = 2; q[1, 1] = 1; q[1, 4] = -3
    ptr, terms, w = b200ret.queries_from_dense(q)
    assert ptr.tolist() == [0, 1, 2] and terms.tolist() == [3, 1] and w.tolist() == [2.0, 1.0]


def test_reference_host_expressions(golden_dir):
    import b200ret
    z = np.load(os.path.join(golden_dir, "bm25_arrays.npz"))
    idf = b200ret.reference_idf(z["indices"], len(z["indptr"]) - 1, len(z["idf"]))
    assert np.array_equal(idf, z["idf"])
    assert b200ret.reference_avgdl(z["doc_lengths"]) == float(z["avgdl"])


def test_docstore_roundtrip(tmp_path):
    import b200ret
    p = tmp_path / "docs.idx"
    with pytest.raises(FileNotFoundError):
        b200ret.MemoryIndex(p)
    ix = b200ret.MemoryIndex(p, create=True)
    assert ix.get_document_count() == 0
    docs = [b200ret.Document(id=f"d{i}", text=("word " * (5 + 80 * (i % 2))).strip(), title=f"T{i}",
                             metadata={"i": i}) for i in range(7)]
    ix.add_documents(docs[:4])
    ix.add_documents(docs[4:])
    ix.close()
    ix = b200ret.MemoryIndex(p)
    assert ix.get_document_count() == 7
    d = ix.get_document("d3")
    assert d.text == docs[3].text and d.title == "T3" and d.metadata == {"i": 3}
    assert ix.get_document("nope") is None
    assert [x.id for x in ix.get_documents(["d6", "d0"])] == ["d6", "d0"]
    assert ix.get_document_ids() == [f"d{i}" for i in range(7)] and ix.contains("d2") and not ix.contains("zz")
    st = ix.get_index_stats()                                   # keys of memory_index.py:482-499
    assert sorted(st) == ["average_doc_size_bytes", "cache_stats", "compression_enabled", "file_size_mb",
                          "memory_mapped", "num_documents"]
    assert st["num_documents"] == 7 and st["memory_mapped"] and st["file_size_mb"] > 0
    ix.close()


def test_docstore_reads_reference_layout(tmp_path):
    """A file written byte-by-byte in the reference's layout (memory_index.py:22-34,159-195)."""
    import pickle, struct, zlib
    import b200ret
    text = ("alpha beta " * 60).encode()
    ztext = zlib.compress(text, 6)
    meta = pickle.dumps({"k": 1})
    blob = struct.pack("QQQB", 2, len(ztext), 1, 1) + b"x1" + ztext + b"T" + struct.pack("Q", len(meta)) + meta
    p = tmp_path / "ref.idx"
    p.write_bytes(struct.pack("QQI", 1, len(blob), 64) + blob)
    ix = b200ret.MemoryIndex(p)
    d = ix.get_document("x1")
    assert d.text == text.decode() and d.title == "T" and d.metadata == {"k": 1}
    ix.close()


def test_shard_range_partitions_everything():
    from b200ret.dist import shard_range
    for n, w in [(10, 3), (8, 8), (5, 8), (1_000_000, 8), (8_800_000, 4)]:
        got = [shard_range(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))


def _gloo_worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from b200ret.dist import gather_candidates, global_statistics, shard_range
    from oracle import np_oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "bm25_arrays.npz"))
    n_docs, n_vocab = len(z["indptr"]) - 1, len(z["idf"])
    lo, hi = shard_range(n_docs, world, rank)
    s, e = z["indptr"][lo], z["indptr"][hi]
    idf, avgdl, n_glob = global_statistics(z["indices"][s:e], z["doc_lengths"][lo:hi], n_vocab)
    assert n_glob == n_docs
    assert np.array_equal(idf, z["idf"]), "global idf must equal the single-index idf"
    assert avgdl == float(z["avgdl"])
    # per-shard top-k of the reference scores -> keys -> gather -> merged top-k == global top-k
    k = 10
    scores = z["ref_scores"]
    keys = np.zeros((scores.shape[0], k), np.uint64)
    for q in range(scores.shape[0]):
        idx, val = np_oracle.topk_canonical(scores[q, lo:hi], k)
        u = val.view(np.uint32).astype(np.uint64)
        u = np.where(val == 0, np.uint64(0), u)
        o = np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
        keys[q, :len(idx)] = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (idx + lo).astype(np.uint64))
    g = gather_candidates(torch.from_numpy(keys.view(np.int64)))
    assert tuple(g.shape) == (world, scores.shape[0], k)
    allk = g.numpy().view(np.uint64)
    for q in range(scores.shape[0]):
        merged = np.sort(allk[:, q, :].reshape(-1))[::-1][:k]
        ids = (np.uint64(0xFFFFFFFF) - (merged & np.uint64(0xFFFFFFFF))).astype(np.int64)
        want, _ = np_oracle.topk_canonical(scores[q], k)
        assert np.array_equal(ids, want), (rank, q)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_sharded_plumbing_gloo_world2(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_product_quantisers_match_the_reference_outputs(golden_dir):
    """b200ret.synthetic.quantize_corpus mirrors _quantize_embeddings (retriever_registry.py:435-447): checked against
    the int8 codes / scales the reference itself produced (tests/golden/int8.npz, written by oracle/gen_golden.py)."""
    from b200ret import synthetic as S
    z = np.load(os.path.join(golden_dir, "int8.npz"))
    q, sc = S.quantize_corpus(z["emb"])
    assert q.dtype == np.int8 and sc.dtype == np.float32
    assert np.array_equal(q, z["d8"][:len(q)]) and np.array_equal(sc, z["dscale"][:len(q)])
    q2, s2 = S.quantize_queries(z["emb"][:4])
    assert q2.dtype == np.int8 and np.abs(q2).max() == 127 and np.allclose(q2 * s2[:, None], z["emb"][:4], atol=s2.max())


# ----------------------------------------------------------------------------------- f2: the reference's own registry
def _reference_root():
    """Where the reference package can be imported from: the mounted checkout, or the files oracle/make_ref.py placed
    under oracle/_ref (they travel to the GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for root in ("/root/reference", os.path.join(here, "oracle", "_ref")):
        if os.path.isfile(os.path.join(root, "rag_system", "core", "retriever_registry.py")):
            return root
    return None


@pytest.mark.skipif(_reference_root() is None, reason="reference package not available (no /root/reference, no oracle/_ref)")
def test_plugin_registers_into_the_reference_registry():
    """The plugin goes into the REFERENCE's RetrieverRegistry (rag_system/core/retriever_registry.py:562-599), is
    created by its create() with the config shape the reference uses, and has the constructor / method signatures of
    the class it replaces (OptimizedBM25Retriever, :120-262).  No CUDA call happens before an index is built."""
    import inspect
    import sys
    root = _reference_root()
    sys.path.insert(0, root)
    try:
        import b200ret
        from b200ret import retriever as plug
        from rag_system.core import retriever_registry as rr
        stock_create = rr.RetrieverRegistry.__dict__["create"]
        try:
            reg = plug.install()
            assert reg is rr.RetrieverRegistry
            assert "bm25_b200" in reg.list_available()["registered_custom"]
            r = reg.create({"type": "bm25_b200", "params": {"k1": 0.9, "b": 0.4}})
            assert isinstance(r, b200ret.B200BM25Retriever) and (r.k1, r.b) == (0.9, 0.4)
            with pytest.raises(ValueError, match="Index not built"):
                r.search({"q": "anything"}, top_k=5)
            with pytest.raises(ValueError, match="Empty corpus"):
                r.build_index_from_corpus({})
            ref_cls = rr.OptimizedBM25Retriever
            for name in ("__init__", "build_index_from_corpus", "search", "clear_cache"):
                want = list(inspect.signature(getattr(ref_cls, name)).parameters)
                got = list(inspect.signature(getattr(b200ret.B200BM25Retriever, name)).parameters)
                assert got == want, (name, got, want)
            # a maintainer's switch-over: the built-in names route to the plugin, everything else stays stock
            plug.install(take_over_bm25=True)
            t = reg.create({"type": "tfidf"})
            assert isinstance(t, b200ret.B200BM25Retriever) and (t.k1, t.b) == (1000.0, 0.0)
            assert isinstance(reg.create("bm25"), b200ret.B200BM25Retriever)
            with pytest.raises(ValueError, match="Unknown retriever"):
                reg.create({"type": "no_such_method"})
            # the pipeline's constructor contract (evaluate_rag_pipeline.py:165-180)
            p = b200ret.B200BM25Retriever.from_pipeline_config({"type": "bm25", "params": {"k1": 1.5, "top_k": 50}},
                                                               {"cores": 8, "memory_gb": 64})
            assert (p.k1, p.b, p.method) == (1.5, 0.75, "bm25")
            assert p.cache_file_for({"d2": {}, "d1": {}}).name.startswith("bm25_index_")
        finally:
            rr.RetrieverRegistry.create = stock_create
            rr.RetrieverRegistry._retrievers.pop("bm25_b200", None)
    finally:
        sys.path.remove(root)


# ----------------------------------------------------------------------------------- f4: batched document fetch
def test_batched_document_fetch_matches_per_id_fetch(tmp_path):
    """docstore.MemoryIndex.get_documents / RetrievalService.get_documents / get_search_results / fetch_results
    (reference: rag_system/core/retrieval.py:356-462, memory_index.py:413-468): request order, duplicates, unknown
    ids, compressed and uncompressed records, the threaded path (>= 64 documents) and the service-level cache."""
    import b200ret
    rng = np.random.default_rng(5)
    path = tmp_path / "docs.idx"
    store = b200ret.MemoryIndex(path, create=True)
    docs = []
    for i in range(300):
        n = int(rng.integers(1, 400))                     # short texts stay raw, long ones are zlib-compressed
        docs.append(b200ret.Document(id=f"d{i}", text=" ".join(f"w{int(x)}" for x in rng.integers(0, 50, n)),
                                     title=f"title {i}" if i % 3 else "", metadata={"i": i} if i % 2 else {}))
    store.add_documents(docs)
    ids = [f"d{int(i)}" for i in rng.integers(0, 300, 500)] + ["nope", "d7", "d7"]
    got = store.get_documents(ids, num_workers=4)
    one_by_one = [store.get_document(d) for d in ids]
    assert got == one_by_one and got[500] is None and got[501] == got[502] == docs[7]
    assert store.get_documents(ids[:10], num_workers=1) == one_by_one[:10]
    assert store.get_documents([]) == []
    store.close()
    with b200ret.RetrievalService(path, cache_size=50) as svc:
        assert svc.get_documents(ids) == one_by_one          # more distinct documents than the cache holds
        assert len(svc._cache) <= 50
        res = [{"doc_id": "d3", "score": 2.5}, {"doc_id": "nope", "score": 1.0}, {"doc_id": "d4", "score": 0.5}]
        rows = svc.get_search_results(res)
        assert [r["id"] for r in rows] == ["d3", "d4"] and rows[0]["text"] == docs[3].text and rows[0]["score"] == 2.5
        assert set(svc.get_search_results(res, include_text=False)[0]) == {"id", "score"}
        batch = svc.fetch_results({"q1": {"d10": 3.0, "d11": 2.0}, "q2": {}, "q3": {"d11": 9.0, "zzz": 1.0}})
        assert [r["id"] for r in batch["q1"]] == ["d10", "d11"] and batch["q2"] == [] and len(batch["q3"]) == 1
        assert batch["q3"][0]["metadata"] == docs[11].metadata


def test_prefilter_error_bound_holds_on_an_emulated_corpus():
    """The f32 pre-filter of the search path (csrc/score_approx.cu) is only allowed to rule documents out under a proven
    bound on |approximate - exact|.  Emulate it in numpy on a seeded Zipfian corpus -- values rounded to the packed
    format's 20 bits, f32 weights, f32 accumulation (in a DIFFERENT term order than the exact chain) -- and check the
    three facts the kernels rely on, for Zipfian and rare-term queries, with the constants of approx_bound_warp:
      1. g(approx) <= exact <= h(approx) for every document (g, h: the interval of the relative bound);
      2. the candidate filter derived from the sample threshold keeps every member of the exact top-k;
      3. the survivor filter derived from the k-th best approximate score keeps every member too.
    (No GPU, no library call: this pins the arithmetic of the bound, the GPU tests pin the kernels.)"""
    from oracle import np_oracle
    sys.path.insert(0, os.path.join(ROOT, "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "b2r_synthetic", os.path.join(ROOT, "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200",
                                      "synthetic.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    n_docs, n_vocab, k, k1, b = 120_000, 6000, 10, 1.2, 0.75
    data, ind, ptr, dl = S.zipf_corpus(n_docs, n_vocab, 50, seed=71)
    idf = np_oracle.idf_from_csr(ind, n_docs, n_vocab)
    avgdl = np_oracle.avgdl_from_lengths(dl)
    assert (idf[:5] < 0).all()
    rows = np.repeat(np.arange(n_docs), np.diff(ptr))
    u = (data.astype(np.float64) * (k1 + 1.0)) / (data.astype(np.float64) + k1 * (1.0 - b + b * dl.astype(np.float64)[rows] / avgdl))
    bits = u.astype(np.float32).view(np.uint32)
    packed = ((bits + np.uint32(0x800)) & np.uint32(0xFFFFF000)).view(np.float32)     # pack_postings_kernel
    u_max = float(np.abs(packed).max()) * 1.0009765625
    order = np.argsort(ind, kind="stable")                                             # term-major view
    t_ptr = np.zeros(n_vocab + 1, np.int64)
    np.cumsum(np.bincount(ind, minlength=n_vocab), out=t_ptr[1:])
    t_doc, t_u, t_pk = rows[order], u[order], packed[order]
    z = S.zipf_queries(24, n_vocab, seed=73)
    r = S.zipf_queries(8, n_vocab, seed=74, uniform=True)
    queries = [(z[1][z[0][i]:z[0][i + 1]], z[2][z[0][i]:z[0][i + 1]]) for i in range(24)]
    queries += [(r[1][r[0][i]:r[0][i + 1]], r[2][r[0][i]:r[0][i + 1]]) for i in range(8)]
    queries += [(np.array([0, 1, 2, 5]), np.ones(4, np.float32))]                      # head terms only: low scores
    sample = np.concatenate([np.arange(t0 * 4096, min(n_docs, (t0 + 1) * 4096)) for t0 in range(0, -(-n_docs // 4096), 16)])
    n_pos_mode = 0
    for terms, w in queries:
        exact = np.zeros(n_docs)
        approx = np.zeros(n_docs, np.float32)
        s_all = s_neg = 0.0
        for t, qw in zip(terms, w):                                                    # exact: ascending term id
            sl = slice(t_ptr[t], t_ptr[t + 1])
            exact[t_doc[sl]] += (np.float64(idf[t]) * t_u[sl]) * np.float64(qw)
        for t, qw in zip(terms[::-1], w[::-1]):                                        # approximate: any order
            sl = slice(t_ptr[t], t_ptr[t + 1])
            wq = np.float64(idf[t]) * np.float64(qw)
            approx[t_doc[sl]] += np.float32(wq) * t_pk[sl]
            s_all += abs(wq)
            s_neg += abs(wq) if wq < 0 else 0.0
        n = len(terms)
        delta = 2.0 ** -12 + (n + 16) * 2.0 ** -23                                     # approx_bound_warp
        dp = delta / (1.0 - delta) * 1.000001
        c2 = 2.0 * dp * (s_neg * u_max * 1.0000001) + s_all * 2.0 ** -135 + (n + 1) * 2.0 ** -140
        a64 = approx.astype(np.float64)
        assert (a64 - dp * np.abs(a64) - c2 <= exact).all() and (exact <= a64 + dp * np.abs(a64) + c2).all()
        f32 = exact.astype(np.float32)
        top = np.lexsort((np.arange(n_docs), -f32))[:k]
        grp = approx[sample[:len(sample) // 16 * 16]].reshape(-1, 16).max(1)
        ta = float(np.sort(grp)[-k])
        L = (ta - dp * ta - c2) * (1.0 - 2.0 ** -22)
        pos_mode = not (ta > 0 and L > c2)
        lo = -(c2 / (1.0 - dp)) * 1.000001 - 2.0 ** -130 if pos_mode else (L - c2) / (1.0 + dp) * (1.0 - 2.0 ** -20)
        touched = np.zeros(n_docs, bool)
        for t in terms:
            touched[t_doc[t_ptr[t]:t_ptr[t + 1]]] = True
        cand = np.flatnonzero(touched & (approx >= np.float32(lo)))
        n_pos_mode += pos_mode
        if pos_mode and not (len(cand) >= k and f32[top[-1]] > 0):
            continue                                                                   # the kernels hand these to the exact fallback
        assert set(top.tolist()) <= set(cand.tolist())
        a_k = float(np.sort(approx[cand])[-k])
        G = a_k - dp * abs(a_k) - c2
        yv = G - abs(G) * 2.0 ** -22 - 2.0 ** -130
        x0 = (yv - c2) / (1.0 + dp) if yv >= c2 else (yv - c2) / (1.0 - dp)
        x0 -= abs(x0) * 2.0 ** -20
        surv = cand[approx[cand] >= np.float32(x0)]
        assert set(top.tolist()) <= set(surv.tolist()) and len(surv) <= k + 64
    assert n_pos_mode >= 1
