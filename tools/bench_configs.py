#!/usr/bin/env python
"""Measurements of the BASELINE.json configurations other than the headline one (they are parity-test
cases, not bench lines): C3 (8.8M docs, top-100), C4 (SPLADE-shape impact dot), C5 (INT8 768-d scan).
Synthetic corpora are generated on the GPU with torch (data generation only); every timed call goes
through libb200ret.  Prints one JSON object per configuration.

    python tools/bench_configs.py int8 [--docs N] [--queries Q ...]
    python tools/bench_configs.py c3   [--docs N]
    python tools/bench_configs.py c4   [--docs N]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import b200ret  # noqa: E402
from b200ret import synthetic as S  # noqa: E402

PEAK = 6540.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


from gpu_synth import zipf_csr_torch  # noqa: E402


def run_int8(args):
    dev = torch.device("cuda")
    n, dim, k = args.docs, 768, args.k
    g = torch.Generator(device=dev); g.manual_seed(42)
    d8 = torch.randint(-127, 128, (n, dim), device=dev, dtype=torch.int8, generator=g)
    ds = torch.rand(n, device=dev, generator=g) + 0.01
    out = []
    if args.pair is not None:
        b200ret.set_int8_pair(bool(args.pair))
    for nq, cl, fm in [(a, b, c) for a in args.queries for b in (args.clusters or [None])
                       for c in (args.fused_modes or [None])]:
        if cl is not None:
            b200ret.set_int8_cluster(cl)
        if fm is not None:
            b200ret.set_int8_fused(fm)
        q8 = torch.randint(-127, 128, (nq, dim), device=dev, dtype=torch.int8, generator=g)
        qs = (torch.rand(nq, device=dev, generator=g) + 0.01) / 127
        ms = timed(lambda: b200ret.int8_scan_topk(q8, d8, qs, ds, k), steps=3, warmup=1)
        ops = 2.0 * nq * n * dim
        rec = {"config": f"c5: INT8 {dim}-d exhaustive scan, {n} vectors, top-{k}", "queries": nq, "ms": ms,
               "queries_per_s": nq / (ms * 1e-3), "int8_tops": ops / (ms * 1e-3) / 1e12,
               "corpus_gbs": n * (dim + 4) / (ms * 1e-3) / 1e9}
        if cl is not None:
            rec["max_cluster"] = cl
        if fm is not None:
            rec["fused_mode"] = fm
        if args.pair is not None:
            rec["cta_pair"] = args.pair
        if nq <= 64 and n <= 2_000_000:    # parity spot check on the first query against exact integer math
            idx, val, _ = b200ret.int8_scan_topk(q8[:1], d8, qs[:1], ds, k)
            dots = (d8.to(torch.int32) * q8[0].to(torch.int32)).sum(1).to(torch.float64)
            sc = ((dots * qs[0].double()) * ds.double()).float()
            top = torch.topk(sc, k, sorted=True)
            rec["top_values_match"] = bool(torch.equal(top.values, val[0]))
        out.append(rec)
        print(json.dumps(rec), flush=True)
    return out


def run_sparse(args, kind):
    dev = torch.device("cuda")
    if kind == "c3":
        n_docs, n_vocab, k, nq = args.docs, 100_000, 100, 1024
        t0 = time.time()
        data, ind, ptr, dl = zipf_csr_torch(n_docs, n_vocab, 60.0, 20260101, dev)
        df = torch.bincount(ind, minlength=n_vocab).cpu().numpy()
        idf = np.log((n_docs - df + 0.5) / (df + 0.5)).astype(np.float32)
        avgdl = float(np.mean(dl.cpu().numpy()))
        q_ptr, q_terms, q_w = S.zipf_queries(nq, n_vocab)
        ix = b200ret.TermMajorIndex.from_csr(data, ind, ptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
        per = 12
        desc = f"c3: synthetic Zipfian {n_docs} docs x 100K vocab, 1024 queries, BM25 top-{k}, 1 GPU"
    else:
        n_docs, n_vocab, k, nq = args.docs, 30522, 10, 256
        t0 = time.time()
        data, ind, ptr, _ = zipf_csr_torch(n_docs, n_vocab, 0, 20260103, dev, distinct_per_doc=120, chunk=1 << 18)
        df = torch.bincount(ind, minlength=n_vocab).cpu().numpy()
        idf = np.ones(n_vocab, np.float32)
        q_ptr, q_terms, q_w = S.impact_queries(nq, n_vocab, 30)
        ix = b200ret.TermMajorIndex.from_csr(data, ind, ptr, None, n_vocab=n_vocab, idf=idf, kind="impact")
        per = 8
        desc = f"c4: SPLADE-shape {n_docs} docs x 30522 vocab, ~120 nnz/doc, 30-term queries, impact dot top-{k}"
    torch.cuda.synchronize()
    build_s = time.time() - t0
    nnz = int(ptr[-1])
    postings = int(df[q_terms].sum())
    d_ptr, d_t, d_w = (torch.from_numpy(a).to(dev) for a in (q_ptr, q_terms, q_w))
    ms = timed(lambda: ix.search(d_ptr, d_t, d_w, k), steps=5, warmup=2)
    b200ret.set_fused_selection(False)
    ms_plain = timed(lambda: ix.search(d_ptr, d_t, d_w, k), steps=3, warmup=1)
    b200ret.set_fused_selection(True)
    alg = per * postings + 8 * nq * n_docs
    rec = {"config": desc, "nnz": nnz, "index_gb": ix.device_bytes() / 1e9, "gen_plus_build_s": build_s,
           "postings_touched": postings, "ms_per_batch": ms, "queries_per_s": nq / (ms * 1e-3),
           "ms_per_batch_plain_path": ms_plain, "algorithmic_gbs": alg / (ms * 1e-3) / 1e9,
           "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / PEAK}
    # parity spot check: first 3 queries against the oracle on the host copy of the CSR
    if args.check:
        from oracle import c_oracle
        c = 3
        h = [t.cpu().numpy() for t in (data, ind, ptr)]
        idx, val = ix.search(q_ptr[:c + 1], q_terms[:q_ptr[c]], q_w[:q_ptr[c]], k)
        ok = True
        for q in range(c):
            qtf = np.zeros(n_vocab, np.float32)
            qtf[q_terms[q_ptr[q]:q_ptr[q + 1]]] = q_w[q_ptr[q]:q_ptr[q + 1]]
            if kind == "c3":
                s = c_oracle.bm25_scores(qtf, h[0], h[1], h[2], dl.cpu().numpy(), idf, 1.2, 0.75, avgdl)
            else:
                s = c_oracle.tfidf_scores(qtf, h[0], h[1], h[2], idf)
            wi, wv = c_oracle.topk(s, k)
            ok &= bool(np.array_equal(idx[q].cpu().numpy(), wi) and
                       np.array_equal(val[q].cpu().numpy().view(np.uint32),
                                      np.where(wv == 0, np.float32(0), wv).view(np.uint32)))
        rec["bit_exact_vs_oracle_first_3_queries"] = ok
    print(json.dumps(rec), flush=True)
    return rec


def run_c3_sharded(args):
    """Config 3 as named: 8.8M docs doc-sharded over the ranks of this torchrun job, BM25 top-100, NCCL all-gather
    of the per-shard candidates + merge kernel.  Every rank generates its own shard on its GPU."""
    import torch.distributed as dist
    from b200ret.dist import ShardedBM25, shard_range
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    n_docs, n_vocab, k, nq = args.docs, 100_000, 100, 1024
    lo, hi = shard_range(n_docs, world, rank)

    def shard(r):
        a, b = shard_range(n_docs, world, r)
        return zipf_csr_torch(b - a, n_vocab, 60.0, 20260101 + 7919 * r, dev)
    data, ind, ptr, dl = shard(rank)
    df = torch.bincount(ind, minlength=n_vocab)
    tot = torch.stack([dl.double().sum(), torch.tensor(float(hi - lo), dtype=torch.float64, device=dev)])
    dist.all_reduce(df)
    dist.all_reduce(tot)
    df_h = df.cpu().numpy()
    idf = np.log((n_docs - df_h + 0.5) / (df_h + 0.5)).astype(np.float32)
    avgdl = float(np.float32(float(tot[0]) / float(tot[1])))
    ix = b200ret.TermMajorIndex.from_csr(data, ind, ptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl, doc_id_base=lo)
    sh = ShardedBM25(ix)
    q_ptr, q_terms, q_w = S.zipf_queries(nq, n_vocab)
    d_ptr, d_t, d_w = (torch.from_numpy(a).to(dev) for a in (q_ptr, q_terms, q_w))
    for _ in range(3):
        sh.search(d_ptr, d_t, d_w, k)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    a.record()
    for _ in range(steps):
        idx, val = sh.search(d_ptr, d_t, d_w, k)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    postings = int(df_h[q_terms].sum())
    rec = None
    if rank == 0:
        alg = 12 * postings + 8 * nq * n_docs
        rec = {"config": f"c3: synthetic Zipfian {n_docs} docs x 100K vocab doc-sharded over {world} B200, 1024 queries, "
                         f"BM25 top-{k}, NCCL all-gather + merge", "n_gpus": world, "ms_per_batch": float(ms),
               "queries_per_s": nq / (float(ms) * 1e-3), "postings_touched_global": postings,
               "algorithmic_gbs_all_gpus": alg / (float(ms) * 1e-3) / 1e9,
               "frac_of_hbm_peak_per_gpu": alg / (float(ms) * 1e-3) / 1e9 / PEAK / world}
        if args.check:      # rebuild the whole corpus on the host (same seeds) and check 2 queries against the oracle
            from oracle import c_oracle
            parts = [shard(r) for r in range(world)]
            h_data = np.concatenate([p[0].cpu().numpy() for p in parts])
            h_ind = np.concatenate([p[1].cpu().numpy() for p in parts])
            h_dl = np.concatenate([p[3].cpu().numpy() for p in parts])
            offs = np.cumsum([0] + [int(p[2][-1]) for p in parts])
            h_ptr = np.concatenate([parts[0][2].cpu().numpy()[:1]] +
                                   [p[2].cpu().numpy()[1:] + offs[i] for i, p in enumerate(parts)])
            del parts
            ok = True
            for q in range(2):
                qtf = np.zeros(n_vocab, np.float32)
                qtf[q_terms[q_ptr[q]:q_ptr[q + 1]]] = q_w[q_ptr[q]:q_ptr[q + 1]]
                s_ = c_oracle.bm25_scores(qtf, h_data, h_ind, h_ptr, h_dl, idf, 1.2, 0.75, avgdl)
                wi, wv = c_oracle.topk(s_, k)
                ok &= bool(np.array_equal(idx[q].cpu().numpy(), wi) and
                           np.array_equal(val[q].cpu().numpy().view(np.uint32), wv.view(np.uint32)))
            rec["bit_exact_vs_oracle_first_2_queries"] = ok
        print(json.dumps(rec), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


def run_int8_sharded(args):
    """Config 5 as named: the INT8 vectors doc-sharded over the ranks of this torchrun job, exhaustive scan with
    fused top-100 per rank, NCCL all-gather of the candidates + merge kernel (dist.sharded_int8_scan)."""
    import torch.distributed as dist
    from b200ret.dist import shard_range, sharded_int8_scan
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    n, dim, k = args.docs, 768, args.k
    lo, hi = shard_range(n, world, rank)
    g = torch.Generator(device=dev); g.manual_seed(42 + 7919 * rank)
    d8 = torch.randint(-127, 128, (hi - lo, dim), device=dev, dtype=torch.int8, generator=g)
    ds = torch.rand(hi - lo, device=dev, generator=g) + 0.01
    gq = torch.Generator(device=dev); gq.manual_seed(43)              # the same queries on every rank
    for nq in args.queries:
        q8 = torch.randint(-127, 128, (nq, dim), device=dev, dtype=torch.int8, generator=gq)
        qs = (torch.rand(nq, device=dev, generator=gq) + 0.01) / 127
        for _ in range(2):
            sharded_int8_scan(q8, d8, qs, ds, k, lo)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        a.record()
        for _ in range(steps):
            idx, val = sharded_int8_scan(q8, d8, qs, ds, k, lo)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # parity: the merged top-k of query 0 equals the top-k of the union of every rank's exact scores
        dots = (d8.to(torch.int32) * q8[0].to(torch.int32)).sum(1).to(torch.float64)
        sc = ((dots * qs[0].double()) * ds.double()).float()
        loc = torch.topk(sc, k, sorted=True)
        allv = [torch.empty_like(loc.values) for _ in range(world)]
        dist.all_gather(allv, loc.values.contiguous())
        want = torch.sort(torch.cat(allv), descending=True).values[:k]
        if rank == 0:
            ms_ = float(ms)
            print(json.dumps({"config": f"c5: INT8 {dim}-d exhaustive scan, {n} vectors doc-sharded over {world} B200, "
                                        f"top-{k}, NCCL all-gather + merge", "n_gpus": world, "queries": nq, "ms": ms_,
                              "queries_per_s": nq / (ms_ * 1e-3), "int8_tops_all_gpus": 2.0 * nq * n * dim / (ms_ * 1e-3) / 1e12,
                              "corpus_gbs_all_gpus": n * (dim + 4) / (ms_ * 1e-3) / 1e9,
                              "top_values_match": bool(torch.equal(want, val[0]))}), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


def run_dense(args):
    """a8: RetrievalService.search_by_vector's fp32 gemv + top-k (retrieval.py:402-436) on a resident embedding matrix:
    b2r_f32_dot_topk at N x 768 f32, Q = 1 / 8 / 64 queries, top-10; bytes = the matrix once per pass of 8 queries."""
    dev = torch.device("cuda")
    n, dim, k = args.docs, 768, 10
    g = torch.Generator(device=dev); g.manual_seed(5)
    emb = torch.randn((n, dim), device=dev, generator=g, dtype=torch.float32)
    for nq in (1, 8, 64):
        q = torch.randn((nq, dim), device=dev, generator=g, dtype=torch.float32)
        ms = timed(lambda: b200ret.dense_topk(emb, q, k), steps=5, warmup=2)
        passes = (nq + 7) // 8
        rec = {"config": f"a8: fp32 {dim}-d gemv + top-{k}, {n} vectors resident in HBM", "queries": nq, "ms": ms,
               "queries_per_s": nq / (ms * 1e-3), "matrix_gbs": passes * n * dim * 4 / (ms * 1e-3) / 1e9,
               "frac_of_hbm_peak": passes * n * dim * 4 / (ms * 1e-3) / 1e9 / PEAK}
        idx, val = b200ret.dense_topk(emb, q[:1], k)
        want = torch.topk(emb @ q[0], k)
        rec["top_ids_match_torch"] = bool(torch.equal(torch.sort(idx[0]).values, torch.sort(want.indices).values))
        print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["int8", "c3", "c4", "dense"])
    ap.add_argument("--docs", type=int, default=None)
    ap.add_argument("--queries", type=int, nargs="+", default=[1, 64, 1024])
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--k", type=int, default=100, help="int8: top-k")
    ap.add_argument("--pair", type=int, default=None, help="int8: 1 = CTA-pair (cta_group::2) fused scan, 0 = single CTA")
    ap.add_argument("--fused-modes", type=int, nargs="*", default=None,
                    help="int8: sweep b2r_set_int8_fused (0 plain, 1 default gate, 2 fused for every batch size)")
    ap.add_argument("--clusters", type=int, nargs="*", default=None,
                    help="int8: sweep the thread-block cluster cap of the tcgen05 kernel (A/B measurement)")
    args = ap.parse_args()
    if args.docs is None:
        args.docs = {"int8": 2_000_000, "c3": 8_800_000, "c4": 2_200_000, "dense": 2_000_000}[args.what]
    if args.what == "c3" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_c3_sharded(args)
    elif args.what == "int8" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_int8_sharded(args)
    elif args.what == "int8":
        run_int8(args)
    elif args.what == "dense":
        run_dense(args)
    else:
        run_sparse(args, args.what)


if __name__ == "__main__":
    main()
