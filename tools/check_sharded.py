"""Multi-process parity check of the doc-sharded search (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_sharded.py [--docs 300000] [--queries 512] [--k 10]

Every rank builds its shard of a seeded Zipfian corpus with GLOBAL idf / avgdl from dist.global_statistics (all-reduce),
runs ShardedBM25.search with the peer-memory exchange kernel AND with the NCCL all-gather fallback, plus
dist.BatchPipeline (two batches in flight); rank 0 compares ALL queries (ids and scores, bit for bit) with the oracle
over the whole corpus.  Prints one JSON line; exit code 1 on any mismatch.  tests/test_gpu_multi.py runs it when the
box has at least two GPUs."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=300_000)
    ap.add_argument("--vocab", type=int, default=50_000)
    ap.add_argument("--queries", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    args = ap.parse_args()
    import b200ret
    from b200ret import synthetic as S
    from b200ret.dist import BatchPipeline, ShardedBM25, global_statistics, shard_range
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    data, indices, indptr, dl = S.zipf_corpus(args.docs, args.vocab, 60, seed=91)
    q_ptr, q_terms, q_w = S.zipf_queries(args.queries, args.vocab, seed=92)
    lo, hi = shard_range(args.docs, world, rank)
    s, e = indptr[lo], indptr[hi]
    idf, avgdl, n_glob = global_statistics(indices[s:e], dl[lo:hi], args.vocab, device=dev)
    ok = {"n_docs_global": n_glob == args.docs}
    ix = b200ret.TermMajorIndex.from_csr(data[s:e], indices[s:e], indptr[lo:hi + 1] - s, dl[lo:hi], n_vocab=args.vocab,
                                         idf=idf, avgdl=avgdl, doc_id_base=lo)
    out = {}
    for mode in ("peer", "nccl"):
        os.environ["B2R_EXCHANGE"] = mode
        sh = ShardedBM25(ix)
        for _ in range(3):
            i_, v_ = sh.search(q_ptr, q_terms, q_w, args.k)
        torch.cuda.synchronize()
        out[mode] = (i_.cpu().numpy(), v_.cpu().numpy(), sh.exchange)
        if sh._peer is not None:
            sh._peer.check()
    os.environ["B2R_EXCHANGE"] = "peer"
    d_in = [torch.from_numpy(a).to(dev) for a in (q_ptr, q_terms, q_w)]
    pipe = BatchPipeline(ix, depth=2)
    lanes = pipe.capture(*d_in, args.k)
    for _ in range(3):
        pipe.replay()
    torch.cuda.synchronize()
    pipe.check()
    out["pipeline"] = [(a.cpu().numpy(), b.cpu().numpy()) for a, b in lanes]
    pipe.close()
    rc = 0
    if rank == 0:
        from oracle import c_oracle
        c_oracle.use_all_host_threads()
        idf_ref = b200ret.reference_idf(indices, args.docs, args.vocab)
        ok["global_idf_equals_reference_idf"] = bool(np.array_equal(idf, idf_ref))
        wi, wv = c_oracle.bm25_search_batch(q_ptr, q_terms, q_w, args.vocab, data, indices, indptr, dl, idf, 1.2, 0.75,
                                            avgdl, args.k)
        wv = np.where(wv == 0, np.float32(0), wv)
        same = lambda i_, v_: bool(np.array_equal(i_, wi) and np.array_equal(v_.view(np.uint32), wv.view(np.uint32)))  # noqa: E731
        ok["peer_exchange"] = same(out["peer"][0], out["peer"][1])
        ok["nccl_all_gather"] = same(out["nccl"][0], out["nccl"][1])
        ok["pipeline_lanes"] = all(same(a, b) for a, b in out["pipeline"])
        rc = 0 if all(ok.values()) else 1
        print(json.dumps({"world": world, "queries_checked": args.queries, "k": args.k, "exchange_used": out["peer"][2],
                          "fallback_used": out["nccl"][2], **ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
