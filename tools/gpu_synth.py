"""Synthetic corpora generated ON THE GPU with torch (data generation only, never on a timed path): the laws of
b200ret/synthetic.py at sizes numpy cannot produce in a test's time budget (8.8M documents).
Different RNG stream than the numpy generators, same distributions.  Used by tests/test_gpu_fullsize.py,
tools/bench_configs.py and bench.py's `secondary` measurements."""
import numpy as np
import torch


def zipf_csr_torch(n_docs, n_vocab, mean_len, seed, dev, distinct_per_doc=None, chunk=1 << 20):
    """Doc-major CSR (data f32, indices i32 sorted per row, indptr i64, doc_lengths f32 | None) on `dev`.
    distinct_per_doc=None: BM25 shape (doc length ~ clip(floor(Gamma(2, mean/2)), 5, 400), tokens Zipf(s=1),
    tf = counts).  distinct_per_doc=D: SPLADE shape (the first D distinct Zipf terms of 3*D draws per document,
    weights Gamma(2, 0.5))."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    p = 1.0 / torch.arange(1, n_vocab + 1, dtype=torch.float64, device=dev)
    cdf = torch.cumsum(p / p.sum(), 0)
    datas, inds, nnz_rows, lens_all = [], [], [], []
    for lo in range(0, n_docs, chunk):
        n = min(chunk, n_docs - lo)
        if distinct_per_doc is None:
            gam = torch.distributions.Gamma(torch.tensor(2.0, device=dev), torch.tensor(2.0 / mean_len, device=dev))
            torch.manual_seed(seed + lo)
            lens = torch.clamp(torch.floor(gam.sample((n,))), 5, 400).to(torch.int64)
        else:
            lens = torch.full((n,), distinct_per_doc * 3, dtype=torch.int64, device=dev)
        tot = int(lens.sum())
        toks = torch.clamp(torch.searchsorted(cdf, torch.rand(tot, device=dev, generator=g, dtype=torch.float64)),
                           max=n_vocab - 1)
        doc = torch.repeat_interleave(torch.arange(n, device=dev), lens)
        key, _ = torch.sort(doc * n_vocab + toks)
        del toks, doc
        uniq, cnt = torch.unique_consecutive(key, return_counts=True)
        del key
        rows = uniq // n_vocab
        if distinct_per_doc is not None:      # keep the first `distinct_per_doc` terms of every row
            start = torch.searchsorted(rows, torch.arange(n, device=dev))
            rank = torch.arange(len(uniq), device=dev) - start[rows]
            keep = rank < distinct_per_doc
            uniq, cnt, rows = uniq[keep], cnt[keep], rows[keep]
        nnz_rows.append(torch.bincount(rows, minlength=n))
        inds.append((uniq % n_vocab).to(torch.int32))
        if distinct_per_doc is None:
            datas.append(cnt.to(torch.float32))
            lens_all.append(lens.to(torch.float32))
        else:
            gam = torch.distributions.Gamma(torch.tensor(2.0, device=dev), torch.tensor(2.0, device=dev))
            datas.append(gam.sample((len(uniq),)).to(torch.float32))
    indptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.cat(nnz_rows), 0, out=indptr[1:])
    return torch.cat(datas), torch.cat(inds), indptr, (torch.cat(lens_all) if lens_all else None)


def global_bm25_stats(indices, doc_lengths, n_docs, n_vocab):
    """idf f32[V] / avgdl with the reference's host expressions (retrieval.py:187-190) from device arrays."""
    df = torch.bincount(indices, minlength=n_vocab).cpu().numpy()
    idf = np.log((n_docs - df + 0.5) / (df + 0.5)).astype(np.float32)
    avgdl = float(np.mean(doc_lengths.cpu().numpy()))
    return df, idf, avgdl


def random_int8_corpus(n, dim, seed, dev):
    """Uniform INT8 vectors + positive f32 scales (timing and exactness of the INT8 scan do not depend on the law)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    d8 = torch.randint(-127, 128, (n, dim), device=dev, dtype=torch.int8, generator=g)
    ds = torch.rand(n, device=dev, generator=g) + 0.01
    return d8, ds


def random_int8_queries(nq, dim, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q8 = torch.randint(-127, 128, (nq, dim), device=dev, dtype=torch.int8, generator=g)
    qs = (torch.rand(nq, device=dev, generator=g) + 0.01) / 127
    return q8, qs
