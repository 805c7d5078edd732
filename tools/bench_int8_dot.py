#!/usr/bin/env python
"""Row a7 on its own: b2r_int8_dot_batch (the quantized_dot_product_batch drop-in,
rag_system/core/retriever_registry.py:90-117) writing the full f32 [Q, N] matrix.  One JSON line per shape.

    python tools/bench_int8_dot.py [--queries 256] [--docs 1000000]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ret  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, nargs="+", default=[256])
    ap.add_argument("--docs", type=int, default=1_000_000)
    args = ap.parse_args()
    dev, dim = torch.device("cuda"), 768
    g = torch.Generator(device=dev); g.manual_seed(3)
    d8 = torch.randint(-127, 128, (args.docs, dim), device=dev, dtype=torch.int8, generator=g)
    ds = torch.rand(args.docs, device=dev, generator=g) + 0.01
    for nq in args.queries:
        q8 = torch.randint(-127, 128, (nq, dim), device=dev, dtype=torch.int8, generator=g)
        qs = (torch.rand(nq, device=dev, generator=g) + 0.01) / 127
        for _ in range(2):
            out = b200ret.quantized_dot_product_batch(q8, d8, qs, ds)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            out = b200ret.quantized_dot_product_batch(q8, d8, qs, ds)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        dots = d8[:4096].to(torch.float64) @ q8[0].to(torch.float64)        # exact: |dot| < 2^24
        ref = ((dots * qs[0].double()) * ds[:4096].double()).float()
        print(json.dumps({"op": "quantized_dot_product_batch", "queries": nq, "docs": args.docs, "dim": dim, "ms": ms,
                          "int8_tops": 2.0 * nq * args.docs * dim / (ms * 1e-3) / 1e12,
                          "output_gbs": nq * args.docs * 4 / (ms * 1e-3) / 1e9,
                          "first_4096_bit_exact": bool(torch.equal(out[0, :4096], ref)),
                          "note": "includes the allocation of the [Q, N] output by the Python wrapper"}), flush=True)


if __name__ == "__main__":
    main()
