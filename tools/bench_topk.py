#!/usr/bin/env python
"""Row a3 on its own: b2r_topk (the fast_topk_selection drop-in, rag_system/core/retrieval.py:79-92) over a
[rows, n] f32 score matrix resident in HBM.  Prints one JSON line per (n, k): ms, GB/s of score bytes, fraction
of the measured HBM peak, parity of row 0 against torch.topk values.

    python tools/bench_topk.py [--rows 1024] [--n 1000000] [--k 10 100]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ret  # noqa: E402

PEAK = 6540.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--n", type=int, nargs="+", default=[1_000_000])
    ap.add_argument("--k", type=int, nargs="+", default=[10, 100])
    args = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(7)
    for n in args.n:
        s = torch.randn((args.rows, n), device=dev, generator=g)
        for k in args.k:
            for _ in range(3):
                idx, val = b200ret.fast_topk_selection(s, k)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                idx, val = b200ret.fast_topk_selection(s, k)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            gbs = args.rows * n * 4 / (ms * 1e-3) / 1e9
            ok = bool(torch.equal(val[0], torch.topk(s[0], k, sorted=True).values))
            print(json.dumps({"op": "fast_topk_selection", "rows": args.rows, "n": n, "k": k, "ms": ms, "score_gbs": gbs,
                              "frac_of_hbm_peak": gbs / PEAK, "row0_values_match_torch_topk": ok,
                              "note": "includes the workspace/out allocations of the Python wrapper"}), flush=True)


if __name__ == "__main__":
    main()
