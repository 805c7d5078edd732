#!/usr/bin/env python
"""Headline benchmark: BM25 top-10 queries/sec at 1M docs (BASELINE.json config 2) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path (score 1024 queries against the corpus, select top-10, and for
N > 1 all-gather + merge the per-shard candidates).  The 1M-document corpus is doc-sharded over the
N ranks (strong scaling: the corpus is fixed, as the metric names it).  Prints ONE JSON line.

  value         queries/s with the index and the query batch resident in HBM (CUDA events, max over ranks)
  e2e           the same through the host-buffer C-ABI call (pinned host queries -> H2D -> score ->
                select -> D2H of ids+scores inside the timed region)
  roofline      score_tiles_kernel: algorithmic bytes per launch (12 B per posting touched + 4 B per
                (query, doc) score written) / its CUDA-event time, against the measured HBM peak
  cpu_baseline  the oracle's C port of the reference's per-query loop (doc-major scan + top-k) on the
                host cores, on a bounded sample of the same queries
  --impl reference : times that CPU path alone (the reference is pure Python/Numba and cannot travel
                to the GPU box; oracle/bm25_oracle.c restates it loop for loop)
"""
import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_docs, n_vocab, mean_len, n_queries, k, description)
    "c2": (1_000_000, 100_000, 60.0, 1024, 10,
           "synthetic Zipfian 1M docs x 100K vocab, avg 60 terms/doc, 1024-query batch of 4-8 terms, BM25 top-10"),
    "c3": (8_800_000, 100_000, 60.0, 1024, 100,
           "MS MARCO-passage-shape synthetic 8.8M docs x 100K vocab, 1024-query batch, BM25 top-100"),
    "small": (100_000, 20_000, 60.0, 256, 10, "small smoke workload (not a bench line)"),
    "c2s8": (125_000, 100_000, 60.0, 1024, 10,
             "one rank's share of c2 at 8 GPUs on a single GPU: fixed per-step overheads (not a bench line)"),
}
FALLBACK_HBM_GBS = 6650.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def make_workload(name, n_queries=None):
    from b200ret import synthetic as S
    import b200ret
    n_docs, n_vocab, mean_len, n_q, k, desc = WORKLOADS[name]
    n_q = n_queries or n_q
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, mean_len)
    idf = b200ret.reference_idf(indices, n_docs, n_vocab)
    avgdl = b200ret.reference_avgdl(dl)
    q_ptr, q_terms, q_w = S.zipf_queries(n_q, n_vocab)
    return dict(name=name, desc=desc, n_docs=n_docs, n_vocab=n_vocab, k=k, data=data, indices=indices,
                indptr=indptr, dl=dl, idf=idf, avgdl=avgdl, q_ptr=q_ptr, q_terms=q_terms, q_w=q_w)


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(w, n_sample, repeats=1):
    """queries/s of the oracle's restatement of the reference loop (retrieval.py:233-284) on the host."""
    from oracle import c_oracle
    n_sample = min(n_sample, len(w["q_ptr"]) - 1)
    qp = w["q_ptr"][:n_sample + 1]
    qt, qw = w["q_terms"][:qp[-1]], w["q_w"][:qp[-1]]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        c_oracle.bm25_search_batch(qp, qt, qw, w["n_vocab"], w["data"], w["indices"], w["indptr"], w["dl"], w["idf"],
                                   1.2, 0.75, w["avgdl"], w["k"])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_sample / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle
    w = make_workload(args.workload, args.n_queries)
    cores = c_oracle.use_all_host_threads()
    # size the per-step sample so that a step takes ~1.5 s
    rate, dt = cpu_reference_rate(w, 2)
    n_sample = int(max(2, min(len(w["q_ptr"]) - 1, round(1.5 * rate))))
    for _ in range(args.warmup):
        cpu_reference_rate(w, n_sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(w, n_sample)
    el = time.perf_counter() - t0
    qps = n_sample * args.steps / el
    sample = f"first {n_sample} of {len(w['q_ptr']) - 1} queries per step, full corpus"
    out = {
        "impl": "reference", "metric": f"bm25_top{w['k']}_queries_per_sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "n_docs": w["n_docs"], "n_vocab": w["n_vocab"],
                   "k": w["k"], "queries_per_step": n_sample},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python/Numba (cannot travel): timed here is oracle/bm25_oracle.c, its "
                "loop-for-loop C+OpenMP restatement (doc-major CSR scan per query + top-k), all host threads",
    }
    emit(out)
    return 0


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200ret
    from b200ret.dist import ShardedBM25, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("B2R_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    w = make_workload(args.workload, args.n_queries)
    n_docs, k = w["n_docs"], w["k"]
    b200ret.set_bank_schedule(bool(args.bank_schedule))
    lo, hi = shard_range(n_docs, world, rank)
    s, e = w["indptr"][lo], w["indptr"][hi]
    ix = b200ret.TermMajorIndex.from_csr(w["data"][s:e], w["indices"][s:e], w["indptr"][lo:hi + 1] - s, w["dl"][lo:hi],
                                         n_vocab=w["n_vocab"], idf=w["idf"], avgdl=w["avgdl"], doc_id_base=lo,
                                         tile_docs=args.tile_docs)
    sharded = ShardedBM25(ix)
    nq = len(w["q_ptr"]) - 1
    d_ptr = torch.from_numpy(w["q_ptr"]).to(dev)
    d_terms = torch.from_numpy(w["q_terms"]).to(dev)
    d_w = torch.from_numpy(w["q_w"]).to(dev)
    lib = b200ret._abi.lib

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """barrier+sync, CUDA events around `steps` calls on the current stream, max over ranks (ms)."""
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return float(ms.item())

    # ---- headline: inputs resident in HBM
    eager_step = lambda: sharded.search(d_ptr, d_terms, d_w, k)  # noqa: E731
    for _ in range(args.warmup):
        eager_step()
    torch.cuda.synchronize()
    step, graphed = eager_step, False
    if args.cuda_graph:
        # the step is a fixed sequence of ~12 launches (+ one NCCL all-gather): replay it as a CUDA graph so that
        # launch latency does not bound the multi-GPU runs (the buffers of the captured step are reused)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eager_step()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    g_out = sharded.search(d_ptr, d_terms, d_w, k)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

            def step():
                g.replay()
                return g_out
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            graphed = True
        except Exception as ex:      # capture is an optimisation, never a requirement
            print(f"[bench] CUDA graph capture unavailable ({type(ex).__name__}: {ex}); timing eager launches",
                  file=sys.stderr)
            step = eager_step
            torch.cuda.synchronize()
    n0 = lib.b2r_launch_count()
    eager_step()
    launches_per_step = int(lib.b2r_launch_count() - n0)
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        ms = timed(step, args.steps)
    launches = launches_per_step * args.steps
    qps = nq * args.steps / (ms * 1e-3)

    # ---- parity gate on what was just timed: a few queries against the oracle (rank 0, whole corpus)
    idx, val = step()
    torch.cuda.synchronize()
    idx, val = idx.clone(), val.clone()
    parity = None
    if rank == 0 and args.check > 0:
        from oracle import c_oracle
        c = min(args.check, nq)
        qp = w["q_ptr"][:c + 1]
        wi, wv = c_oracle.bm25_search_batch(qp, w["q_terms"][:qp[-1]], w["q_w"][:qp[-1]], w["n_vocab"], w["data"],
                                            w["indices"], w["indptr"], w["dl"], w["idf"], 1.2, 0.75, w["avgdl"], k)
        ok = bool(np.array_equal(idx[:c].cpu().numpy(), wi) and
                  np.array_equal(val[:c].cpu().numpy().view(np.uint32), np.where(wv == 0, np.float32(0), wv).view(np.uint32)))
        parity = {"queries_checked": c, "bit_exact_vs_oracle": ok}
        if not ok:
            print("PARITY FAILURE against the oracle", file=sys.stderr)

    # ---- e2e: host buffers in, host buffers out
    h2d = int(w["q_ptr"].nbytes + w["q_terms"].nbytes + w["q_w"].nbytes)
    d2h = int(nq * k * (8 + 4))
    if world == 1:
        e2e_step = lambda: ix.search_host(w["q_ptr"], w["q_terms"], w["q_w"], k)  # noqa: E731
    else:
        hp, ht, hw = (torch.from_numpy(w[n]).pin_memory() for n in ("q_ptr", "q_terms", "q_w"))
        oi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        ov = torch.empty((nq, k), dtype=torch.float32).pin_memory()

        def e2e_step():
            i_, v_ = sharded.search(hp.to(dev, non_blocking=True), ht.to(dev, non_blocking=True),
                                    hw.to(dev, non_blocking=True), k)
            oi.copy_(i_, non_blocking=True)
            ov.copy_(v_, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    for _ in range(max(1, args.warmup)):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps)
    e2e_qps = nq * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel: score_tiles_kernel with the fused-selection epilogue, bracketed
    # by CUDA events on the launching stream inside b2r_search_batch (b2r_set_profiling)
    import ctypes as C
    df_local = np.bincount(w["indices"][s:e], minlength=w["n_vocab"])
    n_samp, t_step, cap = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    lib.b2r_fused_plan(C.byref(ix._desc), k, C.byref(n_samp), C.byref(t_step), C.byref(cap))
    fused = n_samp.value > 0
    df_k = df_local          # the fused launch scores every tile of the shard (the sample launch comes on top)
    postings = int(df_local[w["q_terms"]].sum())
    postings_k = int(df_k[w["q_terms"]].sum())
    lib.b2r_set_profiling(1)
    k_times = []
    dense = None
    if fused:
        for _ in range(3 + args.steps):
            eager_step()
            t_ms = C.c_float(0)
            lib.b2r_profile_fused_ms(C.byref(t_ms), None)
            k_times.append(t_ms.value)
        k_ms = float(np.mean(k_times[3:]))
        alg_bytes = 12 * postings_k          # fused epilogue writes no score vector
        kernel_name = "score_tiles_kernel<BM25, FUSED>"
    else:
        dense = torch.empty((nq, ix.padded_docs), dtype=torch.float32, device=dev)
        score_only = lambda: ix.score_dense(d_ptr, d_terms, d_w, out=dense)  # noqa: E731
        for _ in range(3):
            score_only()
        k_ms = timed(score_only, args.steps) / args.steps
        alg_bytes = 12 * postings + 4 * nq * (hi - lo)
        kernel_name = "score_tiles_kernel<BM25, DENSE>"
    lib.b2r_set_profiling(0)
    del dense
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    step_bytes = 12 * postings + 8 * nq * (hi - lo)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{kernel_name}:{w['name']}:n{world}")
    except Exception:
        pass

    out = None
    if rank == 0:
        out = {
            "metric": f"bm25_top{k}_queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['desc']}", "n_docs": n_docs, "n_vocab": w["n_vocab"], "k": k,
                       "queries_per_step": nq, "sharding": f"doc-sharded x{world}", "tile_docs": args.tile_docs,
                       "l2": "inputs exceed L2: per step the index shard (%.2f GB) plus a %.2f GB score tile stream "
                             "through HBM; no flush needed" % (ix.device_bytes() / 1e9, nq * ix.padded_docs * 4 / 1e9),
                       "postings_touched_per_step_rank0": postings, "cuda_graph_replay": graphed,
                       "selection": ("fused: threshold = k-th largest group maximum of every %dth tile, all tiles scored with the "
                                     "candidate epilogue, cap %d" % (t_step.value, cap.value))
                       if fused else "plain: score vector + streaming select"},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                         "step_achieved_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                         "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "step_algorithmic_bytes": step_bytes,
                         "note": "achieved = 12 B x postings touched by this launch / its CUDA-event time; "
                                 "step_* = SURVEY 8d model 12*P + 8*N per query over the whole step"},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": clk.summary(), "parity": parity,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle
            cores = c_oracle.use_all_host_threads()
            r2, _ = cpu_reference_rate(w, 2)
            n_sample = int(max(2, min(nq, round(12.0 * r2))))
            rate, dt = cpu_reference_rate(w, n_sample)
            out["cpu_baseline"] = {"value": rate, "unit": "queries/s", "cores": cores, "kind": "port",
                                   "sample": f"first {n_sample} of {nq} queries, full corpus, {dt:.1f} s of wall time"}
        emit(out)
    if world > 1:
        # Every collective of the run has completed on every rank by now.  Tearing the NCCL communicator down
        # while a captured graph still references it was seen to hang, so leave without the teardown.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line, on the real stdout (fd 1 is pointed at stderr while the run is in progress so
    that library banners -- e.g. NCCL's version line -- cannot land beside it)."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tile-docs", type=int, default=4096)
    ap.add_argument("--check", type=int, default=4, help="queries checked against the oracle after timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the timed step as a CUDA graph (0 = eager)")
    ap.add_argument("--n-queries", type=int, default=None, help="override the batch size (profiling only)")
    ap.add_argument("--bank-schedule", type=int, default=1,
                    help="0 = build the index without the bank schedule of dense segments (A/B measurement only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "b200":
        sock = socket.socket()
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
        sock.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
