#!/usr/bin/env python
"""Headline benchmark: BM25 top-10 queries/sec at 1M docs (BASELINE.json config 2) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path (score 1024 queries against the corpus, select top-10, and for
N > 1 exchange + merge the per-shard candidates).  The 1M-document corpus is doc-sharded over the
N ranks (strong scaling: the corpus is fixed, as the metric names it).  Prints ONE JSON line.

  value         queries/s with the index and the query batch resident in HBM (CUDA events, max over ranks); the K
                steps run as replays of a captured graph that holds three independent batches on three streams
                (run.batches_in_flight; run.ms_per_step_one_batch_in_flight is the serial figure)
  e2e           the same through the host-buffer C-ABI call (pinned host queries -> H2D -> score ->
                select -> D2H of ids+scores inside the timed region)
  roofline      the scorer (score_tiles_kernel, fused-selection epilogue): algorithmic bytes per launch
                (12 B per posting touched) / its CUDA-event time = an EFFECTIVE bandwidth, beside what
                ncu measured for it (DRAM bytes, L2->SM bytes, busiest pipes; profiles/traffic.json)
  parity        every query of the timed configuration against the oracle (N = 1: all of them, N > 1: 64)
  secondary     driver-visible timings + parity of the other BASELINE configs: C1 (FiQA-shape text corpus through
                RetrievalService.search_bm25, tokeniser included), C3 (8.8M docs, top-100, this job's N GPUs),
                C4 (SPLADE-shape impact index, 8.8M docs x 120 nnz, one GPU), C5 (INT8 10M x 768 scan, top-100,
                sharded over this job's N GPUs)
  cpu_baseline  the reference's per-query loop on the host cores, on a bounded sample of the same queries:
                kind "reference" = the reference's own Numba kernels (oracle/_ref, when the recipe
                oracle/make_ref.py could run), kind "port" = oracle/bm25_oracle.c (C + OpenMP restatement)
  --impl reference : times that CPU path alone, same config
"""
import argparse
import importlib.util
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
sys.path.insert(0, os.path.join(ROOT, "tools"))
PKG = os.path.join(ROOT, "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200")

WORKLOADS = {
    # name: (n_docs, n_vocab, mean_len, n_queries, k, description)
    "c2": (1_000_000, 100_000, 60.0, 1024, 10,
           "synthetic Zipfian 1M docs x 100K vocab, avg 60 terms/doc, 1024-query batch of 4-8 terms, BM25 top-10"),
    "small": (100_000, 20_000, 60.0, 256, 10, "small smoke workload (not a bench line)"),
    "c2s8": (125_000, 100_000, 60.0, 1024, 10,
             "one rank's share of c2 at 8 GPUs on a single GPU: fixed per-step overheads (not a bench line)"),
    "c2s4": (250_000, 100_000, 60.0, 1024, 10, "one rank's share of c2 at 4 GPUs on a single GPU (profiling, not a bench line)"),
    "c2s2": (500_000, 100_000, 60.0, 1024, 10, "one rank's share of c2 at 2 GPUs on a single GPU (profiling, not a bench line)"),
}
FALLBACK_HBM_GBS = 6650.0
# queries of the batch the CPU arm times per step (the first n of the 1024; they are i.i.d. draws)
CPU_SAMPLE = {"port": 128, "reference": 16}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def synthetic_module():
    """<package>/synthetic.py is numpy-only: load it by path, so that the CPU arm never loads libb200ret.so."""
    spec = importlib.util.spec_from_file_location("b2r_synthetic_standalone", os.path.join(PKG, "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_workload(name, n_queries=None):
    S = synthetic_module()
    n_docs, n_vocab, mean_len, n_q, k, desc = WORKLOADS[name]
    n_q = n_queries or n_q
    data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, mean_len)
    # the reference's host expressions (rag_system/core/retrieval.py:187-190)
    df = np.bincount(indices, minlength=n_vocab)
    idf = np.log((n_docs - df + 0.5) / (df + 0.5)).astype(np.float32)
    avgdl = float(np.mean(np.asarray(dl, dtype=np.float32)))
    q_ptr, q_terms, q_w = S.zipf_queries(n_q, n_vocab)
    return dict(name=name, desc=desc, n_docs=n_docs, n_vocab=n_vocab, k=k, data=data, indices=indices,
                indptr=indptr, dl=dl, idf=idf, avgdl=avgdl, q_ptr=q_ptr, q_terms=q_terms, q_w=q_w, df=df)


def workload_config(w):
    """The `config` object: identical in both arms (it names the workload, not the run)."""
    return {"workload": f"{w['name']}: {w['desc']}", "n_docs": w["n_docs"], "n_vocab": w["n_vocab"], "k": w["k"],
            "queries_per_step": len(w["q_ptr"]) - 1,
            "l2": "no flush: the index (0.7 GB of postings + tables, 0.2 GB of packed postings for the search path) is "
                  "larger than the 126 MB L2, so every step re-reads the touched posting lists from HBM once; re-use "
                  "ACROSS the 1024 queries of one step is served by L2 by design (roofline.traffic is the measured DRAM "
                  "volume per launch)"}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_kind():
    from oracle import ref_runner
    if os.environ.get("B2R_CPU_KIND") in ("port", "reference"):
        return os.environ["B2R_CPU_KIND"]
    try:
        if ref_runner.available():
            ref_runner.module()
            return "reference"
    except Exception as ex:
        print(f"[bench] oracle/_ref not usable ({type(ex).__name__}: {ex}); CPU arm = port", file=sys.stderr)
    return "port"


def cpu_search(w, n_sample, kind):
    """(idx, val, seconds) of the first n_sample queries through the CPU arm."""
    from oracle import c_oracle, ref_runner
    n_sample = min(n_sample, len(w["q_ptr"]) - 1)
    qp = w["q_ptr"][:n_sample + 1]
    qt, qw = w["q_terms"][:qp[-1]], w["q_w"][:qp[-1]]
    fn = ref_runner.search_batch if kind == "reference" else c_oracle.bm25_search_batch
    t0 = time.perf_counter()
    idx, val = fn(qp, qt, qw, w["n_vocab"], w["data"], w["indices"], w["indptr"], w["dl"], w["idf"], 1.2, 0.75,
                  w["avgdl"], w["k"])
    return idx, val, time.perf_counter() - t0


def cpu_threads(kind):
    from oracle import c_oracle, ref_runner
    return ref_runner.use_all_host_threads() if kind == "reference" else c_oracle.use_all_host_threads()


def cpu_sample_text(kind, n_sample, nq):
    what = ("the reference's own simd_bm25_score + fast_topk_selection (Numba, oracle/_ref) called per query as "
            "_score_bm25_query does" if kind == "reference" else
            "oracle/bm25_oracle.c, the C+OpenMP restatement of that per-query loop")
    return f"first {n_sample} of the {nq} queries per step, full 1M-doc corpus; {what}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = make_workload(args.workload, args.n_queries)
    nq = len(w["q_ptr"]) - 1
    kind = cpu_kind()
    cores = cpu_threads(kind)
    n_sample = min(CPU_SAMPLE[kind], nq)
    cpu_search(w, 2, kind)                        # JIT / page-in, untimed
    for _ in range(args.warmup):
        cpu_search(w, n_sample, kind)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_search(w, n_sample, kind)
    el = time.perf_counter() - t0
    qps = n_sample * args.steps / el
    out = {
        "impl": "reference", "metric": f"bm25_top{w['k']}_queries_per_sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, n_sample, nq)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if kind == "reference":      # for continuity with round 1, whose CPU arm was the (faster) C port
        cores_p = cpu_threads("port")
        n_p = min(CPU_SAMPLE["port"], nq)
        cpu_search(w, 2, "port")
        _, _, dt = cpu_search(w, n_p, "port")
        out["port_baseline"] = {"value": n_p / dt, "unit": "queries/s", "cores": cores_p, "kind": "port",
                                "sample": cpu_sample_text("port", n_p, nq)}
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


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _fold(v):
    return np.where(v == 0, np.float32(0), v)


# --------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200ret
    from b200ret.dist import BatchPipeline, ShardedBM25, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)      # NCCL_DEBUG is left as the caller set it

    w = make_workload(args.workload, args.n_queries)
    n_docs, k = w["n_docs"], w["k"]
    b200ret.set_bank_schedule(int(args.bank_schedule))
    b200ret.set_approx_prefilter(bool(args.prefilter))
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
    G = {"graph": None, "out": None}          # holder: the graph must be droppable before NCCL teardown
    if args.cuda_graph:
        # the step is a fixed sequence of launches (+ the candidate exchange): replay it as a CUDA graph so that
        # launch latency does not bound the multi-GPU runs (the buffers of the captured step are reused)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eager_step()
                G["graph"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(G["graph"], stream=side):
                    G["out"] = sharded.search(d_ptr, d_terms, d_w, k)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

            def step():
                G["graph"].replay()
                return G["out"]
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            graphed = True
        except Exception as ex:      # capture is an optimisation, never a requirement
            print(f"[bench] CUDA graph capture unavailable ({type(ex).__name__}: {ex}); timing eager launches",
                  file=sys.stderr)
            step = eager_step
            G["graph"] = G["out"] = None
            torch.cuda.synchronize()
    n0 = lib.b2r_launch_count()
    eager_step()
    launches_per_step = int(lib.b2r_launch_count() - n0)
    torch.cuda.synchronize()

    # ---- batch pipelining: `depth` independent batches in flight on `depth` streams (one captured graph = depth
    # steps).  A step is a dependent chain -- sample, threshold, score, select, exchange -- whose small kernels leave
    # the GPU mostly idle; with two batches in flight they run under the other batch's scoring kernel.  Each lane
    # has its own workspace (TermMajorIndex keeps one per stream), outputs and exchange buffers.
    depth = max(1, args.pipeline) if graphed else 1
    pipe, pipe_rem, lane_out = None, None, []
    if depth > 1:
        try:
            pipe = BatchPipeline(ix, depth)
            lane_out = pipe.capture(d_ptr, d_terms, d_w, k)
            for _ in range(2):
                pipe.replay()
            if args.steps % depth > 1:          # the last steps % depth steps of the timed region: a smaller pipeline
                pipe_rem = BatchPipeline(ix, args.steps % depth)
                lane_out = lane_out + pipe_rem.capture(d_ptr, d_terms, d_w, k)
                pipe_rem.replay()
            torch.cuda.synchronize()
        except Exception as ex:
            print(f"[bench] batch pipelining unavailable ({type(ex).__name__}: {ex}); one batch in flight",
                  file=sys.stderr)
            depth, pipe, pipe_rem, lane_out = 1, None, None, []
            torch.cuda.synchronize()

    def run_steps(n):
        if depth > 1:
            for _ in range(n // depth):
                pipe.replay()
            if n % depth > 1 and pipe_rem is not None and pipe_rem.depth == n % depth:
                pipe_rem.replay()
            else:
                for _ in range(n % depth):
                    step()
        else:
            for _ in range(n):
                step()
    with ClockSampler(local) as clk:
        ms = timed(lambda: run_steps(args.steps), 1)
    ms_single = timed(step, args.steps) if depth > 1 else ms      # one batch in flight (for reference)
    launches = launches_per_step * args.steps
    qps = nq * args.steps / (ms * 1e-3)
    if depth > 1:       # every lane must have produced what the single step produces
        ref_i, ref_v = step()
        torch.cuda.synchronize()
        for li, lv in lane_out:
            if not (torch.equal(li, ref_i) and torch.equal(lv, ref_v)):
                print("PARITY FAILURE: a pipelined lane differs from the single step", file=sys.stderr)
                depth = -depth

    idx, val = step()
    torch.cuda.synchronize()
    got_idx, got_val = idx.cpu().numpy().copy(), val.cpu().numpy().copy()

    # ---- e2e: host buffers in, host buffers out
    h2d = int(w["q_ptr"].nbytes + w["q_terms"].nbytes + w["q_w"].nbytes)
    d2h = int(nq * k * (8 + 4))
    if world == 1:
        # the reference-facing call: b2r_search_batch_host, HOST buffers in and out, one call at a time (two host
        # threads with one call in flight each were measured and are slower: 2.90 vs 2.82 ms per step on one box --
        # the hand-off between Python threads costs more than the overlap gains)
        e2e_step = lambda: ix.search_host(w["q_ptr"], w["q_terms"], w["q_w"], k)  # noqa: E731
        hi_, hv_ = e2e_step()
        if not (np.array_equal(hi_, got_idx) and np.array_equal(_bits(hv_), _bits(got_val))):
            print("PARITY FAILURE: host-buffer call differs from the device-buffer step", file=sys.stderr)
        e2e_api = "b2r_search_batch_host (C ABI, host buffers in / host buffers out)"
    else:
        hp, ht, hw = (torch.from_numpy(w[n]).pin_memory() for n in ("q_ptr", "q_terms", "q_w"))
        oi = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        ov = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        if graphed:
            # the sharded step as captured, fed from pinned host memory: H2D into the graph's input tensors, replay
            # (search + candidate exchange + merge), D2H of ids and scores.  With batches in flight every lane has its
            # own input and output buffers and one host synchronisation covers the `depth` steps of a replay.
            e2e_depth = depth if depth > 1 else 1
            e_in = [[t.clone() for t in (d_ptr, d_terms, d_w)] for _ in range(e2e_depth)]
            e_oi = [torch.empty((nq, k), dtype=torch.int64).pin_memory() for _ in range(e2e_depth)]
            e_ov = [torch.empty((nq, k), dtype=torch.float32).pin_memory() for _ in range(e2e_depth)]
            if e2e_depth > 1:
                e_pipe = BatchPipeline(ix, e2e_depth)
                e_out = e_pipe.capture([x[0] for x in e_in], [x[1] for x in e_in], [x[2] for x in e_in], k)
                G["e2e_pipe"] = e_pipe

            def e2e_pair():          # = e2e_depth steps
                for i in range(e2e_depth):
                    e_in[i][0].copy_(hp, non_blocking=True)
                    e_in[i][1].copy_(ht, non_blocking=True)
                    e_in[i][2].copy_(hw, non_blocking=True)
                if e2e_depth > 1:
                    e_pipe.replay()
                    outs = e_out
                else:
                    d_ptr.copy_(e_in[0][0], non_blocking=True)
                    d_terms.copy_(e_in[0][1], non_blocking=True)
                    d_w.copy_(e_in[0][2], non_blocking=True)
                    G["graph"].replay()
                    outs = [G["out"]]
                for i in range(e2e_depth):
                    e_oi[i].copy_(outs[i][0], non_blocking=True)
                    e_ov[i].copy_(outs[i][1], non_blocking=True)
                torch.cuda.current_stream().synchronize()

            def e2e_step():
                e2e_pair()
            e2e_step.steps_per_call = e2e_depth
            e2e_api = ("ShardedBM25.search as a CUDA-graph replay (%d batch(es) in flight), pinned host queries in, pinned "
                       "host ids/scores out" % e2e_depth)
        else:
            def e2e_step():
                i_, v_ = sharded.search(hp.to(dev, non_blocking=True), ht.to(dev, non_blocking=True),
                                        hw.to(dev, non_blocking=True), k)
                oi.copy_(i_, non_blocking=True)
                ov.copy_(v_, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            e2e_api = "ShardedBM25.search (eager launches), pinned host queries in, pinned host ids/scores out"
    for _ in range(max(1, args.warmup)):
        e2e_step()
    per_call = getattr(e2e_step, "steps_per_call", 1)
    e2e_ms = timed(e2e_step, args.steps // per_call) * args.steps / max(1, (args.steps // per_call) * per_call)
    e2e_qps = nq * args.steps / (e2e_ms * 1e-3)

    # ---- the dominant kernel: score_tiles_kernel with the fused-selection epilogue, bracketed by CUDA events on
    # the launching stream inside b2r_search_batch (b2r_set_profiling)
    import ctypes as C
    df_local = np.bincount(w["indices"][s:e], minlength=w["n_vocab"])
    n_samp, t_step, cap = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    lib.b2r_fused_plan(C.byref(ix._desc), k, C.byref(n_samp), C.byref(t_step), C.byref(cap))
    fused = n_samp.value > 0
    postings = int(df_local[w["q_terms"]].sum())
    lib.b2r_set_profiling(1)
    k_times = []
    dense = None
    prefiltered = False
    if fused:
        for _ in range(3 + args.steps):
            eager_step()
            t_ms = C.c_float(0)
            lib.b2r_profile_fused_ms(C.byref(t_ms), None)
            k_times.append(t_ms.value)
        k_ms = float(np.mean(k_times[3:]))
        alg_bytes = 12 * postings          # fused epilogue writes no score vector
        prefiltered = bool(args.prefilter) and "post_pk" in ix._bufs
        kernel_name = "score_approx_kernel<FUSED>" if prefiltered else "score_tiles_kernel<BM25, FUSED>"
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
    layout_bytes = (4 if prefiltered else 12) * postings
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    step_bytes = 12 * postings + 8 * nq * (hi - lo)
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ncu = json.load(f).get(f"{kernel_name}:{w['name']}:n{world}") or {}
            if not isinstance(ncu, dict):
                ncu = {"dram_bytes": ncu}
    except Exception:
        pass
    traffic = ncu.get("dram_bytes")

    # ---- other BASELINE configs, driver-visible (all ranks take part in C3 / C5; C1 is a one-GPU path)
    # the captured graph references the communicator and the index buffers: drop it before anything else
    torch.cuda.synchronize()
    step = eager_step = run_steps = e2e_step = None
    if G.get("e2e_pipe") is not None:
        if world > 1:
            G["e2e_pipe"].check()
        G["e2e_pipe"].close()
    G.clear()
    for p_ in (pipe, pipe_rem):
        if p_ is not None:
            if world > 1:
                p_.check()
            p_.close()
    pipe, pipe_rem, lane_out = None, None, []
    idx = val = None
    fused_info = (t_step.value, cap.value)
    exchange_kind = sharded.exchange
    if world > 1 and sharded._peer is not None:
        sharded._peer.check()          # raises if a wait inside the exchange kernel ever timed out
    sharded.ix = None
    sharded = ix = None
    torch.cuda.empty_cache()
    secondary = None
    if args.secondary:
        secondary = run_secondary(args, world, rank, dev, timed)

    out = None
    if rank == 0:
        from oracle import c_oracle
        # ---- parity gate on what was just timed, against the oracle over the WHOLE corpus
        cpu_base = None
        n_par = min(nq, args.check) if args.check is not None else (nq if world == 1 else min(nq, 64))
        c_oracle.use_all_host_threads()
        t0 = time.perf_counter()
        wi, wv, dt_port = cpu_search(w, n_par, "port")
        ok = bool(np.array_equal(got_idx[:n_par], wi) and np.array_equal(_bits(got_val[:n_par]), _bits(_fold(wv))))
        parity = {"queries_checked": n_par, "bit_exact_vs_oracle": ok, "ids_and_scores": True}
        if not ok:
            print("PARITY FAILURE against the oracle", file=sys.stderr)
        if world == 1 and not args.no_cpu_baseline:
            kind = cpu_kind()
            cores = cpu_threads(kind)
            n_sample = min(CPU_SAMPLE[kind], nq)
            cpu_search(w, 2, kind)
            ri, rv, dt = cpu_search(w, n_sample, kind)
            cpu_base = {"value": n_sample / dt, "unit": "queries/s", "cores": cores, "kind": kind,
                        "sample": cpu_sample_text(kind, n_sample, nq) + f", {dt:.1f} s of wall time"}
            if kind == "reference":
                # the reference's own scores, bit for bit (its order among tied scores is unspecified)
                parity["reference_kernel_scores_bit_exact"] = bool(
                    np.array_equal(_bits(rv), _bits(wv[:n_sample])))
                parity["reference_kernel_queries"] = n_sample
                cpu_base["port_value"] = n_par / dt_port
        cfg = workload_config(w)
        out = {
            "metric": f"bm25_top{k}_queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "run": {"sharding": f"doc-sharded x{world}", "exchange": exchange_kind, "tile_docs": args.tile_docs,
                    "postings_touched_per_step_rank0": postings, "cuda_graph_replay": graphed,
                    "batches_in_flight": depth, "ms_per_step_one_batch_in_flight": ms_single / args.steps,
                    "launches_per_step": launches_per_step,
                    "prefilter": ("f32 pre-filter on 4-byte packed postings (score_approx_kernel), survivors rescored with "
                                  "the exact f64 chain before ranking" if prefiltered else "off: every posting scored in f64"),
                    "selection": ("fused: threshold = k-th largest group maximum of every %dth tile, all tiles scored with "
                                  "the candidate epilogue, cap %d" % fused_info)
                    if fused else "plain: score vector + streaming select"},
            "roofline": {
                "bound": ncu.get("bound", "lsu/shared+l2 (see bound_evidence)"),
                "bound_evidence": ncu.get("evidence"),
                "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_meaning": "EFFECTIVE bandwidth: SURVEY 8d algorithmic bytes (12 B x postings touched) / kernel time / "
                                "measured HBM peak; posting lists are shared by the queries of a step through L2, so this "
                                "is not DRAM utilisation (dram_frac is)",
                "traffic": traffic,
                "dram_frac": (traffic / (k_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "l2_to_sm_bytes": ncu.get("l2_to_sm_bytes"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "kernel_layout_bytes_per_launch": layout_bytes,
                "kernel_layout_frac": layout_bytes / (k_ms * 1e-3) / 1e9 / peak,
                "kernel_layout_note": ("bytes this kernel's own data layout makes it read per launch: 4 B x postings touched "
                                       "(packed f32 pre-filter postings) -- the SURVEY figure above counts the reference "
                                       "layout's 12 B per posting, which the pre-filter no longer moves"
                                       if prefiltered else "12 B x postings touched (u32 doc + f64 value)"),
                "step_model_bytes": step_bytes,
                "step_model_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                "note": "step_model_* = SURVEY 8d step model 12*P + 8*N*Q; the fused path never writes or reads the "
                        "8*N*Q score bytes, so values above 1 mean avoided traffic, not skipped work (parity below "
                        "covers every query of the timed configuration)"},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "api": e2e_api},
            "gpu_launches": launches, "clocks": clk.summary(), "parity": parity,
        }
        if cpu_base is not None:
            out["cpu_baseline"] = cpu_base
        if secondary is not None:
            out["secondary"] = secondary
        emit(out)
    if world > 1:
        # the captured graph that referenced the communicator was dropped above: tear NCCL down properly
        sync_all()
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------- secondary configs
def run_secondary(args, world, rank, dev, timed):
    """C1 / C3 / C5 of BASELINE.json, each timed with CUDA events (C1: wall clock, its tokeniser is host code) and
    checked against the oracle on a sample.  Every rank runs C3 and C5 (sharded); rank 0 alone runs C1."""
    import torch
    import torch.distributed as dist
    import b200ret
    from b200ret.dist import ShardedBM25, shard_range, sharded_int8_scan
    from gpu_synth import global_bm25_stats, random_int8_corpus, random_int8_queries, zipf_csr_torch
    from b200ret import synthetic as S
    sec = {}
    steps = 5

    # ---- C3: 8.8M docs x 100K vocab, 1024 queries, BM25 top-100, doc-sharded over this job's ranks
    try:
        n_docs, n_vocab, k, nq = 8_800_000, 100_000, 100, 1024
        data, ind, ptr, dl = zipf_csr_torch(n_docs, n_vocab, 60.0, 20260101, dev)
        df, idf, avgdl = global_bm25_stats(ind, dl, n_docs, n_vocab)
        q_ptr, q_terms, q_w = S.zipf_queries(nq, n_vocab)
        lo, hi = shard_range(n_docs, world, rank)
        s_, e_ = int(ptr[lo]), int(ptr[hi])
        ix = b200ret.TermMajorIndex.from_csr(data[s_:e_], ind[s_:e_], ptr[lo:hi + 1] - s_, dl[lo:hi], n_vocab=n_vocab,
                                             idf=idf, avgdl=avgdl, doc_id_base=lo)
        host = [t.cpu().numpy() for t in (data, ind, ptr, dl)] if rank == 0 else None
        del data, ind, ptr, dl
        torch.cuda.empty_cache()
        sh = ShardedBM25(ix)
        d_q = [torch.from_numpy(a).to(dev) for a in (q_ptr, q_terms, q_w)]
        for _ in range(3):
            idx, val = sh.search(*d_q, k)
        ms_eager = timed(lambda: sh.search(*d_q, k), steps) / steps
        ws_bytes = int(ix._ws.numel()) if ix._ws is not None else 0
        # the same step with two batches in flight (dist.BatchPipeline: one captured graph = two steps on two streams)
        ms, in_flight = ms_eager, 1
        try:
            from b200ret.dist import BatchPipeline
            pipe = BatchPipeline(ix, 2)
            lanes = pipe.capture(*d_q, k)
            pipe.replay()
            torch.cuda.synchronize()
            if all(torch.equal(a, idx) and torch.equal(b_, val) for a, b_ in lanes):
                ms_pipe = timed(pipe.replay, steps) / (2 * steps)
                if ms_pipe < ms_eager:          # (on one GPU the chain is a small part of a 24 ms step: keep the better)
                    ms, in_flight = ms_pipe, 2
            if world > 1:
                pipe.check()
            pipe.close()
        except Exception as ex:
            print(f"[bench] C3: batch pipelining unavailable ({type(ex).__name__}: {ex})", file=sys.stderr)
        rec = {"workload": f"c3: synthetic Zipfian 8.8M docs x 100K vocab, 1024-query batch, BM25 top-100, doc-sharded "
                           f"x{world}", "ms_per_step": ms, "queries_per_s": nq / (ms * 1e-3), "n_gpus": world,
               "batches_in_flight": in_flight, "ms_per_step_one_batch_eager": ms_eager,
               "postings_touched_per_step": int(df[q_terms].sum()), "index_bytes_rank0": ix.device_bytes(),
               "workspace_bytes_rank0": ws_bytes}
        if rank == 0:
            from oracle import c_oracle
            c_oracle.use_all_host_threads()
            nc = 16
            e = int(q_ptr[nc])
            wi, wv = c_oracle.bm25_search_batch(q_ptr[:nc + 1], q_terms[:e], q_w[:e], n_vocab, host[0], host[1], host[2],
                                                host[3], idf, 1.2, 0.75, avgdl, k)
            rec["parity"] = {"queries_checked": nc, "ids_and_scores": True, "bit_exact_vs_oracle": bool(
                np.array_equal(idx[:nc].cpu().numpy(), wi) and
                np.array_equal(_bits(val[:nc].cpu().numpy()), _bits(_fold(wv))))}
        sec["c3_8p8m_top100"] = rec
        del sh, ix, host, d_q
        torch.cuda.empty_cache()
    except Exception as ex:
        sec["c3_8p8m_top100"] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- C5: INT8 768-d exhaustive scan, 10M vectors sharded over this job's ranks, 1024 queries, top-100
    try:
        n, dim, k, nq = 10_000_000, 768, 100, 1024
        lo, hi = shard_range(n, world, rank)
        d8, ds = random_int8_corpus(n, dim, 42, dev)          # every rank draws the same corpus, keeps its slice
        if world > 1:
            d8, ds = d8[lo:hi].clone(), ds[lo:hi].clone()
            torch.cuda.empty_cache()
        q8, qs = random_int8_queries(nq, dim, 43, dev)
        run = (lambda: sharded_int8_scan(q8, d8, qs, ds, k, lo)) if world > 1 else (
            lambda: b200ret.int8_scan_topk(q8, d8, qs, ds, k)[:2])
        for _ in range(2):
            idx, val = run()
        ms = timed(run, steps) / steps
        rec = {"workload": f"c5: INT8 768-d exhaustive scan, 10M vectors sharded x{world}, 1024 queries, top-100 "
                           "(tcgen05 cta_group::2 pair kernel)", "ms_per_step": ms, "queries_per_s": nq / (ms * 1e-3),
               "int8_pops": 2.0 * nq * n * dim / (ms * 1e-3) / 1e15, "n_gpus": world}
        # parity: exact integer dots + f64 scale chain for 2 queries over ALL 10M vectors (each rank its slice,
        # then the union's top-k on rank 0)
        nc = 2
        loc_keys = []
        for q in range(nc):
            sc = torch.empty(hi - lo, dtype=torch.float32, device=dev)
            for c0 in range(0, hi - lo, 1 << 20):
                c1 = min(hi - lo, c0 + (1 << 20))
                dt_ = (d8[c0:c1].to(torch.int32) * q8[q].to(torch.int32)).sum(1).to(torch.float64)
                sc[c0:c1] = ((dt_ * qs[q].double()) * ds[c0:c1].double()).float()
            order = torch.argsort(sc, descending=True, stable=True)[:k]
            loc_keys.append(torch.stack([sc[order].double(), (order + lo).double()]))
        loc = torch.stack(loc_keys)                                   # [nc, 2, k]
        if world > 1:
            allk = [torch.empty_like(loc) for _ in range(world)]
            dist.all_gather(allk, loc)
            loc = torch.cat(allk, dim=2)
        if rank == 0:
            ok = True
            for q in range(nc):
                v, i = loc[q, 0].cpu().numpy().astype(np.float32), loc[q, 1].cpu().numpy().astype(np.int64)
                o = np.lexsort((i, -v))[:k]
                ok &= bool(np.array_equal(idx[q].cpu().numpy(), i[o]) and
                           np.array_equal(_bits(val[q].cpu().numpy()), _bits(v[o])))
            rec["parity"] = {"queries_checked": nc, "ids_and_scores": True, "bit_exact_vs_exact_integer_evaluation": ok}
        sec["c5_int8_10m_top100"] = rec
        del d8, ds, q8, qs
        torch.cuda.empty_cache()
    except Exception as ex:
        sec["c5_int8_10m_top100"] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- C4: SPLADE-shape impact index at its named size, 8.8M docs x 30,522 vocab x 120 nnz/doc (1.056e9 postings),
    #      256 weighted 30-term queries, impact dot top-10 (rank 0; BASELINE names no sharding for this config)
    if rank == 0:
        try:
            n_docs, n_vocab, k, nq = 8_800_000, 30522, 10, 256
            data, ind, ptr, _ = zipf_csr_torch(n_docs, n_vocab, 0, 20260103, dev, distinct_per_doc=120, chunk=1 << 18)
            idf1 = np.ones(n_vocab, np.float32)
            q_ptr, q_terms, q_w = S.impact_queries(nq, n_vocab, 30)
            ix = b200ret.TermMajorIndex.from_csr(data, ind, ptr, None, n_vocab=n_vocab, idf=idf1, kind="impact")
            df = torch.bincount(ind, minlength=n_vocab).cpu().numpy()
            nc = 4
            host = [t.cpu().numpy() for t in (data, ind, ptr)]
            del data, ind, ptr
            torch.cuda.empty_cache()
            d_q = [torch.from_numpy(a).to(dev) for a in (q_ptr, q_terms, q_w)]
            for _ in range(2):
                idx, val = ix.search(*d_q, k)
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(steps):
                ix.search(*d_q, k)
            b_.record()
            torch.cuda.synchronize()
            ms = a_.elapsed_time(b_) / steps
            from oracle import c_oracle
            c_oracle.use_all_host_threads()
            ok = True
            for q in range(nc):
                qtf = np.zeros(n_vocab, np.float32)
                qtf[q_terms[q_ptr[q]:q_ptr[q + 1]]] = q_w[q_ptr[q]:q_ptr[q + 1]]
                wi, wv = c_oracle.topk(c_oracle.tfidf_scores(qtf, host[0], host[1], host[2], idf1), k)
                ok &= bool(np.array_equal(idx[q].cpu().numpy(), wi) and
                           np.array_equal(_bits(val[q].cpu().numpy()), _bits(_fold(wv))))
            sec["c4_splade_8p8m_top10"] = {
                "workload": "c4: SPLADE-shape 8.8M docs x 30,522 vocab, 120 nnz/doc (1.056e9 postings), 256 queries of 30 "
                            "weighted terms, impact dot top-10, one GPU", "ms_per_step": ms,
                "queries_per_s": nq / (ms * 1e-3), "postings_touched_per_step": int(df[q_terms].sum()),
                "index_bytes": ix.device_bytes(),
                "parity": {"queries_checked": nc, "ids_and_scores": True, "bit_exact_vs_oracle": ok}}
            del ix, host, d_q
            torch.cuda.empty_cache()
        except Exception as ex:
            sec["c4_splade_8p8m_top10"] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- C1: FiQA-shape text corpus through RetrievalService (rank 0; tokeniser and dict building included)
    if rank == 0:
        try:
            import tempfile
            corpus = S.fiqa_shape_corpus()
            queries = S.fiqa_shape_queries()
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "docs.idx")
                b200ret.MemoryIndex(path, create=True).close()
                with b200ret.RetrievalService(path) as svc:
                    t0 = time.perf_counter()
                    svc.build_bm25_index(corpus)
                    build_s = time.perf_counter() - t0
                    got = svc.search_bm25(queries, top_k=10)
                    best = None
                    for _ in range(5):
                        svc.clear_cache()                      # search_bm25 serves repeated texts from its cache
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        svc.search_bm25(queries, top_k=10)
                        dt = time.perf_counter() - t0
                        best = dt if best is None else min(best, dt)
            rec = {"workload": "c1: FiQA-shape synthetic text corpus of the reference's generator (57,638 docs, 648 "
                               "queries, top-10) through RetrievalService.build_bm25_index / search_bm25",
                   "search_bm25_queries_per_s": len(queries) / best, "search_bm25_ms_per_call": best * 1e3,
                   "build_bm25_index_s": build_s, "timing": "wall clock, best of 5 (host tokeniser + packing + one GPU "
                                                            "call + result dicts); query cache cleared before each call",
                   "reference_queries_per_s": {"published_fiqa": 314.67, "survey_probe_8_threads": 133.0}}
            try:
                z = np.load(os.path.join(ROOT, "tests", "golden", "fiqa_shape.npz"))
                ok = True
                for qi, qid in enumerate(queries):
                    keep = z["canon_val"][qi] > 0
                    ok &= list(got[qid]) == [f"doc_{i}" for i in z["canon_idx"][qi][keep]]
                    ok &= list(got[qid].values()) == [float(v) for v in z["canon_val"][qi][keep]]
                rec["parity"] = {"queries_checked": len(queries), "ids_and_scores": True,
                                 "bit_exact_vs_reference_fixture": bool(ok)}
            except FileNotFoundError:
                rec["parity"] = None
            sec["c1_fiqa_shape_search_bm25"] = rec
        except Exception as ex:
            sec["c1_fiqa_shape_search_bm25"] = {"error": f"{type(ex).__name__}: {ex}"}
    return sec


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
    ap.add_argument("--check", type=int, default=None,
                    help="queries checked against the oracle (default: all of them at N = 1, 64 at N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", type=int, default=1, help="0 = skip the C1 / C3 / C5 secondary measurements")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the timed step as a CUDA graph (0 = eager)")
    ap.add_argument("--pipeline", type=int, default=3,
                    help="independent batches in flight on separate streams inside the timed region (1 = none)")
    ap.add_argument("--n-queries", type=int, default=None, help="override the batch size (profiling only)")
    ap.add_argument("--prefilter", type=int, default=1,
                    help="0 = score every posting in f64 (round-1 kernel) instead of the f32 pre-filter + exact rescoring")
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
