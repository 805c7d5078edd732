"""Doc-sharded search across the GPUs of one box (no reference counterpart: the reference is a
single CPU process).

One process per GPU (torch.distributed, NCCL).  Rank r owns the contiguous global document range
shard_range(n_docs, world, r); its TermMajorIndex carries doc_id_base = range start, so candidate
keys hold GLOBAL document indices and the (score desc, doc index asc) rule is global.  idf and avgdl
are global statistics: df is all-reduced (int64[V]) and avgdl is computed from the all-reduced
(sum of lengths, N) -- or passed in when the caller has the full doc_lengths vector, which
reproduces the reference's float(np.mean(f32)) bit-for-bit.  Per batch the only traffic is an
all-gather of u64[Q, k] candidate keys followed by a merge kernel (b2r_merge_candidates).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_range", "global_statistics", "gather_candidates", "merge_shard_candidates", "sharded_int8_scan",
           "ShardedBM25", "PeerExchange", "BatchPipeline"]


def shard_range(n_docs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split: every rank gets ceil(n/world) docs except the tail."""
    per = (n_docs + world - 1) // world
    lo = min(n_docs, rank * per)
    return lo, min(n_docs, lo + per)


def global_statistics(local_indices: np.ndarray, local_doc_lengths: np.ndarray, n_vocab: int, group=None,
                      device: Optional[torch.device] = None):
    """All-reduce df[V], sum(doc_len) and N over the shards.  Returns (idf f32[V], avgdl, n_docs_global)
    with the reference's expressions (rag_system/core/retrieval.py:187-190) applied to the global
    counts.  avgdl here is f32(sum)/N computed in f64 then rounded through f32, which equals
    float(np.mean(f32 array)) whenever the f32 pairwise sum is exact (integer lengths, N < 2^24 * ...)."""
    df = torch.from_numpy(np.bincount(local_indices, minlength=n_vocab).astype(np.int64))
    tot = torch.tensor([float(np.sum(local_doc_lengths.astype(np.float64))), float(len(local_doc_lengths))],
                       dtype=torch.float64)
    if device is not None:
        df, tot = df.to(device), tot.to(device)
    if dist.is_initialized():
        dist.all_reduce(df, group=group)
        dist.all_reduce(tot, group=group)
    df_h = df.cpu().numpy()
    n_global = int(round(float(tot[1])))
    idf = np.log((n_global - df_h + 0.5) / (df_h + 0.5)).astype(np.float32)
    avgdl = float(np.float32(float(tot[0]) / n_global))
    return idf, avgdl, n_global


def gather_candidates(local_keys: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of the [Q, k] candidate keys -> [world, Q, k]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_keys.unsqueeze(0)
    out = torch.empty((world,) + tuple(local_keys.shape), dtype=local_keys.dtype, device=local_keys.device)
    if local_keys.is_cuda:
        dist.all_gather_into_tensor(out, local_keys.contiguous(), group=group)
    else:   # gloo (CPU tests of the plumbing)
        dist.all_gather(list(out.unbind(0)), local_keys.contiguous(), group=group)
    return out


_PEER_CACHE: dict = {}


class PeerExchange:
    """Candidate exchange + merge in one kernel over peer memory (csrc/exchange.cu, b2r_exchange_merge).

    Every rank allocates a receive buffer as PyTorch symmetric memory (CUDA VMM memory mapped into every process of
    the group over NVLink) and hands the kernel the device array of all ranks' buffer addresses; from then on a step
    is ONE launch per rank -- no NCCL call, no host synchronisation, capturable in a CUDA graph.  `create()` returns
    None (and callers keep the NCCL all-gather path) when the allocation or the rendezvous fails, or
    B2R_EXCHANGE=nccl is set."""

    def __init__(self, n_queries: int, k: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _abi
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_queries, self.k = int(n_queries), int(k)
        nbytes = int(_abi.lib.b2r_exchange_bytes(self.world, self.n_queries, self.k))
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        self.handle.barrier()          # every rank's buffer is zeroed before anyone pushes into it

    @staticmethod
    def create(n_queries: int, k: int, device, group=None) -> Optional["PeerExchange"]:
        import os
        if os.environ.get("B2R_EXCHANGE", "peer") == "nccl" or not dist.is_initialized():
            return None
        try:
            return PeerExchange(n_queries, k, device, group)
        except Exception as ex:      # no symmetric memory on this box / build: the NCCL path stays
            import warnings
            warnings.warn(f"b200ret: peer-memory exchange unavailable ({type(ex).__name__}: {ex}); using NCCL all-gather")
            return None

    def fits(self, n_queries: int, k: int) -> bool:
        return int(n_queries) == self.n_queries and int(k) == self.k

    def merge(self, local_keys: torch.Tensor):
        """local_keys i64[Q, k] (u64 payload) -> (idx i64[Q, k] global doc indices, val f32[Q, k]) of the union."""
        from . import _abi
        from .index import _stream_ptr
        nq, k = int(local_keys.shape[0]), int(local_keys.shape[1])
        assert self.fits(nq, k), "PeerExchange was sized for another (n_queries, k)"
        dev = local_keys.device
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        val = torch.empty((nq, k), dtype=torch.float32, device=dev)
        _abi.check(_abi.lib.b2r_exchange_merge(local_keys.contiguous().data_ptr(), self.peer_ptrs.data_ptr(), self.rank,
                                               self.world, nq, k, None, idx.data_ptr(), val.data_ptr(), _stream_ptr(dev)),
                   "exchange + merge")
        return idx, val

    def check(self) -> None:
        """Synchronise and raise if a wait inside the kernel ever timed out."""
        from . import _abi
        from .index import _stream_ptr
        _abi.check(_abi.lib.b2r_exchange_status(self.buf.data_ptr(), _stream_ptr(self.buf.device)), "exchange status")


def merge_shard_candidates(local_keys: torch.Tensor, k: int, group=None):
    """all-gather the ranked [Q, k] candidate keys of every shard and merge them on the GPU
    (b2r_merge_candidates).  Returns (idx i64[Q,k] global doc indices, val f32[Q,k])."""
    from . import _abi
    from .index import _stream_ptr
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1 and local_keys.is_cuda:      # peer-memory exchange when available (cached per batch shape)
        ck = (int(local_keys.shape[0]), int(k), str(local_keys.device), id(group))
        if ck not in _PEER_CACHE:
            _PEER_CACHE[ck] = PeerExchange.create(ck[0], ck[1], local_keys.device, group)
        if _PEER_CACHE[ck] is not None:
            return _PEER_CACHE[ck].merge(local_keys)
    gathered = gather_candidates(local_keys, group)
    nq = int(local_keys.shape[0])
    dev = local_keys.device
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    val = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ws = torch.empty(nq * k * 8 + (1 << 20), dtype=torch.uint8, device=dev)
    _abi.check(_abi.lib.b2r_merge_candidates(gathered.data_ptr(), world, nq, k, None, idx.data_ptr(), val.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "merge candidates")
    return idx, val


def sharded_int8_scan(queries_int8, local_corpus_int8, query_scales, local_corpus_scales, k: int, doc_id_base: int,
                      group=None):
    """Doc-sharded INT8 scan: this rank scans its slice (global ids = doc_id_base + local row), then the
    k candidates per query cross NVLink and are merged."""
    from .kernels import int8_scan_topk
    _i, _v, keys = int8_scan_topk(queries_int8, local_corpus_int8, query_scales, local_corpus_scales, k,
                                  doc_id_base=doc_id_base)
    return merge_shard_candidates(keys, int(keys.shape[1]), group)


class ShardedBM25:
    """Holds this rank's shard and runs search -> all-gather -> merge."""

    def __init__(self, shard_index, group=None):
        self.ix = shard_index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._ws = None
        self._peer: Optional[PeerExchange] = None
        self._peer_tried = False
        self.exchange = "none" if self.world == 1 else "nccl all_gather + merge kernel"

    def search(self, q_ptr, q_terms, q_weights, k: int):
        from . import _abi
        from .index import _stream_ptr
        _idx, _val, keys = self.ix.search(q_ptr, q_terms, q_weights, k, return_keys=True)
        if self.world == 1:
            return _idx, _val
        # candidate exchange: one fused kernel over peer memory when the box offers it (sized on first use: every
        # rank runs the same batches), else NCCL all-gather + merge kernel
        if keys.is_cuda and (self._peer is None or not self._peer.fits(keys.shape[0], k)) and not (
                self._peer_tried and self._peer is None):
            self._peer = PeerExchange.create(int(keys.shape[0]), int(k), keys.device, self.group)
            self._peer_tried = True
            if self._peer is not None:
                self.exchange = "peer-memory exchange+merge kernel (symmetric memory over NVLink)"
        if self._peer is not None and self._peer.fits(keys.shape[0], k):
            return self._peer.merge(keys)
        gathered = gather_candidates(keys, self.group)
        nq = int(keys.shape[0])
        dev = keys.device
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        val = torch.empty((nq, k), dtype=torch.float32, device=dev)
        need = nq * k * 8 + (1 << 20)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _abi.check(_abi.lib.b2r_merge_candidates(gathered.data_ptr(), self.world, nq, k, None, idx.data_ptr(),
                                                 val.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                                 _stream_ptr(dev)), "merge candidates")
        return idx, val


class BatchPipeline:
    """`depth` independent query batches in flight on `depth` CUDA streams, captured as ONE CUDA graph.

    A search step is a dependent chain -- threshold sample, threshold, scoring, selection, candidate exchange -- whose
    small kernels leave the GPU mostly idle (0.09 of 0.45 ms per step on a 1/8 shard of the 1M-doc corpus).  With a
    second batch in flight they run under the other batch's scoring kernel, and that kernel's tail is filled too.
    Every lane is a ShardedBM25 of its own over the SAME shard index: the index keeps one workspace per stream, the
    lane owns its outputs and (multi-GPU) its exchange buffers, so the lanes never share mutable state.

        pipe = BatchPipeline(index, depth=2)
        outs = pipe.capture(q_ptr, q_terms, q_weights, k)    # [(idx, val)] per lane, device tensors (graph outputs)
        ...write new queries into the captured input tensors (same shapes)...
        pipe.replay()                                          # depth steps; results are in `outs`

    All ranks of a doc-sharded job must capture and replay in step (the exchange is collective)."""

    def __init__(self, shard_index, depth: int = 2, group=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.lanes = [ShardedBM25(shard_index, group) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(shard_index.device) for _ in range(depth)]
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.outputs = []

    @property
    def depth(self) -> int:
        return len(self.lanes)

    def capture(self, q_ptr, q_terms, q_weights, k: int):
        """Capture one step per lane on (q_ptr, q_terms, q_weights): CUDA tensors that stay the graph's inputs.  Each of
        the three may be ONE tensor (every lane reads the same batch) or a list with one tensor per lane."""
        def per_lane(x):
            return list(x) if isinstance(x, (list, tuple)) else [x] * self.depth
        q_ptr, q_terms, q_weights = per_lane(q_ptr), per_lane(q_terms), per_lane(q_weights)
        cur = torch.cuda.current_stream()
        for i, (lane, st) in enumerate(zip(self.lanes, self.streams)):   # eager first use: workspaces, exchange buffers
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(2):
                    lane.search(q_ptr[i], q_terms[i], q_weights[i], k)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        self.outputs = []
        with torch.cuda.stream(side):
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                for i, (lane, st) in enumerate(zip(self.lanes, self.streams)):
                    st.wait_stream(side)
                    with torch.cuda.stream(st):
                        self.outputs.append(lane.search(q_ptr[i], q_terms[i], q_weights[i], k))
                for st in self.streams:
                    side.wait_stream(st)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        return self.outputs

    def replay(self) -> None:
        """Enqueue `depth` steps on the current stream."""
        self.graph.replay()

    def check(self) -> None:
        for lane in self.lanes:
            if lane._peer is not None:
                lane._peer.check()

    def close(self) -> None:
        """Drop the graph (it references the process group's memory) and the lanes."""
        self.graph = None
        self.outputs = []
        for lane in self.lanes:
            lane.ix = None
        self.lanes = []
