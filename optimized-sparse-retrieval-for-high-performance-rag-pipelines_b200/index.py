"""Term-major posting index resident in HBM, and the batched search entry points over it.

This is the host side of the hot path that RetrievalService.build_bm25_index / search_bm25
(reference: rag_system/core/retrieval.py:129-296) delegate to: it owns the device buffers as
PyTorch tensors (buffer ownership only -- no torch op touches the data path) and calls
libb200ret.so through the C ABI.  The doc-major scipy CSR the reference builds
(retrieval.py:176-184) is the INPUT; the layout in HBM is term-major and tile-partitioned:

  post_doc  u32[nnz]   local doc index of every posting, grouped by (term, doc tile)
  post_val  f64[nnz]   BM25: (tf*(k1+1))/(tf+k1*(1-b+b*dl/avgdl))   |   f32[nnz] impact weight
  blk_ptr   u32[V*T+1] postings of (term t, tile T) are [blk_ptr[t*T_n+T], blk_ptr[t*T_n+T+1]), doc-ascending
  dense_id  i32[V]     row of dense_ptr for terms averaging >= 64 postings per tile, else -1
  dense_ptr u32[...]   per dense term: offsets of its postings per sub-tile (8 sub-tiles per tile)
  idf       f32[V]
  post_pk   u32[nnz]   (BM25, 4096-doc tiles) packed copy for the f32 pre-filter of the search path: the f32 bits of
                       post_val rounded to 11 mantissa bits | the 12-bit doc offset in its tile (csrc/score_approx.cu)
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi

__all__ = ["TermMajorIndex", "pack_queries", "queries_from_dense", "reference_idf", "reference_avgdl",
           "set_fused_selection", "set_fused_cap", "set_approx_prefilter"]


def set_approx_prefilter(enabled: bool) -> None:
    """Profiling / test hook: False makes search score every posting in f64 (the round-1 kernel) even when the index
    carries the packed copy for the f32 pre-filter (results are bit-identical either way)."""
    _abi.lib.b2r_set_approx_prefilter(1 if enabled else 0)


def set_fused_cap(cap: int) -> None:
    """Test hook: candidate-list capacity of the fused search path (0 = the library's own choice).  A tiny value
    makes every list overflow and sends every query through the exhaustive fallback (results are identical)."""
    _abi.lib.b2r_set_fused_cap(int(cap))


def set_fused_selection(enabled: bool) -> None:
    """Profiling / test hook: False makes search use the plain "score everything, then select" kernels
    instead of the fused-selection pipeline (both are exact)."""
    _abi.lib.b2r_set_fused_selection(1 if enabled else 0)


def _cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b200ret needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def reference_idf(indices: np.ndarray, n_docs: int, n_vocab: int, df: Optional[np.ndarray] = None) -> np.ndarray:
    """retrieval.py:187-189: df = bincount(indices); idf = log((N - df + 0.5)/(df + 0.5)) as f32."""
    if df is None:
        df = np.bincount(indices, minlength=n_vocab)
    return np.log((n_docs - df + 0.5) / (df + 0.5)).astype(np.float32)


def reference_avgdl(doc_lengths: np.ndarray) -> float:
    """retrieval.py:190: float(np.mean(f32 array))."""
    return float(np.mean(np.asarray(doc_lengths, dtype=np.float32)))


def pack_queries(queries: Iterable[Tuple[Sequence[int], Sequence[float]]]):
    """[(term ids, weights)] -> CSR-style (q_ptr i32[Q+1], q_terms i32, q_weights f32).

    Mirrors the dense query_tf vector of the reference (retrieval.py:241-249, :67): one weight per
    term (a later duplicate overwrites an earlier one), only weights > 0 take part, terms ascending.
    """
    ptr = [0]
    terms, weights = [], []
    for t, w in queries:
        t = np.asarray(t, dtype=np.int64).reshape(-1)
        w = np.asarray(w, dtype=np.float32).reshape(-1)
        if len(t):
            # last write wins, like query_tf[term] = w
            _, last = np.unique(t[::-1], return_index=True)
            keep = len(t) - 1 - last
            t, w = t[keep], w[keep]
            order = np.argsort(t, kind="stable")
            t, w = t[order], w[order]
            pos = w > 0
            t, w = t[pos], w[pos]
        terms.append(t.astype(np.int32))
        weights.append(w)
        ptr.append(ptr[-1] + len(t))
    q_terms = np.concatenate(terms) if terms else np.zeros(0, np.int32)
    q_weights = np.concatenate(weights) if weights else np.zeros(0, np.float32)
    return np.asarray(ptr, np.int32), q_terms.astype(np.int32), q_weights.astype(np.float32)


def queries_from_dense(query_tf: np.ndarray):
    """Dense query_tf [V] or [Q, V] (the reference's kernel argument) -> packed queries."""
    q = np.asarray(query_tf, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    rows, cols = np.nonzero(q > 0)
    ptr = np.zeros(q.shape[0] + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=q.shape[0]), out=ptr[1:])
    return ptr, cols.astype(np.int32), q[rows, cols].astype(np.float32)


class TermMajorIndex:
    """One shard of documents laid out term-major in HBM (see module docstring)."""

    def __init__(self):
        self.device = None
        self.kind = "bm25"
        self.n_docs = 0
        self.n_vocab = 0
        self.nnz = 0
        self.tile_docs = 0
        self.n_tiles = 0
        self.doc_id_base = 0
        self.k1 = 1.2
        self.b = 0.75
        self.avgdl = 0.0
        self.idf_host: Optional[np.ndarray] = None
        self._desc = _abi.B2RIndex()
        self._bufs = {}
        # Concurrency (the reference's search_bm25 may be called from several threads): the workspace is kept per
        # CUDA stream, so searches enqueued on different streams never share scratch memory; search_host, which
        # also owns pinned staging buffers, holds _lock from the fill to the copy-out.
        self._ws_by_stream = {}
        self._pinned = {}
        self._lock = threading.Lock()
        self.workspace_cap_bytes = 16 << 30
        self.prefilter_u_max: Optional[float] = None   # largest packed value when the f32 pre-filter copy exists

    # ------------------------------------------------------------------ build
    @classmethod
    def from_csr(cls, data, indices, indptr, doc_lengths=None, *, n_vocab: int, idf=None, avgdl=None,
                 k1: float = 1.2, b: float = 0.75, kind: str = "bm25", doc_id_base: int = 0,
                 tile_docs: int = 4096, device=None, prefilter: bool = True) -> "TermMajorIndex":
        """Build from a doc-major CSR (numpy arrays or CUDA tensors).

        idf / avgdl default to the reference's host expressions over THIS CSR; a doc-sharded build
        passes the global values instead (see dist.py).  prefilter=False skips the packed copy of the search
        path's f32 pre-filter (an index that only serves dense score output does not need it).
        """
        self = cls()
        dev = self.device = _cuda_device(device)
        if kind not in ("bm25", "impact"):
            raise ValueError(f"unknown index kind {kind!r}")
        self.kind = kind
        n_docs = int(len(indptr) - 1)
        if n_docs <= 0:
            raise ValueError("Empty corpus provided")
        nnz = int(indptr[-1])
        self.n_docs, self.n_vocab, self.nnz = n_docs, int(n_vocab), nnz
        self.tile_docs, self.doc_id_base = int(tile_docs), int(doc_id_base)
        self.n_tiles = (n_docs + self.tile_docs - 1) // self.tile_docs
        self.k1, self.b = float(k1), float(b)

        if kind == "bm25":
            if doc_lengths is None:
                raise ValueError("a BM25 index needs doc_lengths")
            if avgdl is None:
                avgdl = reference_avgdl(_to_numpy(doc_lengths))
        self.avgdl = float(avgdl) if avgdl is not None else 0.0
        if idf is None:
            idf = reference_idf(_to_numpy(indices), n_docs, self.n_vocab)
        self.idf_host = np.ascontiguousarray(_to_numpy(idf), dtype=np.float32)
        if len(self.idf_host) != self.n_vocab:
            raise ValueError("idf must have n_vocab entries")

        sizes = _abi.B2RIndexSizes()
        kind_id = _abi.KIND_BM25 if kind == "bm25" else _abi.KIND_IMPACT
        _abi.check(_abi.lib.b2r_index_sizes_for(nnz, n_docs, self.n_vocab, self.tile_docs, kind_id, C.byref(sizes)),
                   "index sizes")
        b_ = self._bufs
        b_["post_doc"] = torch.empty(sizes.post_doc_bytes, dtype=torch.uint8, device=dev)
        b_["post_val"] = torch.empty(sizes.post_val_bytes, dtype=torch.uint8, device=dev)
        b_["blk_ptr"] = torch.empty(sizes.blk_ptr_bytes, dtype=torch.uint8, device=dev)
        b_["dense_id"] = torch.empty(sizes.dense_id_bytes, dtype=torch.uint8, device=dev)
        b_["dense_ptr"] = torch.empty(sizes.dense_ptr_bytes, dtype=torch.uint8, device=dev)
        b_["idf"] = torch.from_numpy(self.idf_host).to(dev)
        scratch = torch.empty(sizes.scratch_bytes, dtype=torch.uint8, device=dev)

        d = self._desc
        d.n_docs, d.doc_id_base, d.nnz = n_docs, self.doc_id_base, nnz
        d.n_vocab, d.tile_docs, d.n_tiles, d.kind = self.n_vocab, self.tile_docs, self.n_tiles, kind_id
        d.post_doc, d.post_val, d.blk_ptr = b_["post_doc"].data_ptr(), b_["post_val"].data_ptr(), b_["blk_ptr"].data_ptr()
        d.dense_id, d.dense_ptr, d.n_dense_max = b_["dense_id"].data_ptr(), b_["dense_ptr"].data_ptr(), int(sizes.n_dense_max)

        # stage the CSR on the device (freed after the build)
        tf_d = _to_device(data, torch.float32, dev)
        ind_d = _to_device(indices, torch.int32, dev)
        ptr_d = _to_device(indptr, torch.int64, dev)
        dl_d = _to_device(doc_lengths, torch.float32, dev) if doc_lengths is not None else None
        st = _stream_ptr(dev)
        _abi.check(_abi.lib.b2r_index_build(C.byref(d), tf_d.data_ptr(), ind_d.data_ptr(), ptr_d.data_ptr(),
                                            dl_d.data_ptr() if dl_d is not None else None, self.k1, self.b,
                                            self.avgdl if kind == "bm25" else 1.0, scratch.data_ptr(),
                                            scratch.numel(), st), "index build")
        _abi.check(_abi.lib.b2r_index_build_status(scratch.data_ptr(), st), "index build")
        del tf_d, ind_d, ptr_d, dl_d, scratch
        if prefilter:
            self._pack_prefilter()
        return self

    def _pack_prefilter(self) -> None:
        """Derive the packed copy the search path's f32 pre-filter reads (BM25 with 4096-document tiles only; it
        is derived data: never saved, rebuilt here after a build or a load)."""
        if self.kind != "bm25" or self.tile_docs != 4096 or self.nnz == 0:
            return
        d = self._desc
        buf = torch.empty(int(_abi.lib.b2r_index_pack_bytes(self.nnz)), dtype=torch.uint8, device=self.device)
        d.post_pk = buf.data_ptr()
        st = _stream_ptr(self.device)
        _abi.check(_abi.lib.b2r_index_pack(C.byref(d), st), "index pack")
        u_max = C.c_float(0.0)
        rc = _abi.lib.b2r_index_pack_status(C.byref(d), st, C.byref(u_max))
        if rc == -4:                 # B2R_ERR_UNSUPPORTED: a value is not finite -- f64 scoring only
            d.post_pk = None
            return
        _abi.check(rc, "index pack")
        self._bufs["post_pk"] = buf
        self.prefilter_u_max = float(u_max.value)

    # ------------------------------------------------------------------ on-disk form
    _IO_CHUNK = 64 << 20

    def save(self, path) -> int:
        """Write the HBM layout verbatim to `path` (format: include/b200ret.h, b2r_index_file_header): a 4096-byte
        header, then post_doc / post_val / blk_ptr / dense_id / dense_ptr / idf, each 4096-aligned and
        checksummed.  The reference can only cache the doc-major CSR (.npz, evaluate_rag_pipeline.py:280-312)
        and rebuilds everything else on load; this file goes back to HBM with no re-layout.  Returns the file size."""
        hdr = _abi.B2RIndexFileHeader()
        d = self._desc
        hdr.n_docs, hdr.doc_id_base, hdr.nnz = d.n_docs, d.doc_id_base, d.nnz
        hdr.n_vocab, hdr.tile_docs, hdr.kind, hdr.n_dense_max = d.n_vocab, d.tile_docs, d.kind, d.n_dense_max
        hdr.k1, hdr.b, hdr.avgdl = self.k1, self.b, self.avgdl
        total = C.c_uint64(0)
        _abi.check(_abi.lib.b2r_index_file_layout(C.byref(hdr), C.byref(total)), "index file layout")
        torch.cuda.synchronize(self.device)
        stage = torch.empty(self._IO_CHUNK, dtype=torch.uint8).pin_memory()
        with open(path, "wb") as f:
            f.truncate(total.value)
            for i, name in enumerate(_abi.SEC_NAMES):
                sec = hdr.sections[i]
                buf = self._bufs[name].view(torch.uint8).reshape(-1)
                if buf.numel() != sec.bytes:
                    raise _abi.B2RError(f"index file: buffer {name} has {buf.numel()} bytes, layout says {sec.bytes}")
                f.seek(sec.offset)
                # chained checksum: every chunk is folded into the running value of the section
                acc = 0
                for o in range(0, int(sec.bytes), self._IO_CHUNK):
                    n = min(self._IO_CHUNK, int(sec.bytes) - o)
                    stage[:n].copy_(buf[o:o + n])
                    acc = _chain_checksum(acc, stage, n)
                    f.write(memoryview(stage.numpy())[:n])
                sec.checksum = acc
            f.seek(0)
            f.write(bytes(hdr))
        return int(total.value)

    @classmethod
    def load(cls, path, *, device=None, doc_id_base: Optional[int] = None, verify: bool = True) -> "TermMajorIndex":
        """Read an index written by save() straight into HBM (no CSR, no build kernels).  `doc_id_base` re-bases
        the shard (the buffers hold shard-local document indices); `verify` checks every section's checksum."""
        import os
        self = cls()
        dev = self.device = _cuda_device(device)
        size = os.path.getsize(path)
        hdr = _abi.B2RIndexFileHeader()
        with open(path, "rb") as f:
            raw = f.read(C.sizeof(hdr))
            if len(raw) < C.sizeof(hdr):
                raise ValueError(f"{path}: not a b200ret index file (too short)")
            C.memmove(C.byref(hdr), raw, C.sizeof(hdr))
            _abi.check(_abi.lib.b2r_index_file_check(C.byref(hdr), size), f"{path}")
            stage = torch.empty(cls._IO_CHUNK, dtype=torch.uint8).pin_memory()
            for i, name in enumerate(_abi.SEC_NAMES):
                sec = hdr.sections[i]
                buf = torch.empty(int(sec.bytes), dtype=torch.uint8, device=dev)
                f.seek(sec.offset)
                acc = 0
                for o in range(0, int(sec.bytes), cls._IO_CHUNK):
                    n = min(cls._IO_CHUNK, int(sec.bytes) - o)
                    got = f.readinto(memoryview(stage.numpy())[:n])
                    if got != n:
                        raise ValueError(f"{path}: section {name} is truncated")
                    if verify:
                        acc = _chain_checksum(acc, stage, n)
                    buf[o:o + n].copy_(stage[:n])
                    torch.cuda.current_stream(dev).synchronize()      # the staging buffer is reused
                if verify and acc != sec.checksum:
                    raise ValueError(f"{path}: checksum mismatch in section {name} (file is corrupt)")
                self._bufs[name] = buf
        self.kind = "bm25" if hdr.kind == _abi.KIND_BM25 else "impact"
        self.n_docs, self.n_vocab, self.nnz = int(hdr.n_docs), int(hdr.n_vocab), int(hdr.nnz)
        self.tile_docs, self.n_tiles = int(hdr.tile_docs), int(hdr.n_tiles)
        self.doc_id_base = int(hdr.doc_id_base if doc_id_base is None else doc_id_base)
        if self.doc_id_base < 0 or self.doc_id_base + self.n_docs >= 0xFFFFFFFF:
            raise ValueError("doc_id_base out of range")
        self.k1, self.b, self.avgdl = float(hdr.k1), float(hdr.b), float(hdr.avgdl)
        self._bufs["idf"] = self._bufs["idf"].view(torch.float32)
        self.idf_host = self._bufs["idf"].cpu().numpy()
        d, b_ = self._desc, self._bufs
        d.n_docs, d.doc_id_base, d.nnz = self.n_docs, self.doc_id_base, self.nnz
        d.n_vocab, d.tile_docs, d.n_tiles, d.kind = self.n_vocab, self.tile_docs, self.n_tiles, int(hdr.kind)
        d.post_doc, d.post_val, d.blk_ptr = b_["post_doc"].data_ptr(), b_["post_val"].data_ptr(), b_["blk_ptr"].data_ptr()
        d.dense_id, d.dense_ptr, d.n_dense_max = b_["dense_id"].data_ptr(), b_["dense_ptr"].data_ptr(), int(hdr.n_dense_max)
        self._pack_prefilter()
        return self

    def view(self) -> "TermMajorIndex":
        """A second handle on the SAME device buffers with its own pinned staging buffers, workspaces and lock: lets
        several host threads run search_host concurrently (each on its own CUDA stream) against one resident index."""
        v = TermMajorIndex()
        for name in ("device", "kind", "n_docs", "n_vocab", "nnz", "tile_docs", "n_tiles", "doc_id_base", "k1", "b",
                     "avgdl", "idf_host", "workspace_cap_bytes", "prefilter_u_max"):
            setattr(v, name, getattr(self, name))
        v._bufs = self._bufs                    # shared, immutable after the build
        C.memmove(C.byref(v._desc), C.byref(self._desc), C.sizeof(self._desc))
        return v

    # ------------------------------------------------------------------ properties
    @property
    def padded_docs(self) -> int:
        return self.n_tiles * self.tile_docs

    @property
    def idf(self) -> torch.Tensor:
        return self._bufs["idf"]

    def set_idf(self, idf) -> None:
        """Replace the idf vector (it is a scoring-time input, like in the reference signature)."""
        idf = np.ascontiguousarray(_to_numpy(idf), dtype=np.float32)
        if len(idf) != self.n_vocab:
            raise ValueError("idf must have n_vocab entries")
        self.idf_host = idf
        self._bufs["idf"] = torch.from_numpy(idf).to(self.device)

    def device_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._bufs.values())

    # ------------------------------------------------------------------ workspace
    def _workspace(self, n_queries: int, k: int, extra: int = 0) -> torch.Tensor:
        mn, full = C.c_size_t(0), C.c_size_t(0)
        _abi.check(_abi.lib.b2r_search_workspace(C.byref(self._desc), n_queries, k, C.byref(mn), C.byref(full)),
                   "search workspace")
        want = min(full.value, max(mn.value, self.workspace_cap_bytes)) + extra
        key = _stream_ptr(self.device)
        ws = self._ws_by_stream.get(key)
        if ws is None or ws.numel() < want:
            self._ws_by_stream.pop(key, None)
            ws = None
            ws = self._ws_by_stream[key] = torch.empty(want, dtype=torch.uint8, device=self.device)
        return ws

    @property
    def _ws(self) -> Optional[torch.Tensor]:
        """Workspace of the current stream (None before the first search on it)."""
        return self._ws_by_stream.get(_stream_ptr(self.device))

    # ------------------------------------------------------------------ search (device buffers)
    def search(self, q_ptr, q_terms, q_weights, k: int, *, return_keys: bool = False):
        """Batched BM25 / impact scoring + top-k.  Returns (idx i64[Q,k], val f32[Q,k]) CUDA tensors
        (idx = global doc index, -1 / -inf where fewer than k docs exist) and optionally the keys."""
        dev = self.device
        q_ptr = _to_device(q_ptr, torch.int32, dev)
        q_terms = _to_device(q_terms, torch.int32, dev)
        q_weights = _to_device(q_weights, torch.float32, dev)
        nq = int(q_ptr.numel() - 1)
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)   # u64 payload
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        val = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ws = self._workspace(nq, k)
        _abi.check(_abi.lib.b2r_search_batch(C.byref(self._desc), q_ptr.data_ptr(), q_terms.data_ptr(),
                                             q_weights.data_ptr(), self.idf.data_ptr(), nq, k, None, 0,
                                             keys.data_ptr(), idx.data_ptr(), val.data_ptr(), ws.data_ptr(),
                                             ws.numel(), _stream_ptr(dev)), "search")
        return (idx, val, keys) if return_keys else (idx, val)

    def score_dense(self, q_ptr, q_terms, q_weights, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Dense f32 scores [Q, n_docs] (a view of a padded [Q, n_tiles*tile_docs] buffer; pass the
        padded buffer as `out` to reuse it)."""
        dev = self.device
        q_ptr = _to_device(q_ptr, torch.int32, dev)
        q_terms = _to_device(q_terms, torch.int32, dev)
        q_weights = _to_device(q_weights, torch.float32, dev)
        nq = int(q_ptr.numel() - 1)
        if out is None:
            out = torch.empty((nq, self.padded_docs), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (nq, self.padded_docs) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous f32 [Q, n_tiles*tile_docs] CUDA tensor")
        _abi.check(_abi.lib.b2r_search_batch(C.byref(self._desc), q_ptr.data_ptr(), q_terms.data_ptr(),
                                             q_weights.data_ptr(), self.idf.data_ptr(), nq, 0, out.data_ptr(),
                                             self.padded_docs, None, None, None, None, 0, _stream_ptr(dev)),
                   "score")
        return out[:, :self.n_docs]

    # ------------------------------------------------------------------ search (host buffers)
    def _pinned_buf(self, name: str, n: int, dtype) -> torch.Tensor:
        t = self._pinned.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[name] = t
        return t

    def search_host(self, q_ptr: np.ndarray, q_terms: np.ndarray, q_weights: np.ndarray, k: int):
        """Plugin-facing call with HOST buffers: host->device copy of the queries, scoring, selection
        and device->host copy of (idx, val) all happen inside b2r_search_batch_host.
        Returns numpy (idx i64[Q,k], val f32[Q,k])."""
        nq = int(len(q_ptr) - 1)
        nt = int(len(q_terms))
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        with self._lock:      # the pinned staging buffers are per index: one host-buffer search at a time
            hp = self._pinned_buf("q_ptr", nq + 1, torch.int32)
            ht = self._pinned_buf("q_terms", nt, torch.int32)
            hw = self._pinned_buf("q_weights", nt, torch.float32)
            hi = self._pinned_buf("idx", nq * k, torch.int64)
            hv = self._pinned_buf("val", nq * k, torch.float32)
            hp[:nq + 1].numpy()[:] = q_ptr
            ht[:nt].numpy()[:] = q_terms
            hw[:nt].numpy()[:] = q_weights
            extra = int(_abi.lib.b2r_search_host_extra_bytes(nq, nt, k))
            ws = self._workspace(nq, k, extra)
            _abi.check(_abi.lib.b2r_search_batch_host(C.byref(self._desc), hp.data_ptr(), ht.data_ptr(), hw.data_ptr(),
                                                      self.idf.data_ptr(), nq, k, None, hi.data_ptr(), hv.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), _stream_ptr(self.device)),
                       "search (host buffers)")
            return (hi[:nq * k].numpy().reshape(nq, k).copy(), hv[:nq * k].numpy().reshape(nq, k).copy())

    # ------------------------------------------------------------------ traffic model
    def postings_touched(self, q_ptr: np.ndarray, q_terms: np.ndarray, df: np.ndarray) -> int:
        """P = sum over queries and their terms of df[t] on this shard."""
        return int(np.asarray(df, dtype=np.int64)[np.asarray(q_terms, dtype=np.int64)].sum())


def _chain_checksum(acc: int, stage: torch.Tensor, n: int) -> int:
    """Running checksum of a section written/read in chunks: fold the chunk's b2r_checksum64 into `acc`."""
    c = int(_abi.lib.b2r_checksum64(stage.data_ptr(), n))
    word = np.array([acc, c], dtype=np.uint64)
    return int(_abi.lib.b2r_checksum64(word.ctypes.data, 16))


def _to_numpy(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _to_device(x, dtype: torch.dtype, dev: torch.device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    a = np.ascontiguousarray(np.asarray(x))
    t = torch.from_numpy(a) if a.size else torch.zeros(0, dtype=dtype)
    return t.to(device=dev, dtype=dtype).contiguous()
