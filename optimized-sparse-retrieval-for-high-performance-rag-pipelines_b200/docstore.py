"""Document store shim at the drop-in boundary.

The reference's RetrievalService opens a MemoryIndex file in its constructor
(rag_system/core/retrieval.py:103; format in rag_system/core/memory_index.py:22-34, 159-195) and
fetches result texts from it (retrieval.py:356-462).  Text blobs are NOT on the scoring path and
the store itself is out of scope (the fetch that FOLLOWS a search is); this
module reads and writes the same on-disk layout so that an index file made by either side opens on
the other, and serves the batched fetch behind get_documents / get_search_results: one pass over the
mmap in file order, every distinct document read once, zlib inflation on a thread pool:

  file header  "QQI"  : num_docs u64, data_size u64, max_id_len u32
  per document "QQQB" : id_len, text_len, title_len, flags ; id ; text ; title ; u64 meta_len ; meta
  flags bit0/1/2 = text/title/metadata zlib-compressed; metadata is a pickled dict.
"""
from __future__ import annotations

import mmap
import pickle
import struct
import zlib
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Union

_FILE_HDR = struct.Struct("QQI")
_DOC_HDR = struct.Struct("QQQB")
_U64 = struct.Struct("Q")
_Z_TEXT, _Z_TITLE, _Z_META = 1, 2, 4
_Z_MIN = 256


@dataclass
class Document:
    """Same fields as the reference's Document (rag_system/core/data_processor.py:14-19)."""
    id: str
    text: str
    title: Optional[str] = None
    metadata: Optional[Dict] = None

    def validate(self) -> bool:
        return bool(self.id and self.text)

    def to_dict(self) -> Dict:
        d = {"id": self.id, "text": self.text}
        if self.title:
            d["title"] = self.title
        if self.metadata:
            d["metadata"] = self.metadata
        return d


def _maybe_z(blob: bytes, level: int):
    if len(blob) > _Z_MIN:
        z = zlib.compress(blob, level)
        if len(z) < len(blob):
            return z, True
    return blob, False


class MemoryIndex:
    """mmap-backed document store; constructor contract of memory_index.py:201-226
    (`create=True` writes an empty file, otherwise a missing file raises FileNotFoundError)."""

    def __init__(self, index_path: Union[str, Path], create: bool = False, max_id_length: int = 64,
                 cache_size: int = 1000):
        self.index_path = Path(index_path)
        self.max_id_length = max_id_length
        self.cache_size = cache_size
        self._file = None
        self._map = None
        self._where: Dict[str, int] = {}      # doc id -> offset of its "QQQB" header
        if create:
            self.index_path.parent.mkdir(parents=True, exist_ok=True)
            with open(self.index_path, "wb") as f:
                f.write(_FILE_HDR.pack(0, 0, max_id_length))
        self._open()

    # -- file handling
    def _open(self) -> None:
        if not self.index_path.exists():
            raise FileNotFoundError(f"Index file not found: {self.index_path}")
        self._file = open(self.index_path, "r+b")
        self._map = mmap.mmap(self._file.fileno(), 0, access=mmap.ACCESS_READ)
        if len(self._map) < _FILE_HDR.size:
            raise ValueError("Invalid index file: too small")
        n_docs, _, self.max_id_length = _FILE_HDR.unpack_from(self._map, 0)
        self._where.clear()
        off, end = _FILE_HDR.size, len(self._map)
        for _ in range(n_docs):
            if off + _DOC_HDR.size > end:
                break                                   # truncated file: keep what is readable
            id_len, text_len, title_len, _flags = _DOC_HDR.unpack_from(self._map, off)
            body = off + _DOC_HDR.size
            if body + id_len > end:
                break
            doc_id = self._map[body:body + id_len].decode("utf-8").rstrip("\x00")
            meta_at = body + id_len + text_len + title_len
            if meta_at + 8 > end:
                break
            (meta_len,) = _U64.unpack_from(self._map, meta_at)
            self._where[doc_id] = off
            off = meta_at + 8 + meta_len

    def _close_map(self) -> None:
        if self._map is not None:
            self._map.close()
            self._map = None
        if self._file is not None:
            self._file.close()
            self._file = None

    def close(self) -> None:
        self._close_map()

    # -- writes (rewrite-all, like the reference's add_documents)
    def add_documents(self, documents: List[Document], compression_level: int = 6, **_ignored) -> None:
        if not documents:
            return
        docs = [self.get_document(d) for d in list(self._where)] + list(documents)
        self._close_map()
        blobs = []
        for doc in docs:
            did = doc.id.encode("utf-8")[: self.max_id_length]
            flags = 0
            text, z = _maybe_z((doc.text or "").encode("utf-8"), compression_level)
            flags |= _Z_TEXT if z else 0
            title, z = _maybe_z((doc.title or "").encode("utf-8"), compression_level)
            flags |= _Z_TITLE if z else 0
            meta, z = _maybe_z(pickle.dumps(doc.metadata or {}), compression_level)
            flags |= _Z_META if z else 0
            blobs.append(_DOC_HDR.pack(len(did), len(text), len(title), flags) + did + text + title +
                         _U64.pack(len(meta)) + meta)
        with open(self.index_path, "wb") as f:
            f.write(_FILE_HDR.pack(len(blobs), sum(map(len, blobs)), self.max_id_length))
            for b in blobs:
                f.write(b)
        self._open()

    # -- reads
    def get_document(self, doc_id: str) -> Optional[Document]:
        off = self._where.get(doc_id)
        if off is None or self._map is None:
            return None
        id_len, text_len, title_len, flags = _DOC_HDR.unpack_from(self._map, off)
        p = off + _DOC_HDR.size + id_len
        text = self._map[p:p + text_len]
        p += text_len
        title = self._map[p:p + title_len]
        p += title_len
        (meta_len,) = _U64.unpack_from(self._map, p)
        meta = self._map[p + 8:p + 8 + meta_len]
        text = (zlib.decompress(text) if flags & _Z_TEXT else text).decode("utf-8")
        title = (zlib.decompress(title) if flags & _Z_TITLE else title).decode("utf-8")
        metadata = pickle.loads(zlib.decompress(meta) if flags & _Z_META else meta)
        return Document(id=doc_id, text=text, title=title, metadata=metadata)

    def _raw(self, off: int):
        """(flags, text bytes, title bytes, metadata bytes) of the record at `off`, still compressed."""
        id_len, text_len, title_len, flags = _DOC_HDR.unpack_from(self._map, off)
        p = off + _DOC_HDR.size + id_len
        text = bytes(self._map[p:p + text_len])
        p += text_len
        title = bytes(self._map[p:p + title_len])
        p += title_len
        (meta_len,) = _U64.unpack_from(self._map, p)
        return flags, text, title, bytes(self._map[p + 8:p + 8 + meta_len])

    @staticmethod
    def _decode(item) -> Document:
        doc_id, (flags, text, title, meta) = item
        text = (zlib.decompress(text) if flags & _Z_TEXT else text).decode("utf-8")
        title = (zlib.decompress(title) if flags & _Z_TITLE else title).decode("utf-8")
        metadata = pickle.loads(zlib.decompress(meta) if flags & _Z_META else meta)
        return Document(id=doc_id, text=text, title=title, metadata=metadata)

    def get_documents(self, doc_ids: List[str], num_workers: int = 4) -> List[Optional[Document]]:
        """memory_index.py:413-468, batched: the distinct ids of the call are looked up once, their records are read
        in FILE ORDER (one forward pass over the mmap instead of one random access per id and per worker), and the
        inflation runs on `num_workers` threads (zlib releases the GIL).  Result order = request order, None for an
        unknown id."""
        if self._map is None or not doc_ids:
            return [None] * len(doc_ids)
        wanted = {}
        for d in doc_ids:
            if d not in wanted:
                off = self._where.get(d)
                if off is not None:
                    wanted[d] = off
        raw = [(d, self._raw(off)) for d, off in sorted(wanted.items(), key=lambda kv: kv[1])]
        if num_workers > 1 and len(raw) >= 64:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=num_workers) as pool:
                docs = list(pool.map(self._decode, raw, chunksize=max(1, len(raw) // (4 * num_workers))))
        else:
            docs = [self._decode(r) for r in raw]
        by_id = {doc.id: doc for doc in docs}
        return [by_id.get(d) for d in doc_ids]

    def get_document_count(self) -> int:
        return len(self._where)

    def get_document_ids(self) -> List[str]:
        """memory_index.py:474-476."""
        return list(self._where.keys())

    def contains(self, doc_id: str) -> bool:
        """memory_index.py:478-480."""
        return doc_id in self._where

    def get_index_stats(self) -> Dict[str, object]:
        """memory_index.py:482-499 (same keys; there is no LRU cache in this shim, so its stats are empty)."""
        size = self.index_path.stat().st_size if self.index_path.exists() else 0
        n = len(self._where)
        return {"num_documents": n, "file_size_mb": size / (1024 * 1024),
                "average_doc_size_bytes": (size - _FILE_HDR.size) / n if n else 0,
                "cache_stats": {"size": 0, "capacity": self.cache_size}, "compression_enabled": True,
                "memory_mapped": self._map is not None}

    def __contains__(self, doc_id: str) -> bool:
        return doc_id in self._where

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
