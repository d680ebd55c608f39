"""Streaming reader / writer of the reference's embedding files.

On-disk format (``src/openmatch/retriever/dense_retriever.py:84-86,160-161``, read back at ``:96-101,171-176`` and by
``scripts/split_embeddings.py:14-21``): ``pickle.dump((encoded: np.float32 [n, d] C-order, lookup: list[str]),
protocol=4)`` in files named ``embeddings.{corpus,query}.rank.{r}``.  The reference materialises the whole matrix on
the host on both sides (a 27 GB ``bytes`` object inside the unpickler for an 8.8 M x 768 shard).  This module keeps the
bytes of the file identical in meaning — ``pickle.load`` of our files returns the same ``(ndarray, list)``, and we read
files written by the reference or by ``split_embeddings.py`` — but never holds more than one chunk on the host:

* :func:`write_embedding_file` emits the pickle opcodes by hand and streams the matrix payload chunk by chunk (from a
  CUDA tensor: one pinned D2H chunk at a time);
* :class:`EmbeddingFile` parses the pickle header (numpy's layout or ours), exposes the payload as a read-only
  ``np.memmap`` and un-pickles only the id list; ``chunks()`` feeds ``FlatIPIndex.add`` without a host copy of the
  whole matrix;
* :func:`split_embedding_file` is ``scripts/split_embeddings.py`` (round-robin rows ``split::num_splits``) on top of both.

Anything the header parser does not recognise falls back to plain ``pickle.load`` (same result, reference memory cost).
"""
from __future__ import annotations

import io
import os
import pickle
import pickletools
import struct
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

_CHUNK_ROWS = 1 << 16


# ------------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------------
def _op_int(v: int) -> bytes:
    if 0 <= v < 256:
        return b"K" + bytes([v])
    if 0 <= v < 65536:
        return b"M" + struct.pack("<H", v)
    if -2 ** 31 <= v < 2 ** 31:
        return b"J" + struct.pack("<i", v)
    raw = v.to_bytes((v.bit_length() + 8) // 8, "little", signed=True)
    return b"\x8a" + bytes([len(raw)]) + raw  # LONG1


def _op_str(s: str) -> bytes:
    raw = s.encode("utf-8", "surrogatepass")
    if len(raw) < 256:
        return b"\x8c" + bytes([len(raw)]) + raw  # SHORT_BINUNICODE
    return b"X" + struct.pack("<I", len(raw)) + raw  # BINUNICODE


def _global(module: str, name: str) -> bytes:
    return _op_str(module) + _op_str(name) + b"\x93"  # STACK_GLOBAL


def _header(n: int, d: int) -> bytes:
    """opcodes up to and including the BINBYTES8 length of the float32 [n, d] payload"""
    out = [b"\x80\x04",  # PROTO 4 (unframed: frames are optional)
           _global("numpy.core.multiarray", "_reconstruct"),  # importable under numpy 1.x and 2.x alike
           _global("numpy", "ndarray"), _op_int(0), b"\x85", b"C\x01b", b"\x87", b"R",  # (ndarray, (0,), b'b') REDUCE
           b"(", _op_int(1), _op_int(n), _op_int(d), b"\x86",  # MARK 1 (n, d)
           _global("numpy", "dtype"), _op_str("f4"), b"\x89\x88\x87R",  # dtype('f4', False, True)
           b"(", _op_int(3), _op_str("<"), b"NNN", _op_int(-1), _op_int(-1), _op_int(0), b"tb",  # dtype state, BUILD
           b"\x89",  # is_fortran = False
           b"\x8e" + struct.pack("<Q", n * d * 4)]  # BINBYTES8 + payload length
    return b"".join(out)


def _rows_chunks(rows, chunk_rows: int) -> Iterator[np.ndarray]:
    """float32 C-contiguous host chunks of a numpy array / torch tensor (CPU or CUDA) / iterable of such"""
    try:
        import torch
    except ImportError:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(rows, torch.Tensor):
        if rows.is_cuda:
            stage = torch.empty((min(chunk_rows, max(1, rows.shape[0])), rows.shape[1]), dtype=torch.float32).pin_memory()
            for lo in range(0, rows.shape[0], chunk_rows):
                part = rows[lo:lo + chunk_rows]
                stage[: part.shape[0]].copy_(part)  # synchronous D2H of one chunk
                yield stage[: part.shape[0]].numpy()
            return
        rows = rows.detach().numpy()
    if isinstance(rows, np.ndarray):
        for lo in range(0, rows.shape[0], chunk_rows):
            yield np.ascontiguousarray(rows[lo:lo + chunk_rows], dtype=np.float32)
        return
    for part in rows:
        yield from _rows_chunks(part, chunk_rows)


def write_embedding_file(path: str, rows, ids: Sequence[str], n: Optional[int] = None, d: Optional[int] = None,
                         chunk_rows: int = _CHUNK_ROWS) -> None:
    """``pickle.dump((float32 [n, d], list(ids)), f, protocol=4)`` without a second copy of the matrix.
    ``rows``: numpy array, torch tensor (CUDA: streamed through one pinned chunk) or an iterable of row blocks (then
    ``n`` and ``d`` must be given)."""
    if n is None or d is None:
        n, d = int(rows.shape[0]), int(rows.shape[1])
    if len(ids) != n:
        raise ValueError("%d ids for %d rows" % (len(ids), n))
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_header(n, d))
        written = 0
        for chunk in _rows_chunks(rows, chunk_rows):
            if chunk.ndim != 2 or chunk.shape[1] != d:
                raise ValueError("row block of shape %s in a [%d, %d] file" % (chunk.shape, n, d))
            f.write(memoryview(chunk).cast("B"))
            written += chunk.shape[0]
        if written != n:
            raise ValueError("wrote %d rows, header promised %d" % (written, n))
        f.write(b"tb")  # TUPLE (ndarray state) BUILD
        f.write(b"]")   # EMPTY_LIST
        for lo in range(0, n, 1000):  # APPENDS in batches like pickle's save_list
            f.write(b"(" + b"".join(_op_str(str(s)) for s in ids[lo:lo + 1000]) + b"e")
        f.write(b"\x86.")  # TUPLE2 STOP
    os.replace(tmp, path)


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------
class _Sink:
    """stands in for the ndarray while only the id list of a file is un-pickled"""

    def __setstate__(self, state):
        pass


def _sink():
    return _Sink()


class EmbeddingFile:
    """One ``embeddings.*.rank.*`` file: ``shape``, ``rows`` (read-only float32 memmap of the payload, no copy),
    ``ids`` (list[str], parsed lazily), ``chunks(rows_per_chunk)``."""

    def __init__(self, path: str):
        self.path = path
        self._ids: Optional[List[str]] = None
        self._fallback = None
        try:
            self._parse_header()
        except Exception:  # unknown layout: the reference's way (whole file through pickle)
            with open(path, "rb") as f:
                enc, ids = pickle.load(f)
            self._fallback = np.ascontiguousarray(enc, dtype=np.float32)
            self._ids = list(ids)
            self.shape = tuple(self._fallback.shape)
            self.streaming = False

    def _parse_header(self):
        with open(self.path, "rb") as f:
            head = f.read(1 << 16)
        pos_after, ints, strs, memo = 0, [], [], 0
        shape = None
        try:
            for op, arg, pos in pickletools.genops(io.BytesIO(head)):
                name = op.name
                if name in ("MEMOIZE",):
                    memo += 1
                if name in ("BININT", "BININT1", "BININT2", "LONG1"):
                    ints.append(int(arg))
                if name in ("SHORT_BINUNICODE", "BINUNICODE"):
                    strs.append(arg)
                if name == "TUPLE2" and shape is None and len(ints) >= 3 and "ndarray" in strs:
                    shape = (ints[-2], ints[-1])  # MARK 1 (n, d): the first 2-tuple after the constructor call
                if name in ("SHORT_BINBYTES", "BINBYTES", "BINBYTES8") and shape is not None and "f4" in strs:
                    # small payload fully inside the prefix we read
                    self._finish(pos + {"SHORT_BINBYTES": 2, "BINBYTES": 5, "BINBYTES8": 9}[name], len(arg), shape, memo)
                    return
                nxt = f"{name}"
                pos_after = None  # recomputed below from the stream position
                del nxt
        except ValueError:
            pass  # ran into the (truncated) payload: the op that failed starts where the last good one ended
        if shape is None or "f4" not in strs or "<" not in strs:
            raise ValueError("not a float32 little-endian ndarray pickle")
        # position of the failing opcode = end of the last complete op: re-walk to find it
        end = 0
        try:
            gen = pickletools.genops(io.BytesIO(head))
            last = None
            for op, arg, pos in gen:
                last = (op, arg, pos)
        except ValueError:
            pass
        # the failing op begins right after `last`: scan forward from last's position for a BINBYTES / BINBYTES8 opcode
        start = last[2] if last else 0
        for p in range(start + 1, min(len(head) - 9, start + 64)):
            if head[p] == 0x8e:  # BINBYTES8
                size = struct.unpack_from("<Q", head, p + 1)[0]
                if size == shape[0] * shape[1] * 4:
                    self._finish(p + 9, size, shape, memo)
                    return
            if head[p] == ord("B"):  # BINBYTES
                size = struct.unpack_from("<I", head, p + 1)[0]
                if size == shape[0] * shape[1] * 4:
                    self._finish(p + 5, size, shape, memo)
                    return
        raise ValueError("payload opcode not found")

    def _finish(self, data_off: int, size: int, shape, memo: int):
        if size != shape[0] * shape[1] * 4:
            raise ValueError("payload size does not match the shape")
        self.shape = (int(shape[0]), int(shape[1]))
        self._data_off, self._size, self._memo = data_off, size, memo
        self.streaming = True

    @property
    def rows(self) -> np.ndarray:
        if self._fallback is not None:
            return self._fallback
        if self.shape[0] == 0:
            return np.zeros(self.shape, np.float32)
        return np.memmap(self.path, dtype="<f4", mode="r", offset=self._data_off, shape=self.shape)

    @property
    def ids(self) -> List[str]:
        if self._ids is None:
            with open(self.path, "rb") as f:
                f.seek(self._data_off + self._size)
                tail = f.read()
            # replay the tail on a stack that looks like the original one: [sink, MARK, ..., payload] with the memo
            # table as long as the original's (MEMOIZE indices are implicit in protocol 4)
            pre = b"\x80\x04" + b"N\x940" * self._memo + _global(__name__, "_sink") + b")R" + b"(N"
            obj = pickle.loads(pre + tail)
            self._ids = list(obj[1])
            if len(self._ids) != self.shape[0]:
                raise ValueError("%s: %d ids for %d rows" % (self.path, len(self._ids), self.shape[0]))
        return self._ids

    def chunks(self, rows_per_chunk: int = _CHUNK_ROWS) -> Iterator[np.ndarray]:
        rows = self.rows
        for lo in range(0, self.shape[0], rows_per_chunk):
            yield rows[lo:lo + rows_per_chunk]


def read_embedding_file(path: str) -> Tuple[np.ndarray, List[str]]:
    """``pickle.load`` equivalent that maps the matrix instead of copying it"""
    ef = EmbeddingFile(path)
    return ef.rows, ef.ids


def split_embedding_file(input_embedding: str, output_embeddings: str, num_splits: int = 2) -> List[str]:
    """``scripts/split_embeddings.py:14-21``: split ``i`` holds rows ``i::num_splits`` and their ids, written as
    ``<output_embeddings>.<i>`` in the same format; streamed, so a shard larger than host memory can be split."""
    ef = EmbeddingFile(input_embedding)
    ids, rows = ef.ids, ef.rows
    outs = []
    for s in range(num_splits):
        n_s = len(range(s, ef.shape[0], num_splits))

        def blocks(s=s):
            for lo in range(0, ef.shape[0], _CHUNK_ROWS * num_splits):
                block = rows[lo:lo + _CHUNK_ROWS * num_splits]
                first = (s - lo) % num_splits
                yield np.ascontiguousarray(block[first::num_splits], dtype=np.float32)

        out = "%s.%d" % (output_embeddings, s)
        write_embedding_file(out, blocks(), ids[s::num_splits], n=n_s, d=ef.shape[1])
        outs.append(out)
    return outs


if __name__ == "__main__":  # drop-in for scripts/split_embeddings.py
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_embedding", type=str)
    ap.add_argument("--output_embeddings", type=str)
    ap.add_argument("--num_splits", type=int, default=2)
    a = ap.parse_args()
    for p in split_embedding_file(a.input_embedding, a.output_embeddings, a.num_splits):
        print(p)
