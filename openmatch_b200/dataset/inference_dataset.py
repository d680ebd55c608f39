"""Streaming inference datasets (reference: ``src/openmatch/dataset/inference_dataset.py``).

Same contract as the reference: ``InferenceDataset.load`` dispatches on the file extension (``.json`` ->
JSON lines, ``.tsv``/``.txt`` -> tab separated with the configured column names), examples are rendered
through the query / doc template, tokenised to a fixed ``max_length`` padding, and ranks take interleaved
blocks of ``batch_size`` examples (:99-115).  Files are streamed with plain Python I/O.
``PretokenizedDataset`` is the device-friendly ingest format: int32 token ids in a ``.npy`` memory map.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterator

import numpy as np
from torch.utils.data import IterableDataset

from ..arguments import DataArguments
from ..utils import fill_template, find_all_markers


def get_idx(obj) -> str:
    example_id = obj.get("_id", None) or obj.get("id", None)
    return str(example_id) if example_id is not None else None


class InferenceDataset(IterableDataset):
    def __init__(self, tokenizer, data_args: DataArguments, is_query: bool = False, final: bool = True,
                 stream: bool = True, batch_size: int = 1, num_processes: int = 1, process_index: int = 0,
                 cache_dir: str = None):
        super().__init__()
        self.tokenizer = tokenizer
        self.data_files = [data_args.query_path] if is_query else [data_args.corpus_path]
        self.max_len = data_args.q_max_len if is_query else data_args.p_max_len
        self.template = data_args.query_template if is_query else data_args.doc_template
        self.all_markers = find_all_markers(self.template)
        self.final, self.stream = final, stream
        self.batch_size, self.num_processes, self.process_index = batch_size, num_processes, process_index

    @classmethod
    def load(cls, tokenizer, data_args: DataArguments, is_query: bool = False, final: bool = True, stream: bool = True,
             batch_size: int = 1, num_processes: int = 1, process_index: int = 0, cache_dir: str = None):
        path = data_args.query_path if is_query else data_args.corpus_path
        ext = os.path.splitext(path)[1]
        target = {".json": JsonlDataset, ".jsonl": JsonlDataset, ".tsv": TsvDataset, ".txt": TsvDataset,
                  ".npy": PretokenizedDataset}.get(ext)
        if target is None:
            raise ValueError("Unsupported dataset file extension {}".format(ext))
        return target(tokenizer=tokenizer, data_args=data_args, is_query=is_query, final=final, stream=stream,
                      batch_size=batch_size, num_processes=num_processes, process_index=process_index,
                      cache_dir=cache_dir)

    def _records(self) -> Iterator[Dict]:
        raise NotImplementedError

    def process_one(self, example):
        text = fill_template(self.template, example, self.all_markers, allow_not_found=True)
        tok = self.tokenizer(text, add_special_tokens=self.final, padding="max_length" if self.final else False,
                             truncation=True, max_length=self.max_len, return_attention_mask=self.final,
                             return_token_type_ids=self.final)
        return {"text_id": get_idx(example), **tok}

    def __iter__(self):
        group = self.batch_size * self.num_processes
        lo, hi = self.process_index * self.batch_size, (self.process_index + 1) * self.batch_size
        pending = []
        for rec in self._records():
            pending.append(rec)
            if len(pending) == group:
                for rec_ in pending[lo:hi]:
                    yield self.process_one(rec_)
                pending = []
        for rec_ in pending[lo:hi]:
            yield self.process_one(rec_)


class JsonlDataset(InferenceDataset):
    def _records(self):
        with open(self.data_files[0]) as f:
            for line in f:
                if line.strip():
                    yield json.loads(line)


class TsvDataset(InferenceDataset):
    def __init__(self, tokenizer, data_args: DataArguments, is_query: bool = False, **kwargs):
        super().__init__(tokenizer, data_args, is_query, **kwargs)
        self.all_columns = (data_args.query_column_names if is_query else data_args.doc_column_names).split(",")

    def _records(self):
        with open(self.data_files[0]) as f:
            for line in f:
                yield dict(zip(self.all_columns, line.rstrip("\n").split("\t")))


class PretokenizedDataset(InferenceDataset):
    """``<name>.npy``: int32 ``[n, L]`` token ids (0 = padding); optional ``<name>.ids.txt`` with one id per
    row.  No tokenizer involved: rows are sliced to ``max_len`` and the mask is ``ids != 0``.

    Two ways out: the reference's per-example iterator (``__iter__`` -> DataLoader + DRInferenceCollator, same
    interleaving of ``batch_size`` blocks over the ranks, :99-115), and ``iter_batches()``, which hands whole
    ``[B, L]`` int32 slices of the memory map to ``Retriever`` (one memcpy into a pinned buffer per batch, no
    per-row Python objects) — the ingest path that can feed the encoder at tens of thousands of passages/s."""

    def _open(self):
        ids = np.load(self.data_files[0], mmap_mode="r")
        if ids.ndim != 2 or ids.dtype != np.int32:
            raise ValueError("%s: expected an int32 [n, L] array, got %s %s" % (self.data_files[0], ids.dtype, ids.shape))
        names_path = os.path.splitext(self.data_files[0])[0] + ".ids.txt"
        names = None
        if os.path.exists(names_path):
            with open(names_path) as f:
                names = f.read().split("\n")
            if len(names) < ids.shape[0]:
                raise ValueError("%s holds %d ids for %d rows" % (names_path, len(names), ids.shape[0]))
        return ids, names

    def _records(self):
        ids, names = self._open()
        for i in range(ids.shape[0]):
            yield {"id": names[i] if names else str(i), "row": ids[i]}

    def num_local_rows(self) -> int:
        """rows this rank will see (blocks ``process_index, process_index + W, ...`` of ``batch_size`` rows)"""
        n = np.load(self.data_files[0], mmap_mode="r").shape[0]
        bs, W, r = self.batch_size, self.num_processes, self.process_index
        full, rest = divmod(n, bs * W)
        return full * bs + max(0, min(bs, rest - r * bs))

    def iter_batches(self):
        """Yields ``(text_ids: list[str], ids: int32 [b, max_len] C-contiguous view or padded copy)`` for this rank's
        blocks, in the order ``__iter__`` would produce the same examples."""
        ids, names = self._open()
        n, width = ids.shape
        bs, W, r = self.batch_size, self.num_processes, self.process_index
        for b0 in range(r * bs, n, W * bs):
            b1 = min(n, b0 + bs)
            block = ids[b0:b1, : self.max_len]
            if width < self.max_len:  # stored narrower than the model's padded length: pad like the reference does
                padded = np.zeros((b1 - b0, self.max_len), dtype=np.int32)
                padded[:, :width] = block
                block = padded
            yield (names[b0:b1] if names else [str(i) for i in range(b0, b1)]), block

    def process_one(self, example):
        row = np.asarray(example["row"][: self.max_len], dtype=np.int64)
        if row.shape[0] < self.max_len:
            row = np.pad(row, (0, self.max_len - row.shape[0]))
        return {"text_id": get_idx(example), "input_ids": row.tolist(), "attention_mask": (row != 0).astype(np.int64).tolist(),
                "token_type_ids": [0] * self.max_len}
