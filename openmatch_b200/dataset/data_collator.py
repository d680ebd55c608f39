"""Batch collation for the hot path's inputs (reference: ``src/openmatch/dataset/data_collator.py``).
Both collators emit exactly the int64 ``[B, L]`` tensors the encoder consumes, always padded to the
configured maximum length like the reference does."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List

import torch


def _stack(features: List[Dict[str, Any]], keys) -> Dict[str, torch.Tensor]:
    return {k: torch.tensor([f[k] for f in features], dtype=torch.long) for k in keys}


@dataclass
class DRInferenceCollator:
    """``(text_ids, {input_ids, attention_mask, token_type_ids})`` (data_collator.py:78-83)."""

    def __call__(self, features):
        text_ids = [f["text_id"] for f in features]
        keys = [k for k in features[0].keys() if k != "text_id" and features[0][k] is not None]
        return text_ids, _stack(features, keys)


@dataclass
class QPCollator:
    """``[{query, passages}] -> (query batch, passage batch)`` padded to ``max_q_len`` / ``max_p_len``
    (data_collator.py:8-40)."""
    tokenizer: Any = None
    max_q_len: int = 32
    max_p_len: int = 128

    def _pad(self, items, max_len):
        if self.tokenizer is not None:
            return dict(self.tokenizer.pad(items, padding="max_length", max_length=max_len, return_tensors="pt"))
        ids = torch.zeros((len(items), max_len), dtype=torch.long)
        mask = torch.zeros((len(items), max_len), dtype=torch.long)
        for i, it in enumerate(items):
            seq = list(it["input_ids"])[:max_len]
            ids[i, : len(seq)] = torch.tensor(seq, dtype=torch.long)
            mask[i, : len(seq)] = 1
        return {"input_ids": ids, "attention_mask": mask}

    def __call__(self, features):
        queries = [f["query"] for f in features]
        passages = [f["passages"] for f in features]
        if isinstance(queries[0], list):
            queries = [q for group in queries for q in group]
        if isinstance(passages[0], list):
            passages = [p for group in passages for p in group]
        return self._pad(queries, self.max_q_len), self._pad(passages, self.max_p_len)
