"""Host-side helpers with the reference's names and semantics (``src/openmatch/utils.py``): pooling
reference implementation for the torch training path, result merging, TREC I/O and text templates."""
from __future__ import annotations

import warnings
from typing import Dict, List

import torch


def mean_pooling(token_embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Masked mean over tokens in fp32, denominator clamped at 1e-9 (utils.py:233-235).  Used by the torch
    (training) path; inference pools inside the CUDA encoder."""
    weights = attention_mask.unsqueeze(-1).expand(token_embeddings.size()).float()
    return (token_embeddings * weights).sum(1) / weights.sum(1).clamp(min=1e-9)


def merge_retrieval_results_by_score(results: List[Dict[str, Dict[str, float]]], topk: int = 100):
    """Union of per-partition results: the first score seen for a (query, doc) wins, then a stable
    descending sort keeps ``topk`` docs per query (utils.py:215-229)."""
    pooled: Dict[str, Dict[str, float]] = {}
    for part in results:
        for qid, docs in part.items():
            bucket = pooled.setdefault(qid, {})
            for did, score in docs.items():
                bucket.setdefault(did, score)
    return {qid: dict(sorted(docs.items(), key=lambda kv: kv[1], reverse=True)[:topk]) for qid, docs in pooled.items()}


def save_as_trec(rank_result: Dict[str, Dict[str, float]], output_path: str, run_id: str = "OpenMatch"):
    """``<qid> Q0 <docid> <rank> <score> <run_id>`` lines, docs ordered by descending score (utils.py:126-136)."""
    with open(output_path, "w") as out:
        for qid, docs in rank_result.items():
            ranked = sorted(docs.items(), key=lambda kv: kv[1], reverse=True)
            out.writelines("{} Q0 {} {} {} {}\n".format(qid, did, r + 1, s, run_id) for r, (did, s) in enumerate(ranked))


def save_trec_arrays(query_ids, doc_names, I, D, output_path: str, run_id: str = "OpenMatch"):
    """Array form of :func:`save_as_trec` for search output that is still in (I, D) form: ``I`` int64 [nq, k] rows
    into ``doc_names`` (-1 = padding), ``D`` float32 [nq, k] sorted descending per row.  Writes byte-for-byte what
    ``save_as_trec`` writes for the equivalent ``{qid: {docid: score}}`` dict (scores print as ``float(np.float32)``,
    utils.py:126-136 / dense_retriever.py:183-188) without building the 7 M-entry dict of a top-1000 MS MARCO run.
    Precondition (checked by the caller): doc ids are unique, so the dict would not have merged any entries."""
    import numpy as np
    names = np.asarray(doc_names)
    I = np.asarray(I)
    scores = np.asarray(D, dtype=np.float64)
    with open(output_path, "w") as out:
        for qi, qid in enumerate(query_ids):
            row = I[qi]
            n = int((row >= 0).sum()) if (row < 0).any() else row.shape[0]  # padding is a suffix
            docs = names[row[:n]].tolist()
            vals = scores[qi, :n].tolist()
            qid = str(qid)
            out.write("".join("{} Q0 {} {} {} {}\n".format(qid, d, r + 1, v, run_id)
                              for r, (d, v) in enumerate(zip(docs, vals))))


def load_from_trec(input_path: str, as_list: bool = False, max_len_per_q: int = None):
    """Reads 6-column TREC runs or 3-column ``qid docid score`` files (utils.py:139-169)."""
    result, seen = {}, 0
    with open(input_path) as src:
        for line in src:
            cols = line.split()
            if len(cols) == 6:
                qid, _, did, _, score, _ = cols
            elif len(cols) == 3:
                qid, did, score = cols
            else:
                raise ValueError("Invalid run format")
            if qid not in result:
                result[qid] = [] if as_list else {}
                seen = 0
            if max_len_per_q is None or seen < max_len_per_q:
                if as_list:
                    result[qid].append((did, float(score)))
                else:
                    result[qid][did] = float(score)
            seen += 1
    return result


def find_all_markers(template: str) -> List[str]:
    """Names between '<' and '>' in a template, left to right (utils.py:172-187)."""
    names, pos = [], 0
    while True:
        lo = template.find("<", pos)
        hi = template.find(">", lo) if lo != -1 else -1
        if lo == -1 or hi == -1:
            return names
        names.append(template[lo + 1:hi])
        pos = hi + 1


def fill_template(template: str, data: Dict, markers: List[str] = None, allow_not_found: bool = False) -> str:
    """Substitutes ``<a.b>`` markers with ``data['a']['b']`` (utils.py:190-212)."""
    for marker in (markers if markers is not None else find_all_markers(template)):
        value = data
        for key in marker.split("."):
            value = value.get(key, None) if isinstance(value, dict) else None
            if value is None:
                break
        if value is None:
            if not allow_not_found:
                raise ValueError("Cannot find the marker '{}' in the data".format(marker))
            warnings.warn("Marker '{}' not found in data. Replacing it with an empty string.".format(marker),
                          RuntimeWarning)
            value = ""
        template = template.replace("<{}>".format(marker), str(value))
    return template
