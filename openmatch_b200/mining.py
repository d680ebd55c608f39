"""Hard-negative mining: the step after retrieval in the reference's documented workflow
(``docs/dr-msmarco-passage.md:98-157``): retrieve over the TRAIN queries with the current model, then turn the run into
training examples whose negatives are the top-ranked non-relevant passages (``scripts/msmarco/build_hn.py:13-37`` —
``load_ranking`` — and ``:73-93`` — sharded ``splitNN.hn.jsonl`` output in the ``{"query", "positives", "negatives"}``
token-id format that ``DRTrainDataset`` reads).

Two entry points with the same selection rule (negatives = ranked passages that are not relevant, in rank order; keep the
first ``depth``; shuffle; keep ``n_sample``):
  * :func:`load_ranking` walks a TREC run file exactly like the reference's generator (same grouping by consecutive query
    id, same use of the ``random`` module — seed it to reproduce the reference's draw);
  * :func:`negatives_from_run` works on the arrays ``Retriever.search(as_arrays=True)`` returns, so the mining loop needs
    no TREC round trip.
The reference tokenises query / passage text here; in this tier's scope text preprocessing is upstream, so the writer
takes PRE-TOKENISED stores (``{id: [token ids]}`` or the ``.npy`` + ``.ids.txt`` pair of ``PretokenizedDataset``).
"""
from __future__ import annotations

import json
import os
import random
from typing import Dict, Iterable, Iterator, List, Mapping, Sequence, Tuple

import numpy as np


def read_qrel(relevance_file: str) -> Dict[str, List[str]]:
    """``SimpleTrainPreProcessor.read_qrel`` (src/openmatch/utils.py:48-59): tab separated ``qid, _, docid, rel``"""
    qrel: Dict[str, List[str]] = {}
    with open(relevance_file, encoding="utf8") as f:
        for line in f:
            if not line.strip():
                continue
            topicid, _, docid, rel = line.rstrip("\n").split("\t")
            assert rel == "1"
            qrel.setdefault(topicid, []).append(docid)
    return qrel


def _finish(negatives: List[str], n_sample: int, depth: int, rng) -> List[str]:
    negatives = negatives[:depth]
    rng.shuffle(negatives)
    return negatives[:n_sample]


def load_ranking(rank_file: str, relevance: Mapping[str, Sequence[str]], n_sample: int, depth: int, rng=random
                 ) -> Iterator[Tuple[str, Sequence[str], List[str]]]:
    """Yields ``(qid, relevant ids, sampled hard negatives)`` per run of consecutive lines with the same query id
    (``build_hn.py:13-37``).  ``rng``: the ``random`` module (reference behaviour) or a ``random.Random``."""
    curr_q, negatives = None, []
    with open(rank_file) as rf:
        for line in rf:
            parts = line.strip().split()
            if len(parts) != 6:
                raise ValueError("not a TREC run line: %r" % line)
            q, p = parts[0], parts[2]
            if q != curr_q:
                if curr_q is not None:
                    yield curr_q, relevance[curr_q], _finish(negatives, n_sample, depth, rng)
                curr_q, negatives = q, []
            if p not in relevance[q]:
                negatives.append(p)
    if curr_q is not None:
        yield curr_q, relevance[curr_q], _finish(negatives, n_sample, depth, rng)


def negatives_from_run(query_ids: Sequence[str], doc_names: np.ndarray, I: np.ndarray,
                       relevance: Mapping[str, Sequence[str]], n_sample: int = 30, depth: int = 200, rng=random
                       ) -> Iterator[Tuple[str, Sequence[str], List[str]]]:
    """Same rule on search output arrays: row ``i`` of ``I`` ranks rows of ``doc_names`` for ``query_ids[i]`` (-1 = pad)."""
    doc_names = np.asarray(doc_names)
    for qi, qid in enumerate(query_ids):
        row = I[qi]
        ranked = doc_names[row[row >= 0]].tolist()
        rel = set(relevance[qid])
        yield qid, relevance[qid], _finish([p for p in ranked if p not in rel], n_sample, depth, rng)


class TokenStore:
    """id -> token ids, backed by a dict or by the ``<name>.npy`` (int32 [n, L], 0 = pad) + ``<name>.ids.txt`` pair"""

    def __init__(self, source):
        if isinstance(source, Mapping):
            self._map, self._rows = source, None
        else:
            self._rows = np.load(source, mmap_mode="r")
            names_path = os.path.splitext(source)[0] + ".ids.txt"
            names = open(names_path).read().split("\n") if os.path.exists(names_path) else [str(i) for i in range(self._rows.shape[0])]
            self._map = {n: i for i, n in enumerate(names[: self._rows.shape[0]])}

    def __getitem__(self, key: str) -> List[int]:
        v = self._map[key]
        if self._rows is None:
            return list(v)
        row = np.asarray(self._rows[v])
        return row[row != 0].tolist()


def write_hn_shards(examples: Iterable[Tuple[str, Sequence[str], Sequence[str]]], queries, passages, save_to: str,
                    shard_size: int = 45000, truncate: int = 128, query_max_len: int = 32) -> List[str]:
    """``build_hn.py:73-93``: one JSON line per query, ``shard_size`` lines per ``splitNN.hn.jsonl``."""
    queries, passages = TokenStore(queries), TokenStore(passages)
    os.makedirs(save_to, exist_ok=True)
    paths, f, counter, shard = [], None, 0, 0
    for qid, pos, neg in examples:
        if f is None:
            paths.append(os.path.join(save_to, "split%02d.hn.jsonl" % shard))
            f = open(paths[-1], "w")
        f.write(json.dumps({"query": queries[qid][:query_max_len],
                            "positives": [passages[p][:truncate] for p in pos],
                            "negatives": [passages[n][:truncate] for n in neg]}) + "\n")
        counter += 1
        if counter == shard_size:
            f.close()
            f, counter, shard = None, 0, shard + 1
    if f is not None:
        f.close()
    return paths
