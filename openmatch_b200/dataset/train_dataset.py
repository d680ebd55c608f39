"""Training examples for contrastive DR (reference: ``src/openmatch/dataset/train_dataset.py:48-119``):
JSON lines ``{"query": [ids], "positives": [[ids], ...], "negatives": [[ids], ...]}`` of pre-tokenised
text; one positive and ``train_n_passages - 1`` negatives are drawn per query and epoch."""
from __future__ import annotations

import glob
import json
import os
import random
from typing import List

from torch.utils.data import IterableDataset

from ..arguments import DataArguments


class DRTrainDataset(IterableDataset):
    def __init__(self, tokenizer, data_args: DataArguments, trainer=None, shuffle_seed: int = None, cache_dir: str = None):
        super().__init__()
        self.tokenizer, self.data_args, self.trainer = tokenizer, data_args, trainer
        self.data_files = [data_args.train_path] if data_args.train_dir is None else sorted(
            glob.glob(os.path.join(data_args.train_dir, "*.jsonl")))
        self.shuffle_seed = shuffle_seed
        self.neg_num = data_args.train_n_passages - 1

    def __len__(self):
        return sum(1 for path in self.data_files for _ in open(path))

    def create_one_example(self, token_ids: List[int], is_query: bool = False):
        max_len = self.data_args.q_max_len if is_query else self.data_args.p_max_len
        if self.tokenizer is not None:
            # the reference calls tokenizer.encode_plus(ids, truncation='only_first', max_length=...) (:62-70);
            # that API no longer exists in transformers 5.x, so special tokens are added explicitly
            if not hasattr(self, "_affix"):
                empty = list(self.tokenizer("", add_special_tokens=True)["input_ids"])  # e.g. [CLS, SEP] / [</s>]
                k = 1 if len(empty) >= 2 else 0
                self._affix = (empty[:k], empty[k:])
            prefix, suffix = self._affix
            room = max(max_len - len(prefix) - len(suffix), 0)
            return {"input_ids": prefix + list(token_ids)[:room] + suffix}
        return {"input_ids": list(token_ids)[:max_len]}

    def _pick(self, example, epoch: int, hashed_seed):
        positives, negatives = example["positives"], example["negatives"]
        if self.data_args.positive_passage_no_shuffle or hashed_seed is None:
            chosen = [positives[0]]
        else:
            chosen = [positives[(hashed_seed + epoch) % len(positives)]]
        want = self.neg_num
        if len(negatives) < want:
            negs = random.choices(negatives, k=want) if hashed_seed is not None else (list(negatives) * 2)[:want]
        elif self.data_args.train_n_passages == 1:
            negs = []
        elif self.data_args.negative_passage_no_shuffle:
            negs = negatives[:want]
        else:
            offset = epoch * want % len(negatives)
            pool = list(negatives)
            if hashed_seed is not None:
                random.Random(hashed_seed).shuffle(pool)
            negs = (pool * 2)[offset: offset + want]
        return chosen + list(negs)

    def __iter__(self):
        epoch = int(self.trainer.state.epoch) if self.trainer is not None else 0
        hashed_seed = hash(self.trainer.args.seed) if self.trainer is not None else self.shuffle_seed
        records = (json.loads(line) for path in self.data_files for line in open(path) if line.strip())
        if self.shuffle_seed is not None:  # buffered shuffle, like datasets' streaming shuffle
            rng, buf, source = random.Random(self.shuffle_seed + epoch), [], records

            def shuffled():
                for rec in source:
                    buf.append(rec)
                    if len(buf) >= 10_000:
                        yield buf.pop(rng.randrange(len(buf)))
                while buf:
                    yield buf.pop(rng.randrange(len(buf)))
            records = shuffled()
        for example in records:
            passages = self._pick(example, epoch, hashed_seed)
            assert len(passages) == self.data_args.train_n_passages
            yield {"query": self.create_one_example(example["query"], is_query=True),
                   "passages": [self.create_one_example(p) for p in passages]}
