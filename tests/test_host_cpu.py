"""CPU: host-side logic of the drop-in package (no CUDA needed): argument parsing, TREC I/O, templates,
datasets / collators, DROutput, result merging, checkpoint formats, batch sharding."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from openmatch_b200 import utils
from openmatch_b200.arguments import DataArguments, DRTrainingArguments, InferenceArguments, ModelArguments


def test_drop_in_alias_modules():
    import openmatch
    from openmatch.loss import DistributedContrastiveLoss, SimpleContrastiveLoss  # noqa: F401
    from openmatch.modeling import DRModel, DRModelForInference, DROutput, LinearHead  # noqa: F401
    from openmatch.retriever import FaissRetriever, Retriever, SuccessiveRetriever
    from openmatch.driver import build_index, retrieve, successive_retrieve, train_dr
    assert FaissRetriever is Retriever and issubclass(SuccessiveRetriever, Retriever)
    for mod in (build_index, retrieve, successive_retrieve, train_dr):
        assert callable(mod.main)
    assert openmatch.modeling.DRModel is DRModel


def test_arguments_parse_like_reference():
    from transformers import HfArgumentParser
    parser = HfArgumentParser((ModelArguments, DataArguments, InferenceArguments))
    m, d, e = parser.parse_args_into_dataclasses(
        ["--model_name_or_path", "bert-base-uncased", "--output_dir", "/tmp/x", "--per_device_eval_batch_size", "256",
         "--fp16", "--use_gpu", "--q_max_len", "32", "--p_max_len", "128", "--pooling", "mean", "--normalize",
         "--doc_template", "<title> <text>", "--dataloader_num_workers", "1", "--trec_save_path", "/tmp/x/run.trec"])
    assert (m.pooling, m.normalize, d.p_max_len, e.per_device_eval_batch_size, e.fp16, e.use_gpu) == (
        "mean", True, 128, 256, True, True)
    assert e.world_size == 1 and e.process_index == 0 and e.local_process_index == 0
    parser = HfArgumentParser((ModelArguments, DataArguments, DRTrainingArguments))
    m, d, t = parser.parse_args_into_dataclasses(
        ["--model_name_or_path", "x", "--output_dir", "/tmp/y", "--train_n_passages", "8", "--negatives_x_device",
         "--per_device_train_batch_size", "64", "--bf16", "--learning_rate", "5e-6"])
    assert t.negatives_x_device and t.warmup_ratio == 0.1 and d.train_n_passages == 8 and t.bf16


def test_trec_roundtrip_and_merge(tmp_path):
    res = {"q1": {"d3": 1.5, "d1": 9.25, "d2": 1.5}, "q2": {"d9": -0.5}}
    path = str(tmp_path / "run.trec")
    utils.save_as_trec(res, path)
    lines = open(path).read().splitlines()
    assert lines[0] == "q1 Q0 d1 1 9.25 OpenMatch" and lines[1].startswith("q1 Q0 d3 2 1.5")
    back = utils.load_from_trec(path)
    assert back == {"q1": {"d1": 9.25, "d3": 1.5, "d2": 1.5}, "q2": {"d9": -0.5}}
    assert utils.load_from_trec(path, as_list=True, max_len_per_q=2)["q1"] == [("d1", 9.25), ("d3", 1.5)]
    import oracle
    r1 = {"q1": {"d1": 3.0, "d2": 1.0}}
    r2 = {"q1": {"d2": 9.0, "d4": 2.5}, "q3": {"d1": 1.0}}
    assert utils.merge_retrieval_results_by_score([r1, r2], 2) == oracle.merge_retrieval_results_by_score([r1, r2], 2)
    assert list(utils.merge_retrieval_results_by_score([r1, r2], 2)["q1"].items()) == [("d1", 3.0), ("d4", 2.5)]


def test_templates():
    assert utils.find_all_markers("Title: <title> Text: <text>") == ["title", "text"]
    assert utils.fill_template("<a.b>-<c>", {"a": {"b": 1}, "c": "x"}) == "1-x"
    with pytest.warns(RuntimeWarning):
        assert utils.fill_template("<title>|<text>", {"text": "t"}, allow_not_found=True) == "|t"
    with pytest.raises(ValueError):
        utils.fill_template("<nope>", {})


def test_mean_pooling_matches_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "misc.npz"))
    got = utils.mean_pooling(torch.from_numpy(z["mp_hidden"]), torch.from_numpy(z["mp_mask"]))
    np.testing.assert_allclose(got.numpy(), z["mp_out"], rtol=1e-6, atol=1e-7)


class _Tok:
    """whitespace 'tokenizer' with the HF call signature the datasets use"""

    def __call__(self, text, add_special_tokens=True, padding=False, truncation=True, max_length=8,
                 return_attention_mask=True, return_token_type_ids=True):
        ids = [101] + [1000 + len(w) for w in text.split()][: max_length - 2] + [102]
        mask = [1] * len(ids)
        if padding == "max_length":
            pad = max_length - len(ids)
            ids, mask = ids + [0] * pad, mask + [0] * pad
        return {"input_ids": ids, "attention_mask": mask, "token_type_ids": [0] * len(ids)}


def test_inference_datasets_and_collator(tmp_path):
    from openmatch_b200.dataset import DRInferenceCollator, InferenceDataset
    jl = tmp_path / "corpus.json"
    jl.write_text("\n".join(json.dumps({"id": i, "title": "t%d" % i, "text": "a bb ccc"}) for i in range(10)))
    tsv = tmp_path / "queries.tsv"
    tsv.write_text("\n".join("q%d\tsome query %d" % (i, i) for i in range(5)))
    da = DataArguments(corpus_path=str(jl), query_path=str(tsv), p_max_len=8, q_max_len=6)
    # two processes, batch 2: rank 0 sees docs 0,1,4,5,8,9 ; rank 1 sees 2,3,6,7 (reference's interleaving)
    seen = [[ex["text_id"] for ex in InferenceDataset.load(_Tok(), da, batch_size=2, num_processes=2, process_index=r)]
            for r in range(2)]
    assert seen == [["0", "1", "4", "5", "8", "9"], ["2", "3", "6", "7"]]
    ds = InferenceDataset.load(_Tok(), da, is_query=True, batch_size=4)
    items = list(ds)
    assert [it["text_id"] for it in items] == ["q0", "q1", "q2", "q3", "q4"] and len(items[0]["input_ids"]) == 6
    ids, batch = DRInferenceCollator()(items[:3])
    assert ids == ["q0", "q1", "q2"] and batch["input_ids"].shape == (3, 6) and batch["input_ids"].dtype == torch.int64
    # pre-tokenised memory-mapped format
    np.save(tmp_path / "c.npy", np.arange(40, dtype=np.int32).reshape(5, 8) % 7)
    da2 = DataArguments(corpus_path=str(tmp_path / "c.npy"), p_max_len=6)
    rows = list(InferenceDataset.load(None, da2, batch_size=5))
    assert len(rows) == 5 and len(rows[0]["input_ids"]) == 6 and rows[0]["attention_mask"][0] == 0
    with pytest.raises(ValueError):
        InferenceDataset.load(None, DataArguments(corpus_path="x.parquet"))


def test_train_dataset_and_qp_collator(tmp_path):
    from openmatch_b200.dataset import DRTrainDataset, QPCollator
    path = tmp_path / "train.jsonl"
    recs = [{"query": [5, 6, 7], "positives": [[1, 2], [3]], "negatives": [[9, 9, 9, 9, 9, 9], [8], [7, 7]]} for _ in range(6)]
    path.write_text("\n".join(json.dumps(r) for r in recs))
    da = DataArguments(train_path=str(path), train_n_passages=4, q_max_len=4, p_max_len=5)
    ds = DRTrainDataset(None, da)
    assert len(ds) == 6
    ex = next(iter(ds))
    assert len(ex["passages"]) == 4 and ex["passages"][0]["input_ids"] == [1, 2] and ex["query"]["input_ids"] == [5, 6, 7]
    q, p = QPCollator(None, max_q_len=4, max_p_len=5)([ex, ex])
    assert q["input_ids"].shape == (2, 4) and p["input_ids"].shape == (8, 5)
    assert p["attention_mask"][1].tolist() == [1, 1, 1, 1, 1]  # truncated to p_max_len


def test_batch_sharding_matches_reference_layout():
    from openmatch_b200.trainer.dense_trainer import _ShardByBatch
    got = [list(_ShardByBatch(range(10), per_device=2, world=2, rank=r)) for r in range(2)]
    assert got[0] == [0, 1, 4, 5, 8, 9] and got[1][:4] == [2, 3, 6, 7] and len(got[0]) == len(got[1])


def test_droutput_and_linear_head_checkpoint(tmp_path):
    from openmatch_b200.modeling import DROutput, LinearHead
    out = DROutput(q_reps=torch.ones(2, 3), p_reps=None)
    assert out["q_reps"] is out.q_reps and out.keys() == ["q_reps"] and out[0] is out.q_reps
    head = LinearHead(8, 4)
    head.save(str(tmp_path))
    assert json.load(open(tmp_path / "head_config.json")) == {"input_dim": 8, "output_dim": 4}
    again = LinearHead.load(str(tmp_path))
    assert torch.equal(again.linear.weight, head.linear.weight) and again.linear.bias is None


def test_embedding_pickle_format_is_the_references(tmp_path):
    # (float32 [n, d] C-order, list[str]) at protocol 4 under embeddings.corpus.rank.{r}
    enc = np.arange(12, dtype=np.float32).reshape(3, 4)
    path = tmp_path / "embeddings.corpus.rank.0"
    with open(path, "wb") as f:
        pickle.dump((enc, ["a", "b", "c"]), f, protocol=4)
    with open(path, "rb") as f:
        got, ids = pickle.load(f)
    assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"] and ids == ["a", "b", "c"]


def test_model_refuses_cpu_tensors():
    """The product has no CPU path: encoding CPU tensors must fail loudly, not fall back."""
    from openmatch_b200.modeling import DRModelForInference

    class Dummy(torch.nn.Module):
        pass

    m = DRModelForInference(lm_q=Dummy(), lm_p=Dummy())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.encode_passage({"input_ids": torch.zeros(1, 4, dtype=torch.long), "attention_mask": torch.ones(1, 4, dtype=torch.long)})


def test_rank_arrays_and_vectorised_trec_writer_match_the_dict_path(tmp_path):
    # Retriever.search(as_arrays=True) + RankArrays.save_trec must write byte-for-byte what the reference's
    # dict-of-dicts + save_as_trec writes (dense_retriever.py:183-188, utils.py:126-136), padding included
    import types

    import oracle
    from openmatch_b200.retriever.dense_retriever import RankArrays, Retriever, _results_dict
    rng = np.random.default_rng(5)
    x = rng.standard_normal((40, 16)).astype(np.float32)
    q = rng.standard_normal((7, 16)).astype(np.float32)
    idx = oracle.FlatIPIndex(16)
    idx.add(x)
    r = Retriever.__new__(Retriever)
    r.index = idx
    r.args = types.SimpleNamespace(world_size=1)
    r.doc_lookup = ["doc%03d" % i for i in range(40)]
    r.query_lookup = ["q%d" % i for i in range(7)]
    r._load_queries = lambda: q
    for topk in (5, 40, 64):  # 64 > ntotal: -1 padded tail
        as_dict = r.search(topk)
        arrays = r.search(topk, as_arrays=True)
        assert isinstance(arrays, RankArrays) and arrays.to_dict() == as_dict
        assert all(len(v) == min(topk, 40) for v in as_dict.values())
        a, b = tmp_path / ("dict.%d.trec" % topk), tmp_path / ("arr.%d.trec" % topk)
        utils.save_as_trec(as_dict, str(a))
        arrays.save_trec(str(b))
        assert a.read_bytes() == b.read_bytes()
    # duplicated doc-id strings: the dict merges them, so the writer must fall back to the dict path
    dup = RankArrays(["q0"], np.array(["a", "b", "a"]), np.array([[3.0, 2.0, 1.0]], np.float32), np.array([[0, 1, 2]]))
    assert not dup.unique_doc_ids()
    dup.save_trec(str(tmp_path / "dup.trec"))
    utils.save_as_trec(_results_dict(["q0"], np.array(["a", "b", "a"]), dup.D, dup.I), str(tmp_path / "dup2.trec"))
    assert (tmp_path / "dup.trec").read_bytes() == (tmp_path / "dup2.trec").read_bytes()


def test_pretokenized_block_iterator_matches_the_per_example_iterator(tmp_path):
    # PretokenizedDataset.iter_batches (whole [B, L] int32 slices, the ingest path of Retriever) must hand out exactly the
    # examples, ids and rank interleaving of the reference-style per-example iterator (inference_dataset.py:99-115)
    from openmatch_b200.dataset import InferenceDataset
    rng = np.random.default_rng(0)
    n, width = 53, 20
    ids = rng.integers(1, 1000, (n, width)).astype(np.int32)
    for r in range(n):
        ids[r, rng.integers(3, width):] = 0
    path = tmp_path / "corpus.npy"
    np.save(path, ids)
    (tmp_path / "corpus.ids.txt").write_text("\n".join("doc%d" % i for i in range(n)))
    for p_max_len in (16, 20, 24):  # narrower than / equal to / wider than the stored width
        d = DataArguments(corpus_path=str(path), p_max_len=p_max_len)
        total = 0
        for W in (1, 3):
            for r in range(W):
                ds = InferenceDataset.load(None, d, is_query=False, batch_size=8, num_processes=W, process_index=r)
                per_row = list(ds)
                blocks = list(ds.iter_batches())
                names = [x for b in blocks for x in b[0]]
                rows = np.concatenate([b[1] for b in blocks]) if blocks else np.zeros((0, p_max_len), np.int32)
                assert names == [e["text_id"] for e in per_row]
                assert rows.dtype == np.int32 and rows.shape == (len(per_row), p_max_len)
                assert (rows == np.array([e["input_ids"] for e in per_row]).reshape(len(per_row), p_max_len)).all()
                assert ds.num_local_rows() == len(per_row)
                assert all(b[1].shape[0] <= 8 for b in blocks)
                total += len(per_row) if W == 3 else 0
        assert total == n
