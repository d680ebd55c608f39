"""GPU, end to end through the three CLI drivers exactly as the reference documents them
(docs/dr-msmarco-passage.md): train_dr -> build_index -> retrieve, on a tiny BERT checkpoint and a toy corpus
created on the fly (no network, no pretrained weights)."""
import json
import os
import pickle
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WORDS = ["the", "a", "of", "river", "bank", "money", "loan", "water", "fish", "tree", "green", "blue", "sky", "rain",
         "city", "road", "car", "train", "music", "piano", "guitar", "river", "stone", "bread", "cheese", "wine"]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from transformers import BertConfig, BertModel, BertTokenizer
    root = tmp_path_factory.mktemp("om_drivers")
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "title", "text", ":"] + sorted(set(WORDS))
    (root / "vocab.txt").write_text("\n".join(vocab))
    tok = BertTokenizer(str(root / "vocab.txt"), do_lower_case=True)
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=len(vocab), hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                     intermediate_size=256, max_position_embeddings=64)
    model_dir = root / "model"
    BertModel(cfg).save_pretrained(str(model_dir))
    tok.save_pretrained(str(model_dir))
    rng = np.random.default_rng(0)

    def sent(n):
        return " ".join(rng.choice(WORDS, n))

    with open(root / "corpus.tsv", "w") as f:
        for i in range(60):
            f.write(f"d{i}\t{sent(2)}\t{sent(12)}\n")
    with open(root / "queries.tsv", "w") as f:
        for i in range(7):
            f.write(f"q{i}\t{sent(4)}\n")
    with open(root / "train.jsonl", "w") as f:
        for i in range(32):
            enc = lambda s: tok.encode(s, add_special_tokens=False)  # noqa: E731
            f.write(json.dumps({"query": enc(sent(4)), "positives": [enc(sent(10))],
                                "negatives": [enc(sent(10)) for _ in range(5)]}) + "\n")
    return root


def _run(main, argv):
    old = sys.argv
    sys.argv = ["prog"] + [str(a) for a in argv]
    try:
        main()
    finally:
        sys.argv = old


def test_train_build_retrieve(workdir):
    from openmatch.driver import build_index, retrieve, train_dr
    from openmatch.utils import load_from_trec
    ckpt = workdir / "ckpt"
    _run(train_dr.main, ["--output_dir", ckpt, "--model_name_or_path", workdir / "model", "--do_train",
                         "--train_path", workdir / "train.jsonl", "--per_device_train_batch_size", 4,
                         "--train_n_passages", 4, "--learning_rate", "1e-3", "--q_max_len", 8, "--p_max_len", 16,
                         "--max_steps", 6, "--logging_steps", 2, "--save_steps", 1000, "--bf16",
                         "--dataloader_num_workers", 0])
    cfg = json.load(open(ckpt / "openmatch_config.json"))
    assert cfg["tied"] and cfg["pooling"] == "first" and not cfg["linear_head"]
    assert os.path.exists(ckpt / "config.json") and os.path.exists(ckpt / "tokenizer_config.json")

    emb = workdir / "emb"
    common = ["--output_dir", emb, "--model_name_or_path", ckpt, "--per_device_eval_batch_size", 16, "--q_max_len", 8,
              "--p_max_len", 32, "--dataloader_num_workers", 0]
    _run(build_index.main, common + ["--corpus_path", workdir / "corpus.tsv", "--doc_template", "<title> <text>",
                                     "--doc_column_names", "id,title,text", "--fp16"])
    with open(emb / "embeddings.corpus.rank.0", "rb") as f:
        enc, ids = pickle.load(f)
    assert enc.shape == (60, 128) and enc.dtype == np.float32 and ids[0] == "d0" and np.isfinite(enc).all()

    run = workdir / "run.trec"
    _run(retrieve.main, common + ["--query_path", workdir / "queries.tsv", "--query_template", "<text>",
                                  "--query_column_names", "id,text", "--trec_save_path", run, "--retrieve_depth", 10,
                                  "--use_gpu"])
    result = load_from_trec(str(run))
    assert len(result) == 7 and all(len(v) == 10 for v in result.values())
    # the TREC scores are exact inner products of the pickled embeddings
    with open(emb / "embeddings.query.rank.0", "rb") as f:
        qenc, qids = pickle.load(f)
    s = qenc @ enc.T
    for qi, qid in enumerate(qids):
        top = np.argsort(-s[qi], kind="stable")[:10]
        got = result[qid]
        assert list(got)[0] == ids[top[0]]
        assert abs(list(got.values())[0] - s[qi, top[0]]) <= 1e-3 * max(1.0, abs(s[qi, top[0]]))


def test_training_reduces_loss(workdir):
    """DRTrainer: a few AdamW steps on a repeated batch must drive the contrastive loss down (fused loss kernel
    gradients flow into the HF encoder)."""
    import types

    from openmatch.arguments import DataArguments, DRTrainingArguments, ModelArguments
    from openmatch.dataset import DRTrainDataset, QPCollator
    from openmatch.modeling import DRModel
    from openmatch.trainer import DRTrainer
    margs = ModelArguments(model_name_or_path=str(workdir / "model"))
    dargs = DataArguments(train_path=str(workdir / "train.jsonl"), train_n_passages=4, q_max_len=8, p_max_len=16)
    targs = DRTrainingArguments(output_dir=str(workdir / "ckpt2"), per_device_train_batch_size=8, learning_rate=2e-3,
                                max_steps=12, logging_steps=1, save_steps=0, warmup_ratio=0.0, dataloader_num_workers=0)
    model = DRModel.build(margs, dargs, targs)
    ds = DRTrainDataset(None, dargs)
    trainer = DRTrainer(model=model, args=targs, train_dataset=ds, data_collator=QPCollator(None, 8, 16))
    ds.trainer = trainer
    trainer.train()
    losses = [e["loss"] for e in trainer.state.log_history]
    assert len(losses) == 12 and np.isfinite(losses).all()
    assert np.mean(losses[-3:]) < np.mean(losses[:3]), losses


def test_grad_cache_matches_plain_step(workdir):
    """GCDenseTrainer (chunked, gradient cache) must produce the same parameter gradients as one plain step."""
    from openmatch.arguments import DataArguments, DRTrainingArguments, ModelArguments
    from openmatch.dataset import DRTrainDataset, QPCollator
    from openmatch.modeling import DRModel
    from openmatch.trainer import DRTrainer, GCDenseTrainer
    margs = ModelArguments(model_name_or_path=str(workdir / "model"))
    dargs = DataArguments(train_path=str(workdir / "train.jsonl"), train_n_passages=4, q_max_len=8, p_max_len=16)

    def grads(trainer_cls, **extra):
        targs = DRTrainingArguments(output_dir=str(workdir / "ckpt3"), per_device_train_batch_size=8, max_steps=1,
                                    dataloader_num_workers=0, **extra)
        torch.manual_seed(0)
        model = DRModel.build(margs, dargs, targs).cuda()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        trainer = trainer_cls(model=model, args=targs, train_dataset=DRTrainDataset(None, dargs),
                              data_collator=QPCollator(None, 8, 16))
        trainer._scaler = None
        batch = next(iter(trainer.get_train_dataloader()))
        loss = trainer.training_step(model, batch)
        g = model.lm_q.encoder.layer[1].output.dense.weight.grad.detach().float().cpu().numpy().copy()
        return float(loss), g

    l0, g0 = grads(DRTrainer)
    l1, g1 = grads(GCDenseTrainer, grad_cache=True, gc_q_chunk_size=3, gc_p_chunk_size=8)
    assert abs(l0 - l1) <= 1e-3 * max(1.0, abs(l0))
    assert np.linalg.norm(g0 - g1) <= 2e-2 * np.linalg.norm(g0)


def test_hard_negative_mining_loop_with_untied_encoders(workdir):
    """docs/dr-msmarco-passage.md:98-157 end to end: train (untied query / passage encoders) -> build_index -> retrieve
    over the TRAIN queries -> build_hn -> second training round on the mined shards.  Also checks the untied checkpoint
    layout (dense_retrieval_model.py:189-213,230-245) and that the two towers really differ after training."""
    from transformers import BertTokenizer

    from openmatch.arguments import ModelArguments
    from openmatch.driver import build_hn, build_index, retrieve, train_dr
    from openmatch.modeling import DRModelForInference
    from openmatch.utils import load_from_trec
    tok = BertTokenizer(str(workdir / "vocab.txt"), do_lower_case=True)
    ckpt = workdir / "ckpt_untied"
    _run(train_dr.main, ["--output_dir", ckpt, "--model_name_or_path", workdir / "model", "--do_train", "--untie_encoder",
                         "--train_path", workdir / "train.jsonl", "--per_device_train_batch_size", 4,
                         "--train_n_passages", 4, "--learning_rate", "1e-3", "--q_max_len", 8, "--p_max_len", 16,
                         "--max_steps", 4, "--logging_steps", 2, "--save_steps", 1000, "--bf16",
                         "--dataloader_num_workers", 0])
    cfg = json.load(open(ckpt / "openmatch_config.json"))
    assert cfg["tied"] is False
    assert os.path.exists(ckpt / "query_model" / "config.json") and os.path.exists(ckpt / "passage_model" / "config.json")
    model = DRModelForInference.build(ModelArguments(model_name_or_path=str(ckpt)))
    assert model.lm_q is not model.lm_p and not model.tied
    wq = model.lm_q.encoder.layer[0].attention.self.query.weight
    wp = model.lm_p.encoder.layer[0].attention.self.query.weight
    assert not torch.equal(wq, wp), "towers did not diverge"
    ids = torch.tensor([[2, 9, 10, 11, 3, 0, 0, 0]]).cuda()
    batch = {"input_ids": ids, "attention_mask": (ids != 0).long(), "token_type_ids": torch.zeros_like(ids)}
    model = model.cuda().eval()
    rq, rp = model(query=batch).q_reps, model(passage=batch).p_reps
    assert not torch.allclose(rq, rp, atol=1e-3), "query and passage towers give identical representations"
    del model

    emb = workdir / "emb_hn"
    common = ["--output_dir", emb, "--model_name_or_path", ckpt, "--per_device_eval_batch_size", 16, "--q_max_len", 8,
              "--p_max_len", 32, "--dataloader_num_workers", 0]
    _run(build_index.main, common + ["--corpus_path", workdir / "corpus.tsv", "--doc_template", "<title> <text>",
                                     "--doc_column_names", "id,title,text"])
    run = workdir / "train.trec"
    _run(retrieve.main, common + ["--query_path", workdir / "queries.tsv", "--query_template", "<text>",
                                  "--query_column_names", "id,text", "--trec_save_path", run, "--retrieve_depth", 30,
                                  "--use_gpu"])
    ranked = load_from_trec(str(run))
    # qrels: pretend the 3rd-ranked passage of every train query is its positive
    with open(workdir / "qrels.tsv", "w") as f:
        for q, docs in ranked.items():
            f.write("%s\t0\t%s\t1\n" % (q, list(docs)[2]))
    # pre-tokenised stores (text preprocessing is upstream of this tier's scope)
    def store(path_tsv, stem, text_cols):
        names, rows = [], []
        for line in open(path_tsv):
            parts = line.rstrip("\n").split("\t")
            names.append(parts[0])
            rows.append(tok.encode(" ".join(parts[c] for c in text_cols), add_special_tokens=False)[:32])
        arr = np.zeros((len(rows), 32), np.int32)
        for i, r in enumerate(rows):
            arr[i, : len(r)] = r
        np.save(workdir / (stem + ".npy"), arr)
        (workdir / (stem + ".ids.txt")).write_text("\n".join(names))
    store(workdir / "corpus.tsv", "collection_tok", (1, 2))
    store(workdir / "queries.tsv", "queries_tok", (1,))
    hn = workdir / "hn"
    _run(build_hn.main, ["--hn_file", run, "--qrels", workdir / "qrels.tsv", "--queries", workdir / "queries_tok.npy",
                         "--collection", workdir / "collection_tok.npy", "--save_to", hn, "--n_sample", 5, "--depth", 20,
                         "--seed", 3])
    rows = [json.loads(line) for line in open(hn / "split00.hn.jsonl")]
    assert len(rows) == len(ranked)
    col = np.load(workdir / "collection_tok.npy")
    col_ids = (workdir / "collection_tok.ids.txt").read_text().split("\n")
    doc_tok = {n: col[i][col[i] != 0].tolist() for i, n in enumerate(col_ids)}
    for (q, docs), r in zip(ranked.items(), rows):
        pos = list(docs)[2]
        assert r["positives"] == [doc_tok[pos]] and len(r["negatives"]) == 5
        allowed = [doc_tok[d] for d in [d for d in docs if d != pos][:20]]  # first `depth` non-relevant passages of the run
        assert all(n in allowed for n in r["negatives"]) and doc_tok[pos] not in r["negatives"]
    ckpt2 = workdir / "ckpt_round2"
    _run(train_dr.main, ["--output_dir", ckpt2, "--model_name_or_path", ckpt, "--do_train", "--train_dir", hn,
                         "--per_device_train_batch_size", 4, "--train_n_passages", 4, "--learning_rate", "1e-3",
                         "--q_max_len", 8, "--p_max_len", 16, "--max_steps", 3, "--logging_steps", 1, "--save_steps", 1000,
                         "--bf16", "--dataloader_num_workers", 0])
    assert json.load(open(ckpt2 / "openmatch_config.json"))["tied"] is False
