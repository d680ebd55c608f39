"""GPU parity: CUDA encoder (C ABI, bf16 tensor-core compute) vs the fp32 CPU oracle (oracle/encoder.py)
and vs the golden vectors the reference's own DRModelForInference produced (tests/golden/).

Tolerance (stated in SURVEY.md section 8c, anchored on the reference's own bf16-autocast-vs-fp32 drift of
rel-L2 5.3e-3 / cosine >= 0.99998): rel-L2 <= 1e-2 and per-row cosine >= 0.9999 on reps and on
last_hidden_state of attended tokens."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.encoder import EncoderSpec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc_mod():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from openmatch_b200 import encoder
    return encoder


def _check(got, want, what, rel_tol=1e-2, cos_tol=0.9999):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    got2, want2 = got.reshape(-1, got.shape[-1]), want.reshape(-1, want.shape[-1])
    rel = np.linalg.norm(got2 - want2) / max(np.linalg.norm(want2), 1e-30)
    cos = (got2 * want2).sum(1) / np.maximum(np.linalg.norm(got2, axis=1) * np.linalg.norm(want2, axis=1), 1e-30)
    assert np.isfinite(got).all(), what + ": non-finite output"
    assert rel <= rel_tol, "%s: rel-L2 %.3e > %.1e" % (what, rel, rel_tol)
    assert cos.min() >= cos_tol, "%s: min cosine %.6f < %.4f" % (what, cos.min(), cos_tol)
    print("[parity] %-28s rel-L2 %.2e (tol %.0e)  min cosine %.6f" % (what, rel, rel_tol, cos.min()))
    return rel, cos.min()


def _golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, sd


def test_bert_small_matches_reference_golden(enc_mod, golden_dir):
    z, sd = _golden(golden_dir, "bert_small.npz")
    spec = dict(arch="bert", layers=2, hidden=128, heads=2, ffn=512, vocab=512, max_pos=128, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="first", max_batch_tokens=1024)
    assert any(n.startswith("pooler") for n in enc.ignored)
    ids, mask, tt = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask", "token_type_ids"))
    hidden, reps = enc.encode(ids, mask, tt, return_hidden=True)
    m = z["attention_mask"].astype(bool)
    _check(reps.cpu().numpy(), z["reps"], "reps vs reference")
    _check(hidden.cpu().numpy()[m], z["hidden"][m], "hidden vs reference")


def test_t5_small_matches_reference_golden(enc_mod, golden_dir):
    z, sd = _golden(golden_dir, "t5_small.npz")
    spec = dict(arch="t5", layers=2, hidden=128, heads=2, ffn=512, vocab=512, ln_eps=1e-6, rel_buckets=32,
                rel_max_distance=128)
    enc = enc_mod.CudaEncoder(spec, sd, head_weight=torch.from_numpy(z["head_weight"]), pooling="mean", normalize=True,
                              max_batch_tokens=1024)
    ids, mask = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask"))
    hidden, reps = enc.encode(ids, mask, return_hidden=True)
    m = z["attention_mask"].astype(bool)
    _check(hidden.cpu().numpy()[m], z["hidden"][m], "hidden vs reference")
    _check(reps.cpu().numpy(), z["reps"], "reps vs reference")
    assert np.abs(reps.cpu().numpy() - z["reps"]).max() <= 2e-3  # normalised reps: max-abs bound


def _rand_bert_sd(gen, layers, H, F, vocab, max_pos, std=0.02):
    def w(*shape):
        return torch.randn(*shape, generator=gen) * std

    def ln():
        return 1.0 + 0.1 * torch.randn(H, generator=gen), 0.05 * torch.randn(H, generator=gen)

    sd = {"embeddings.word_embeddings.weight": w(vocab, H), "embeddings.position_embeddings.weight": w(max_pos, H),
          "embeddings.token_type_embeddings.weight": w(2, H)}
    sd["embeddings.LayerNorm.weight"], sd["embeddings.LayerNorm.bias"] = ln()
    for i in range(layers):
        p = f"encoder.layer.{i}."
        for n, (o, k) in {"attention.self.query": (H, H), "attention.self.key": (H, H), "attention.self.value": (H, H),
                          "attention.output.dense": (H, H), "intermediate.dense": (F, H),
                          "output.dense": (H, F)}.items():
            sd[p + n + ".weight"], sd[p + n + ".bias"] = w(o, k), w(o)
        sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"] = ln()
        sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"] = ln()
    return sd


def _rand_t5_sd(gen, layers, H, heads, F, vocab):
    I = heads * 64

    def w(o, k, std):
        return torch.randn(o, k, generator=gen) * std

    sd = {"shared.weight": torch.randn(vocab, H, generator=gen),
          "encoder.final_layer_norm.weight": 1.0 + 0.1 * torch.randn(H, generator=gen),
          "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": torch.randn(32, heads, generator=gen)}
    for i in range(layers):
        p = f"encoder.block.{i}.layer."
        sd[p + "0.SelfAttention.q.weight"] = w(I, H, (H * 64) ** -0.5)
        sd[p + "0.SelfAttention.k.weight"] = w(I, H, H ** -0.5)
        sd[p + "0.SelfAttention.v.weight"] = w(I, H, H ** -0.5)
        sd[p + "0.SelfAttention.o.weight"] = w(H, I, I ** -0.5)
        sd[p + "0.layer_norm.weight"] = 1.0 + 0.1 * torch.randn(H, generator=gen)
        sd[p + "1.DenseReluDense.wi.weight"] = w(F, H, H ** -0.5)
        sd[p + "1.DenseReluDense.wo.weight"] = w(H, F, F ** -0.5)
        sd[p + "1.layer_norm.weight"] = 1.0 + 0.1 * torch.randn(H, generator=gen)
    return sd


def _ids(gen, B, L, vocab, ragged=True):
    ids = torch.randint(5, vocab, (B, L), generator=gen)
    mask = torch.ones(B, L, dtype=torch.long)
    if ragged:
        lens = torch.randint(min(2, L), L + 1, (B,), generator=gen)
        lens[0] = L
        for b in range(B):
            mask[b, lens[b]:] = 0
            ids[b, lens[b]:] = 0
    return ids, mask


@pytest.mark.parametrize("L,B", [(128, 6), (32, 9), (100, 3), (17, 11), (64, 4), (1, 5)])
def test_bert_base_lengths_vs_oracle(enc_mod, L, B):
    gen = torch.Generator().manual_seed(100 + L)
    layers, H, F, vocab = 3, 768, 3072, 2000
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 128)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=12, ffn=F, vocab=vocab, max_pos=128, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="first", max_batch_tokens=2048)
    ids, mask = _ids(gen, B, L, vocab)
    tt = torch.randint(0, 2, (B, L), generator=gen)
    hidden, reps = enc.encode(ids.cuda(), mask.cuda(), tt.cuda(), return_hidden=True)
    ospec = EncoderSpec("bert", layers, H, 12, F, 1e-12, pooling="first")
    oh, oreps = oracle.encode_reps(sd, ospec, ids, mask, tt)
    m = mask.numpy().astype(bool)
    _check(reps.cpu().numpy(), oreps.numpy(), "reps")
    _check(hidden.cpu().numpy()[m], oh.numpy()[m], "hidden")


def test_bert_base_full_depth_mean_pool_bf16_out(enc_mod):
    gen = torch.Generator().manual_seed(7)
    layers, H, F, vocab = 12, 768, 3072, 3000
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 128)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=12, ffn=F, vocab=vocab, max_pos=128, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="mean", normalize=True, max_batch_tokens=4096)
    ids, mask = _ids(gen, 8, 128, vocab)
    reps = enc.encode(ids.cuda(), mask.cuda())
    ospec = EncoderSpec("bert", layers, H, 12, F, 1e-12, pooling="mean", normalize=True)
    _, oreps = oracle.encode_reps(sd, ospec, ids, mask)
    _check(reps.cpu().numpy(), oreps.numpy(), "reps (12 layers)")
    # strided bf16 output straight into a wider buffer (index-shard style)
    buf = torch.zeros(8, 1024, dtype=torch.bfloat16, device="cuda")
    enc.encode(ids.cuda(), mask.cuda(), out=buf[:, :768])
    _check(buf[:, :768].float().cpu().numpy(), oreps.numpy(), "bf16 reps", rel_tol=1.2e-2)
    assert (buf[:, 768:] == 0).all()


def test_bert_large_shape(enc_mod):
    gen = torch.Generator().manual_seed(8)
    layers, H, F, vocab = 2, 1024, 4096, 1500
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 128)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=16, ffn=F, vocab=vocab, max_pos=128, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="first", max_batch_tokens=1024)
    ids, mask = _ids(gen, 5, 128, vocab)
    reps = enc.encode(ids.cuda(), mask.cuda())
    _, oreps = oracle.encode_reps(sd, EncoderSpec("bert", layers, H, 16, F, 1e-12), ids, mask)
    _check(reps.cpu().numpy(), oreps.numpy(), "reps (bert-large width)")


def test_bert_large_full_depth_vs_oracle(enc_mod):
    # BASELINE.json configs[4]'s encoder: all 24 layers of bert-large (1024 hidden, 16 heads, 4096 ffn), ragged batch
    gen = torch.Generator().manual_seed(24)
    layers, H, F, vocab = 24, 1024, 4096, 1200
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 128)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=16, ffn=F, vocab=vocab, max_pos=128, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="first", max_batch_tokens=1024)
    ids, mask = _ids(gen, 6, 48, vocab)
    reps = enc.encode(ids.cuda(), mask.cuda())
    _, oreps = oracle.encode_reps(sd, EncoderSpec("bert", layers, H, 16, F, 1e-12), ids, mask)
    _check(reps.cpu().numpy(), oreps.numpy(), "reps (bert-large, 24 layers)")


def test_bench_sized_batch_vs_oracle(enc_mod):
    # the batch geometry bench.py runs — B = 256 x L = 128 = 32 768 tokens, 256 m-tiles, every GEMM several waves deep,
    # the static tile schedule and the residual epilogue's next-tile prefetch all exercised — on 2 layers (the CPU oracle
    # finishes in seconds); every 7th sequence is checked
    gen = torch.Generator().manual_seed(256)
    layers, H, F, vocab = 2, 768, 3072, 2500
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 128)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=12, ffn=F, vocab=vocab, max_pos=128, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="first", max_batch_tokens=256 * 128)
    ids, mask = _ids(gen, 256, 128, vocab)
    reps = enc.encode(ids.cuda(), mask.cuda()).cpu().numpy()
    sel = list(range(0, 256, 7))
    _, oreps = oracle.encode_reps(sd, EncoderSpec("bert", layers, H, 12, F, 1e-12), ids[sel], mask[sel])
    _check(reps[sel], oreps.numpy(), "reps (B = 256)")
    # batch composition must not matter: the same sequences encoded alone give the same representations
    alone = enc.encode(ids[sel].cuda(), mask[sel].cuda()).cpu().numpy()
    np.testing.assert_allclose(alone, reps[sel], rtol=0, atol=2e-2)


@pytest.mark.parametrize("L,B", [(128, 4), (32, 7), (50, 3)])
def test_t5_base_vs_oracle(enc_mod, L, B):
    gen = torch.Generator().manual_seed(200 + L)
    layers, H, heads, F, vocab = 3, 768, 12, 3072, 2000
    sd = _rand_t5_sd(gen, layers, H, heads, F, vocab)
    head_w = torch.randn(768, 768, generator=gen) * 768 ** -0.5
    spec = dict(arch="t5", layers=layers, hidden=H, heads=heads, ffn=F, vocab=vocab, ln_eps=1e-6, rel_buckets=32,
                rel_max_distance=128)
    enc = enc_mod.CudaEncoder(spec, sd, head_weight=head_w, pooling="mean", normalize=True, max_batch_tokens=1024)
    ids, mask = _ids(gen, B, L, vocab)
    hidden, reps = enc.encode(ids.cuda(), mask.cuda(), return_hidden=True)
    ospec = EncoderSpec("t5", layers, H, heads, F, 1e-6, pooling="mean", normalize=True)
    oh, oreps = oracle.encode_reps(sd, ospec, ids, mask, head_weight=head_w)
    m = mask.numpy().astype(bool)
    _check(hidden.cpu().numpy()[m], oh.numpy()[m], "hidden")
    _check(reps.cpu().numpy(), oreps.numpy(), "reps")


@pytest.mark.parametrize("L,B", [(256, 3), (384, 2), (512, 2)])
def test_bert_long_sequences_vs_oracle(enc_mod, L, B):
    # sequences longer than one attention tile: online softmax over 128-key tiles (attn_long_kernel)
    gen = torch.Generator().manual_seed(300 + L)
    layers, H, F, vocab = 2, 768, 3072, 2000
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 512)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=12, ffn=F, vocab=vocab, max_pos=512, type_vocab=2,
                ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, pooling="first", max_batch_tokens=B * L)
    ids, mask = _ids(gen, B, L, vocab)
    mask[1, 5:] = 0  # whole key tiles masked out: the running max must survive tiles without any allowed key
    ids[1, 5:] = 0
    tt = torch.randint(0, 2, (B, L), generator=gen)
    hidden, reps = enc.encode(ids.cuda(), mask.cuda(), tt.cuda(), return_hidden=True)
    ospec = EncoderSpec("bert", layers, H, 12, F, 1e-12, pooling="first")
    oh, oreps = oracle.encode_reps(sd, ospec, ids, mask, tt)
    m = mask.numpy().astype(bool)
    _check(reps.cpu().numpy(), oreps.numpy(), "reps")
    _check(hidden.cpu().numpy()[m], oh.numpy()[m], "hidden")


@pytest.mark.parametrize("L,B", [(256, 2), (512, 1)])
def test_t5_long_sequences_vs_oracle(enc_mod, L, B):
    # relative-position bias beyond +-127 (bucket saturation at max_distance) through the 1023-entry table
    gen = torch.Generator().manual_seed(400 + L)
    layers, H, heads, F, vocab = 2, 768, 12, 3072, 2000
    sd = _rand_t5_sd(gen, layers, H, heads, F, vocab)
    head_w = torch.randn(768, 768, generator=gen) * 768 ** -0.5
    spec = dict(arch="t5", layers=layers, hidden=H, heads=heads, ffn=F, vocab=vocab, ln_eps=1e-6, rel_buckets=32,
                rel_max_distance=128)
    enc = enc_mod.CudaEncoder(spec, sd, head_weight=head_w, pooling="mean", normalize=True, max_batch_tokens=B * L)
    ids, mask = _ids(gen, B, L, vocab)
    hidden, reps = enc.encode(ids.cuda(), mask.cuda(), return_hidden=True)
    ospec = EncoderSpec("t5", layers, H, heads, F, 1e-6, pooling="mean", normalize=True)
    oh, oreps = oracle.encode_reps(sd, ospec, ids, mask, head_weight=head_w)
    m = mask.numpy().astype(bool)
    _check(hidden.cpu().numpy()[m], oh.numpy()[m], "hidden")
    _check(reps.cpu().numpy(), oreps.numpy(), "reps")


def test_encoder_errors(enc_mod):
    gen = torch.Generator().manual_seed(1)
    sd = _rand_bert_sd(gen, 1, 128, 256, 100, 64)
    spec = dict(arch="bert", layers=1, hidden=128, heads=2, ffn=256, vocab=100, max_pos=64, type_vocab=2, ln_eps=1e-12)
    enc = enc_mod.CudaEncoder(spec, sd, max_batch_tokens=256)
    ids, mask = _ids(gen, 2, 16, 100)
    with pytest.raises(RuntimeError):
        enc.encode(ids, mask)  # CPU tensors: no CPU path
    with pytest.raises(RuntimeError):
        enc.encode(torch.zeros(1, 130, dtype=torch.long).cuda(), torch.ones(1, 130, dtype=torch.long).cuda())  # 128 < L, L % 128 != 0
    with pytest.raises(RuntimeError):
        enc.encode(torch.zeros(64, 16, dtype=torch.long).cuda(), torch.ones(64, 16, dtype=torch.long).cuda())
    del sd["encoder.layer.0.output.dense.bias"]
    with pytest.raises(RuntimeError, match="missing"):
        enc_mod.CudaEncoder(spec, sd, max_batch_tokens=256)
