"""Scratch probe (not part of the product): encoder throughput on the GPU, synthetic bert-base / t5-base."""
import sys

import torch

sys.path.insert(0, "tests")
from test_encoder_gpu import _rand_bert_sd, _rand_t5_sd  # noqa: E402

from openmatch_b200.encoder import CudaEncoder  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "bert"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
gen = torch.Generator().manual_seed(0)
H, F, layers, vocab = 768, 3072, 12, 30522
if arch == "bert":
    sd = _rand_bert_sd(gen, layers, H, F, vocab, 512)
    spec = dict(arch="bert", layers=layers, hidden=H, heads=12, ffn=F, vocab=vocab, max_pos=512, type_vocab=2, ln_eps=1e-12)
    enc = CudaEncoder(spec, sd, pooling="first", max_batch_tokens=B * L)
else:
    sd = _rand_t5_sd(gen, layers, H, 12, F, 32128)
    spec = dict(arch="t5", layers=layers, hidden=H, heads=12, ffn=F, vocab=32128, ln_eps=1e-6)
    enc = CudaEncoder(spec, sd, head_weight=torch.randn(768, 768) * 0.03, pooling="mean", normalize=True, max_batch_tokens=B * L)
ids = torch.randint(1000, 30000, (B, L), device="cuda")
mask = torch.ones(B, L, dtype=torch.long, device="cuda")
out = torch.empty(B, 768, device="cuda")
for _ in range(3):
    enc.encode(ids, mask, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    enc.encode(ids, mask, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = layers * L * (24 * H * H + 4 * L * H) * B
print(f"{arch} B={B} L={L}: {ms:.3f} ms/batch  {B/ms*1e3:.0f} seq/s  {flop/ms/1e9:.0f} TFLOP/s (algorithmic)")
