"""Scratch probe (not part of the product): encoder throughput on the GPU, synthetic bert-base / t5-base / bert-large.
  python tools/encoder_probe.py [bert|t5|large] [B] [L] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmatch_b200 import synthetic  # noqa: E402
from openmatch_b200.encoder import CudaEncoder  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "bert"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
if arch == "t5":
    spec = dict(synthetic.T5_BASE)
    enc = CudaEncoder(spec, synthetic.t5_state_dict(spec, seed=0), head_weight=torch.randn(768, 768) * 0.03, pooling="mean",
                      normalize=True, max_batch_tokens=B * L)
else:
    spec = dict(synthetic.BERT_LARGE if arch == "large" else synthetic.BERT_BASE)
    enc = CudaEncoder(spec, synthetic.bert_state_dict(spec, seed=0), pooling="first", max_batch_tokens=B * L)
H, layers = spec["hidden"], spec["layers"]
ids, mask = synthetic.token_batch(B, L, spec["vocab"], seed=1, bert=arch != "t5", device="cuda")
out = torch.empty(B, enc.rep_dim, device="cuda")
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
