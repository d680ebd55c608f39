"""CUDA encoder handle: host-side marshalling for ``om_encoder_*`` / ``om_encode`` (csrc/encoder.cu).

Replaces the HF forward + pooling + head + normalise sequence of ``DRModel.encode``
(``src/openmatch/modeling/dense_retrieval_model.py:133-155``) for inference.  Weights are handed over by
their HuggingFace ``state_dict`` names; the library keeps packed bf16 / fp32 copies in HBM.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib

_BERT_KEYS = ("num_hidden_layers", "hidden_size", "num_attention_heads", "intermediate_size", "vocab_size",
              "max_position_embeddings", "type_vocab_size", "layer_norm_eps")


def spec_from_hf_config(config) -> Dict:
    """Translate a HF ``BertConfig`` / ``T5Config`` into the plain dict ``CudaEncoder`` consumes."""
    mt = getattr(config, "model_type", "")
    if mt == "bert":
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise ValueError("CUDA encoder supports hidden_act='gelu' (erf) only, got %r" % config.hidden_act)
        if getattr(config, "position_embedding_type", "absolute") not in (None, "absolute"):
            raise ValueError("CUDA encoder supports absolute position embeddings only")
        return dict(arch="bert", layers=config.num_hidden_layers, hidden=config.hidden_size,
                    heads=config.num_attention_heads, ffn=config.intermediate_size, vocab=config.vocab_size,
                    max_pos=config.max_position_embeddings, type_vocab=config.type_vocab_size,
                    ln_eps=config.layer_norm_eps)
    if mt == "t5":
        if config.d_kv != 64:
            raise ValueError("CUDA encoder supports d_kv == 64 only")
        if getattr(config, "feed_forward_proj", "relu") != "relu":
            raise ValueError("CUDA encoder supports the non-gated ReLU T5 feed-forward only")
        return dict(arch="t5", layers=config.num_layers, hidden=config.d_model, heads=config.num_heads,
                    ffn=config.d_ff, vocab=config.vocab_size, max_pos=0, type_vocab=0,
                    ln_eps=config.layer_norm_epsilon, rel_buckets=config.relative_attention_num_buckets,
                    rel_max_distance=getattr(config, "relative_attention_max_distance", 128))
    raise ValueError("CUDA encoder supports BERT and T5-encoder backbones, got model_type=%r" % mt)


class CudaEncoder:
    def __init__(self, spec: Dict, state_dict: Dict[str, torch.Tensor], head_weight: Optional[torch.Tensor] = None,
                 pooling: str = "first", normalize: bool = False, max_batch_tokens: int = 256 * 128):
        if pooling not in ("first", "mean"):
            raise ValueError("Unknown pooling type: {}".format(pooling))
        self._lib = _lib.load()
        head_head_out = int(head_weight.shape[0]) if head_weight is not None else 0
        if spec["heads"] * 64 != spec["hidden"] and spec["arch"] == "bert":
            raise ValueError("CUDA encoder needs 64-wide attention heads")
        desc = _lib.EncoderDesc(
            arch=_lib.OM_ARCH_BERT if spec["arch"] == "bert" else _lib.OM_ARCH_T5ENC, layers=spec["layers"],
            hidden=spec["hidden"], heads=spec["heads"], ffn=spec["ffn"], vocab=spec["vocab"],
            max_pos=spec.get("max_pos", 0), type_vocab=spec.get("type_vocab", 0), ln_eps=float(spec["ln_eps"]),
            pooling=_lib.OM_POOL_MEAN if pooling == "mean" else _lib.OM_POOL_FIRST,
            has_head=1 if head_weight is not None else 0, head_out=head_head_out, normalize=1 if normalize else 0,
            rel_buckets=spec.get("rel_buckets", 32), rel_max_distance=spec.get("rel_max_distance", 128),
            max_batch_tokens=int(max_batch_tokens))
        h = ctypes.c_void_p()
        _lib.check(self._lib.om_encoder_create(ctypes.byref(desc), ctypes.byref(h)))
        self._h = h
        self.spec = dict(spec)
        self.hidden = spec["hidden"]
        self.max_batch_tokens = int(max_batch_tokens)
        self.ignored = []
        for name, t in state_dict.items():
            self._set(name, t)
        if head_weight is not None:
            self._set("head.linear.weight", head_weight)
        _lib.check(self._lib.om_encoder_finalize(self._h))
        self.rep_dim = int(self._lib.om_encoder_rep_dim(self._h))

    @classmethod
    def from_hf(cls, lm, head=None, pooling="first", normalize=False, max_batch_tokens=256 * 128):
        head_w = head.linear.weight if head is not None else None
        return cls(spec_from_hf_config(lm.config), lm.state_dict(), head_w, pooling, normalize, max_batch_tokens)

    def _set(self, name: str, t: torch.Tensor):
        t = t.detach()
        if not t.is_floating_point():
            return
        t = t.to(torch.float32).contiguous()
        kind = _lib.OM_DEVICE if t.is_cuda else _lib.OM_HOST
        shape = (ctypes.c_int64 * t.dim())(*t.shape)
        rc = _lib.check(self._lib.om_encoder_set_weight(self._h, name.encode(), t.data_ptr(), kind, shape, t.dim()))
        if rc == 1:
            self.ignored.append(name)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.om_encoder_destroy(h)

    @torch.no_grad()
    def encode(self, input_ids: torch.Tensor, attention_mask: torch.Tensor,
               token_type_ids: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               out_dtype: torch.dtype = torch.float32, return_hidden: bool = False):
        """int64 [B, L] CUDA tensors in -> reps [B, rep_dim] (and last_hidden_state fp32 [B, L, H])."""
        if not input_ids.is_cuda:
            raise RuntimeError("openmatch_b200 encoder runs on CUDA tensors only (no CPU path)")
        B, L = input_ids.shape
        ids = input_ids.to(torch.int64).contiguous()
        mask = attention_mask.to(torch.int64).contiguous()
        tt = token_type_ids.to(torch.int64).contiguous() if token_type_ids is not None else None
        if out is None:
            out = torch.empty((B, self.rep_dim), dtype=out_dtype, device=ids.device)
        if out.dtype not in (torch.float32, torch.bfloat16) or out.stride(1) != 1:
            raise ValueError("out must be a row-major fp32 / bf16 CUDA tensor")
        hidden = torch.empty((B, L, self.hidden), dtype=torch.float32, device=ids.device) if return_hidden else None
        _lib.check(self._lib.om_encode(
            self._h, ids.data_ptr(), mask.data_ptr(), tt.data_ptr() if tt is not None else None, B, L,
            out.data_ptr(), _lib.OM_F32 if out.dtype == torch.float32 else _lib.OM_BF16, out.stride(0),
            hidden.data_ptr() if hidden is not None else None, _lib.current_stream_ptr()))
        return (hidden, out) if return_hidden else out
