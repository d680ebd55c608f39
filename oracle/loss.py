"""Oracle: in-batch-negatives contrastive loss (fp32, CPU).

Follows ``src/openmatch/loss.py:7-15`` (``SimpleContrastiveLoss``: ``logits = x @ y.T``;
``F.cross_entropy(logits, target, reduction)``; default ``target = arange(0, nq*tpq, tpq)`` with
``tpq = y.size(0) // x.size(0)``) and the identical arithmetic in ``DRModel.forward``
(``src/openmatch/modeling/dense_retrieval_model.py:113-125``: ``scores = q_reps @ p_reps.T``,
``target = arange(nq) * train_n_passages``, mean cross-entropy, ``* world_size`` when negatives are
gathered across devices).
"""
from __future__ import annotations

import numpy as np


def _default_target(nq: int, n_p: int) -> np.ndarray:
    tpq = n_p // nq
    return np.arange(0, nq * tpq, tpq, dtype=np.int64)


def contrastive_loss(x: np.ndarray, y: np.ndarray, target=None, reduction: str = "mean", scale: float = 1.0):
    """Loss value only (float64 accumulation of the fp32 logits for a tight reference)."""
    loss, _, _, _ = contrastive_loss_fwd_bwd(x, y, target, reduction, scale)
    return loss


def contrastive_loss_fwd_bwd(x, y, target=None, reduction: str = "mean", scale: float = 1.0):
    """Returns (loss, dX, dY, logits).

    loss = scale * reduce_i( logsumexp_j s_ij - s_i,t_i ),  s = x @ y.T
    dS   = scale * w * (softmax(s) - onehot(t)),  w = 1/nq for 'mean', 1 for 'sum'
    dX   = dS @ y,   dY = dS.T @ x
    """
    x64 = np.asarray(x, dtype=np.float64)
    y64 = np.asarray(y, dtype=np.float64)
    nq, n_p = x64.shape[0], y64.shape[0]
    if target is None:
        target = _default_target(nq, n_p)
    target = np.asarray(target, dtype=np.int64)
    s = (np.asarray(x, np.float32) @ np.asarray(y, np.float32).T).astype(np.float64)
    m = s.max(axis=1, keepdims=True)
    e = np.exp(s - m)
    z = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(z))[:, 0]
    per_row = lse - s[np.arange(nq), target]
    if reduction == "mean":
        w = 1.0 / nq
        loss = per_row.mean()
    elif reduction == "sum":
        w = 1.0
        loss = per_row.sum()
    else:
        raise ValueError("reduction must be 'mean' or 'sum'")
    g = e / z
    g[np.arange(nq), target] -= 1.0
    g *= w * scale
    dx = g @ y64
    dy = g.T @ x64
    return float(loss * scale), dx.astype(np.float32), dy.astype(np.float32), s.astype(np.float32)
