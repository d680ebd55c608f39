"""CPU, world_size 2 over gloo: the host-side exchange logic of the row-sharded search (shard offsets,
rank-ordered all-gather of per-shard top-k lists, merge) with the oracle standing in for the CUDA kernels,
and the autograd-aware all-gather used for cross-device negatives."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_merge(Dp, Ip, k):
    D, I = oracle.merge_topk([(Dp[i].numpy(), Ip[i].numpy()) for i in range(Dp.shape[0])], k)
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from openmatch_b200.index import exchange_and_merge, shard_offsets
        rng = np.random.default_rng(0)
        x = rng.integers(-4, 5, (1000, 32)).astype(np.float32)
        q = rng.integers(-4, 5, (13, 32)).astype(np.float32)
        bounds = [0, 380, 1000]  # uneven shards
        mine = x[bounds[rank]:bounds[rank + 1]]
        offset, total = shard_offsets(mine.shape[0])
        assert (offset, total) == (bounds[rank], 1000)
        D, I = oracle.flat_ip_search(q, mine, 25)
        I = np.where(I >= 0, I + offset, I)
        Dm, Im = exchange_and_merge(torch.from_numpy(D), torch.from_numpy(I), 25, merge=_oracle_merge)
        D0, I0 = oracle.flat_ip_search(q, x, 25)
        assert (Im.numpy() == I0).all() and (Dm.numpy() == D0).all()

        # the three-phase protocol (range MAX, histogram SUM, pruned re-score, narrow exchange) through the very
        # function the ranks run on GPUs, with the oracle's restatement of the three kernels
        from openmatch_b200.index import sharded_search_device
        for lo_v, hi_v, k in ((-4, 5, 25), (0, 2, 40)):  # second case: heavy score ties
            xs = rng.integers(lo_v, hi_v, (3000, 32)).astype(np.float32)
            qs = rng.integers(lo_v, hi_v, (9, 32)).astype(np.float32)
            b2 = [0, 1700, 3000]
            shard = oracle.ShardPhases(xs[b2[rank]:b2[rank + 1]], slack=16)
            Dm, Im = sharded_search_device(shard, torch.from_numpy(qs), k, b2[rank], merge=_oracle_merge)
            D0, I0 = oracle.flat_ip_search(qs, xs, k)
            assert (Im.numpy() == I0).all() and (Dm.numpy() == D0).all()

        # cross-device negatives: gathered tensor is rank-major and only the local slice carries gradient
        from openmatch_b200.modeling.dense_retrieval_model import DRModel
        m = DRModel.__new__(DRModel)
        torch.nn.Module.__init__(m)
        m.world_size, m.process_rank = world, rank
        t = torch.full((2, 3), float(rank + 1), requires_grad=True)
        g = m.dist_gather_tensor(t)
        assert g.shape == (4, 3) and g[:2].eq(1).all() and g[2:].eq(2).all()
        g.sum().backward()
        assert t.grad.eq(1).all()
        with open(os.path.join(tmp, "ok.%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / ("ok.%d" % r)) for r in range(2))
