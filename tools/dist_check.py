"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): row-sharded exact search through
openmatch_b200.index.ShardedFlatIPIndex vs the CPU oracle on the unsharded corpus, and the cross-device
contrastive loss (all-gather + fused kernel) vs the oracle on the gathered batch.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from openmatch_b200.index import ShardedFlatIPIndex  # noqa: E402
from openmatch_b200.loss import DistributedContrastiveLoss  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
rng = np.random.default_rng(0)  # same data on every rank
n, d, nq, k = 40000, 128, 77, 100
x = rng.integers(-6, 7, (n, d)).astype(np.float32)
q = rng.integers(-6, 7, (nq, d)).astype(np.float32)
bounds = np.linspace(0, n, world + 1).astype(int)
bounds[1:-1] += 37  # uneven shards
idx = ShardedFlatIPIndex(d)
idx.add_local(x[bounds[rank]:bounds[rank + 1]])
idx.finalize_offsets()
assert idx.ntotal == n and idx.offset == bounds[rank]
D, I = idx.search(q, k)
D0, I0 = oracle.flat_ip_search(q, x, k)
assert (I == I0).all() and (D == D0).all(), "sharded search differs from the oracle on rank %d" % rank

# cross-device negatives: each rank holds 4 queries x 8 passages
g = torch.Generator().manual_seed(5)
allx = (torch.randn(world * 4, 64, generator=g) * 0.5).to(torch.bfloat16)
ally = (torch.randn(world * 32, 64, generator=g) * 0.5).to(torch.bfloat16)
lx = allx[rank * 4:(rank + 1) * 4].cuda().requires_grad_()
ly = ally[rank * 32:(rank + 1) * 32].cuda().requires_grad_()
loss = DistributedContrastiveLoss()(lx, ly)
loss.backward()
want, dx, dy, _ = oracle.contrastive_loss_fwd_bwd(allx.float().numpy(), ally.float().numpy())
assert abs(loss.item() - want * world) <= 1e-3 * max(1.0, abs(want * world)), (loss.item(), want * world)
gx = lx.grad.float().cpu().numpy() / world
ref = dx[rank * 4:(rank + 1) * 4]
assert np.linalg.norm(gx - ref) <= 1e-2 * np.linalg.norm(ref), "x-device grad mismatch"
dist.barrier()
if rank == 0:
    print("DIST CHECK OK: world=%d sharded search exact, x-device loss %.5f (oracle %.5f x %d)" % (world, loss.item(), want, world))
dist.destroy_process_group()
