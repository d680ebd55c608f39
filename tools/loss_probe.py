"""Scratch probe: contrastive loss fwd+bwd latency / one call per shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmatch_b200 import _lib as om_lib  # noqa: E402

lib = om_lib.load()
dev = torch.device("cuda:0")
for bq, bp in ((64, 512), (512, 4096), (512, 4096), (128, 1024)):
    g = torch.Generator().manual_seed(1)
    xq = (torch.randn(bq, 768, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    xp = (torch.randn(bp, 768, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    lo = torch.empty((), device=dev)
    dq, dp = torch.empty(bq, 768, device=dev), torch.empty(bp, 768, device=dev)
    reps = int(os.environ.get("OM_REPS", 1))
    which = os.environ.get("OM_PROBE_GRADS", "both")
    dq_ptr = dq.data_ptr() if which in ("both", "dq") else None
    dp_ptr = dp.data_ptr() if which in ("both", "dp") else None
    for it in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            om_lib.check(lib.om_contrastive_loss_fwd_bwd(xq.data_ptr(), xp.data_ptr(), om_lib.OM_BF16, bq, bp, 768, None,
                                                         om_lib.OM_REDUCE_MEAN, 1.0, lo.data_ptr(), dq_ptr,
                                                         dp_ptr, None, om_lib.current_stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
    import ctypes
    ph = (ctypes.c_uint64 * 4)()
    om_lib.check(lib.om_debug_loss_phase_ns(ph))
    print("loss %dx%d: %.1f us/call, loss=%.5f  phases ns prep/logits/softmax/grads = %s" %
          (bq, bp, e0.elapsed_time(e1) / reps * 1e3, lo.item(), list(ph)))
