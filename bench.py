#!/usr/bin/env python
"""Headline benchmark of the dense-retrieval hot path (BASELINE.json):

  metric  : queries/sec, top-1000 over an 8.8M x 768 corpus (configs[1]: bert-base 768-d, 6 980 queries,
            brute force on 1 x B200); with --gpus N the same corpus is row-sharded over N GPUs and the
            per-shard top-k lists are all-gathered over NCCL and merged (configs[2]) -> strong scaling.
  also    : passages encoded/sec (bert-base, L=128, batch 256 per GPU) in the "encode" object.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # CPU arm: oracle port of faiss IndexFlatIP on a bounded sample

A step = one search of the whole query batch against the HBM-resident corpus (value: inputs resident in HBM;
e2e: host fp32 queries in, host (D, I) out, copies inside the timed region).  Synthetic data: corpus and
queries i.i.d. N(0,1) fp32 (seeded), random-init bert-base weights.  Timed with CUDA events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--corpus", type=int, default=8_800_000)
    ap.add_argument("--nq", type=int, default=6980)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--encode-batch", type=int, default=256)
    ap.add_argument("--skip-encode", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-train", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))), "hbm": float(p["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons, "samples": len(sm)}


def cpu_reference_search(args, steps, warmup, threads=None):
    """The reference's CPU path for this step = faiss IndexFlatIP.search; faiss is not installable here, so the
    oracle port (blocked fp32 SGEMM + exact top-k, all host threads via BLAS) is timed on a bounded sample:
    `nq_s` queries against `n_s` rows, extrapolated linearly in corpus rows to the full corpus."""
    import numpy as np

    import oracle
    cores = os.cpu_count() or 1
    n_s, nq_s = min(args.corpus, 400_000), min(args.nq, 256)
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((n_s, args.dim), dtype=np.float32)
    q = rng.standard_normal((nq_s, args.dim), dtype=np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.flat_ip_search(q, x, min(args.k, n_s))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    qps_full = nq_s / t * (n_s / args.corpus)
    return {"value": qps_full, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": "oracle.flat_ip_search (numpy BLAS SGEMM + exact top-%d), %d queries x %d rows x %d dims per step, "
                      "%.2f s/step, extrapolated linearly in rows to %d" % (min(args.k, n_s), nq_s, n_s, args.dim, t, args.corpus),
            "ms_per_step": t * 1e3}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        base = cpu_reference_search(args, max(1, args.steps), max(0, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": "queries/sec top-1000 over 8.8M x 768 corpus", "value": base["value"],
                "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: top-%d over %d x %d, %d queries (CPU sample extrapolated)" % (
                    args.k, args.corpus, args.dim, args.nq), "corpus_rows": args.corpus, "dim": args.dim, "k": args.k},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist

    from openmatch_b200 import synthetic
    from openmatch_b200.encoder import CudaEncoder
    from openmatch_b200.index import FlatIPIndex, sharded_search_device

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    assert world == args.gpus or world == 1, "--gpus must match WORLD_SIZE under torchrun"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- corpus shard: rows [lo, hi) of the global corpus, generated straight into HBM ----------------
    d, k, nq = args.dim, args.k, args.nq
    per = (args.corpus + world - 1) // world
    lo, hi = rank * per, min(args.corpus, (rank + 1) * per)
    idx = FlatIPIndex(d)
    chunk = 550_000
    for c0 in range(lo, hi, chunk):
        n = min(chunk, hi - c0)
        rows = idx.reserve_rows(n)
        g = torch.Generator(device=dev).manual_seed(1234 + c0 // chunk + 7919 * rank)
        rows.normal_(generator=g)
        idx.commit_rows(n)
    gq = torch.Generator(device=dev).manual_seed(99)
    q_dev = torch.randn(nq, d, generator=gq, device=dev)
    q_host = q_dev.cpu().pin_memory()
    D_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    I_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()

    def search_step(q):
        # world > 1: local scan -> all-reduce(MAX) of per-query floors -> pruned fp32 re-score -> NCCL all-gather of
        # the [nq, k] lists -> merge kernel
        return sharded_search_device(idx, q, k, lo)

    def e2e_step():
        if world == 1:
            idx.search_pinned(q_host, k, D_host, I_host)  # C-ABI call with HOST buffers: H2D + search + D2H inside
        else:
            D, I = search_step(q_host.to(dev, non_blocking=True))
            D_host.copy_(D, non_blocking=True)
            I_host.copy_(I, non_blocking=True)
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- device-resident search (value) + per-kernel device time for the roofline ----------------
    idx.set_param("profile", 1)
    scan_ns = select_ns = final_ns = launches = rounds = 0

    def value_step():
        nonlocal scan_ns, select_ns, final_ns, launches, rounds
        search_step(q_dev)
        scan_ns += idx.stat("scan_ns")
        select_ns += idx.stat("select_ns")
        final_ns += idx.stat("finalize_ns")
        launches += idx.stat("launches") + (1 if world > 1 else 0)  # + merge kernel
        rounds = idx.stat("rounds")

    for _ in range(args.warmup):
        search_step(q_dev)
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(value_step, args.steps, 0)
    idx.set_param("profile", 0)
    ms_per_step = total_ms / args.steps
    qps = nq / (ms_per_step * 1e-3)

    e2e_ms = timed(e2e_step, args.steps, min(args.warmup, 2)) / args.steps
    e2e_qps = nq / (e2e_ms * 1e-3)

    # ---------------- streaming regime: few queries per pass -> one sweep of the bf16 shard is
    # HBM-bound.  Local shard only (no exchange), k = 100; roofline = rows_local * d * 2 B per sweep / measured HBM
    streaming = None
    if not args.skip_encode:
        try:
            streaming = {"unit": "ms per search over this GPU's shard", "k": 100, "cases": {}}
            hbm = measured_peaks().get("hbm")
            sweep_bytes = (hi - lo) * d * 2.0
            for snq in (1, 16, 64):
                qs = q_dev[:snq].contiguous()

                def stream_step():
                    idx.search_device(qs, 100)

                sms_ = timed(stream_step, 10, 3) / 10
                case = {"ms": sms_, "queries_per_s": world * snq / (sms_ * 1e-3), "achieved_gbs": sweep_bytes / (sms_ * 1e-3) / 1e9}
                if hbm:
                    case["frac_of_hbm_peak"] = case["achieved_gbs"] / hbm
                streaming["cases"]["nq=%d" % snq] = case
        except Exception as e:  # informational leg
            streaming = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---------------- encoder throughput (bert-base, L=128, one batch per step per GPU) ----------------
    encode = None
    if not args.skip_encode:
        spec = dict(synthetic.BERT_BASE)
        B, L = args.encode_batch, 128
        enc = CudaEncoder(spec, synthetic.bert_state_dict(spec, seed=0), pooling="first", max_batch_tokens=B * L)
        ids, mask = synthetic.token_batch(B, L, spec["vocab"], seed=1234 + rank, device=dev)
        out = torch.empty((B, 768), dtype=torch.float32, device=dev)
        ids_h, mask_h = ids.cpu().pin_memory(), mask.cpu().pin_memory()
        out_h = torch.empty((B, 768), dtype=torch.float32).pin_memory()

        def enc_step():
            enc.encode(ids, mask, out=out)

        def enc_e2e():
            i, m = ids_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True)
            enc.encode(i, m, out=out)
            out_h.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

        enc_ms = timed(enc_step, max(args.steps, 5), 3) / max(args.steps, 5)
        enc_e2e_ms = timed(enc_e2e, max(args.steps, 5), 2) / max(args.steps, 5)
        flop = spec["layers"] * L * (24 * 768 * 768 + 4 * L * 768) * B
        peaks = measured_peaks()
        encode = {"metric": "passages encoded/sec (bert-base, L=128)", "value": world * B / (enc_ms * 1e-3),
                  "unit": "passages/s", "ms_per_step": enc_ms, "batch_per_gpu": B,
                  "e2e": {"value": world * B / (enc_e2e_ms * 1e-3), "unit": "passages/s",
                          "h2d_bytes_per_step": 2 * B * L * 8, "d2h_bytes_per_step": B * 768 * 4},
                  "roofline": {"bound": "tensor", "achieved": flop / (enc_ms * 1e-3) / 1e12, "peak": peaks["tflops"],
                               "unit": "TFLOP/s", "frac": flop / (enc_ms * 1e-3) / 1e12 / peaks["tflops"],
                               "note": "whole encoder step (22.35 GFLOP/passage algorithmic) / step time; " + peaks["source"]}}
        del enc
        # C3's encoder: t5-base (GTR) + masked mean pooling + bias-free 768x768 head + L2 normalisation
        try:
            tspec = dict(synthetic.T5_BASE)
            tenc = CudaEncoder(tspec, synthetic.t5_state_dict(tspec, seed=0),
                               head_weight=torch.randn(768, 768, generator=torch.Generator().manual_seed(3)) * 0.03,
                               pooling="mean", normalize=True, max_batch_tokens=B * L)
            tids, tmask = synthetic.token_batch(B, L, tspec["vocab"], seed=4321 + rank, bert=False, device=dev)

            def t5_step():
                tenc.encode(tids, tmask, out=out)

            t5_ms = timed(t5_step, max(args.steps, 5), 3) / max(args.steps, 5)
            encode["t5_base_gtr"] = {"value": world * B / (t5_ms * 1e-3), "unit": "passages/s", "ms_per_step": t5_ms,
                                     "frac_of_peak": flop / (t5_ms * 1e-3) / 1e12 / peaks["tflops"]}
            del tenc
        except Exception as e:  # informational leg
            encode["t5_base_gtr"] = {"error": "%s: %s" % (type(e).__name__, e)}

    # contrastive loss fwd+bwd (C4): local negatives [64, 768] x [512, 768] and the cross-device-at-8 shape
    # [512, 768] x [4096, 768]; one cooperative tcgen05 kernel per call (latency-bound: reported in microseconds)
    loss_obj = None
    if not args.skip_encode:
        from openmatch_b200 import _lib as om_lib
        lib = om_lib.load()
        loss_obj = {"metric": "contrastive loss fwd+bwd latency", "unit": "us", "kernel_launches_per_call": 1, "shapes": {}}
        for name, (bq, bp) in {"local_64x512": (64, 512), "xdevice8_512x4096": (512, 4096)}.items():
            g = torch.Generator(device="cpu").manual_seed(1234)
            xq = (torch.randn(bq, 768, generator=g) * 0.5).to(torch.bfloat16).to(dev)
            xp = (torch.randn(bp, 768, generator=g) * 0.5).to(torch.bfloat16).to(dev)
            lo_t = torch.empty((), dtype=torch.float32, device=dev)
            dxq, dxp = torch.empty(bq, 768, device=dev), torch.empty(bp, 768, device=dev)
            reps = 50

            def loss_step():
                for _ in range(reps):
                    om_lib.check(lib.om_contrastive_loss_fwd_bwd(
                        xq.data_ptr(), xp.data_ptr(), om_lib.OM_BF16, bq, bp, 768, None, om_lib.OM_REDUCE_MEAN, 1.0,
                        lo_t.data_ptr(), dxq.data_ptr(), dxp.data_ptr(), None, om_lib.current_stream_ptr()))

            us = timed(loss_step, 3, 3) / 3 / reps * 1e3
            loss_obj["shapes"][name] = {"us": us, "tflops": 6.0 * bq * bp * 768 / (us * 1e-6) / 1e12}

    # contrastive training step (C4): bert-base, 64 queries (L=32) x 8 passages (L=128) per GPU, bf16 autocast.
    # Encoder forward/backward = the HF torch module under autograd (our encoder kernels are forward-only, DESIGN
    # section 6); loss forward+backward = loss_fused_kernel; AdamW step included; DDP all-reduce when world > 1.
    train_obj = None
    if not args.skip_encode and not args.skip_train:
        try:
            import types
            from transformers import BertConfig, BertModel
            from openmatch_b200.modeling import DRModel
            torch.manual_seed(0)
            lm = BertModel(BertConfig(), add_pooling_layer=False).to(dev)
            model = DRModel(lm, lm, tied=True, pooling="first",
                            data_args=types.SimpleNamespace(train_n_passages=8),
                            train_args=types.SimpleNamespace(negatives_x_device=False)).to(dev).train()
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if world > 1 else model
            opt = torch.optim.AdamW(net.parameters(), lr=5e-6, fused=True)
            qi, qm = synthetic.token_batch(64, 32, 30522, seed=77 + rank, device=dev)
            pi, pm = synthetic.token_batch(512, 128, 30522, seed=177 + rank, device=dev)
            qb_ = {"input_ids": qi, "attention_mask": qm, "token_type_ids": torch.zeros_like(qi)}
            pb_ = {"input_ids": pi, "attention_mask": pm, "token_type_ids": torch.zeros_like(pi)}

            def train_step():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = net(qb_, pb_).loss
                loss.backward()
                opt.step()
                opt.zero_grad(set_to_none=True)

            tr_ms = timed(train_step, 5, 3) / 5
            flop = 3 * (64 * 12 * 32 * (24 * 768 * 768 + 4 * 32 * 768) + 512 * 12 * 128 * (24 * 768 * 768 + 4 * 128 * 768))
            train_obj = {"metric": "train queries/sec (bert-base, 64 q x 8 psg per GPU, bf16)", "value": world * 64 / (tr_ms * 1e-3),
                         "unit": "queries/s", "ms_per_step": tr_ms, "tflops_per_gpu": flop / (tr_ms * 1e-3) / 1e12,
                         "note": "encoder fwd/bwd: HF torch module under autograd (cuBLAS/SDPA); loss fwd+bwd: "
                                 "loss_fused_kernel; fused AdamW; DDP all-reduce when n_gpus > 1"}
            del model, net, opt, lm
            torch.cuda.empty_cache()
        except Exception as e:  # informational leg: never take the search line down with it
            train_obj = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    n_local = hi - lo
    scan_flops = 2.0 * nq * n_local * d * args.steps
    scan_s = scan_ns * 1e-9
    achieved = scan_flops / scan_s / 1e12 if scan_s > 0 else None
    traffic = traffic_note = None
    tpath = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = "%s: %.3g B DRAM vs %.3g B algorithmic (x%.3f); %s" % (
            tj.get("launch"), traffic, tj.get("algorithmic_bytes_per_launch"), tj.get("ratio_traffic_over_algorithmic"),
            tj.get("source"))
    line = {
        "metric": "queries/sec top-1000 over 8.8M x 768 corpus", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1] search: top-%d over %d x %d fp32 corpus resident in HBM, %d queries per step"
                               % (k, args.corpus, d, nq), "corpus_rows": args.corpus, "rows_per_gpu": n_local, "dim": d,
                   "k": k, "nq": nq, "candidate_stage": "bf16 tensor-core scan + fp32 re-score", "rounds": rounds,
                   "parallelism": "index row-sharded x%d, NCCL all-gather of per-shard top-k + merge" % world if world > 1
                   else "single shard", "l2_policy": "inputs_exceed_l2 (bf16 scan copy %.1f GB per GPU)" % (n_local * d * 2 / 1e9)},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": nq * d * 4,
                "d2h_bytes_per_step": nq * k * 12},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops"] if achieved else None, "traffic": traffic,
                     "traffic_note": traffic_note,
                     "kernel": "gemm_bf16_tn_kernel<256,4,1,8,EpiScan> (fused Q*X^T + top-k filter)",
                     "note": "2*nq*rows*d FLOPs per sweep / CUDA-event time of the scan launches on the launching "
                             "stream; " + peaks["source"],
                     "phase_ms_per_step": {"scan": scan_ns / 1e6 / args.steps, "select": select_ns / 1e6 / args.steps,
                                           "finalize_rescore": final_ns / 1e6 / args.steps}},
        "clocks": clocks.summary(),
    }
    if encode:
        line["encode"] = encode
    if loss_obj:
        line["loss"] = loss_obj
    if train_obj:
        line["train"] = train_obj
    if streaming:
        line["streaming"] = streaming
    if world == 1 and not args.skip_cpu:
        base = cpu_reference_search(args, 1, 0)
        line["cpu_baseline"] = {k_: base[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
