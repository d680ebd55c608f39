"""Contrastive dense-retrieval trainer with the reference's surface (``src/openmatch/trainer/dense_trainer.py``):
``DRTrainer(model, args, train_dataset, data_collator, tokenizer, callbacks)`` with ``train()``,
``save_model()``, ``compute_loss``, ``training_step``, ``is_world_process_zero``.

The reference subclasses HF ``Trainer``; that class is unusable here (needs ``accelerate``, and its 5.x
hooks no longer match the reference's overrides), so this is a small standalone data-parallel loop: one
process per GPU under torchrun, ``DistributedDataParallel`` over NCCL, AdamW + linear warm-up / decay,
bf16 / fp16 autocast for the HF encoder, and the fused CUDA contrastive loss
(``openmatch_b200.loss``) for logits, log-softmax, loss and rep gradients.
"""
from __future__ import annotations

import logging
import math
import os
import types
from contextlib import nullcontext
from typing import Optional

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, IterableDataset

logger = logging.getLogger(__name__)
TRAINING_ARGS_NAME = "training_args.bin"


class _ShardByBatch(IterableDataset):
    """Each rank takes its ``per_device`` slice of every global batch (what HF's IterableDatasetShard does
    for the reference, dense_trainer.py:75-82)."""

    def __init__(self, dataset, per_device: int, world: int, rank: int):
        self.dataset, self.per_device, self.world, self.rank = dataset, per_device, world, rank

    def __iter__(self):
        group, pending = self.per_device * self.world, []
        for item in self.dataset:
            pending.append(item)
            if len(pending) == group:
                yield from pending[self.rank * self.per_device:(self.rank + 1) * self.per_device]
                pending = []
        if pending:  # pad the tail cyclically so every rank sees the same number of batches
            n0 = len(pending)
            while len(pending) < group:
                pending.append(pending[len(pending) % n0])
            yield from pending[self.rank * self.per_device:(self.rank + 1) * self.per_device]


class DRTrainer:
    def __init__(self, model, args, train_dataset=None, eval_dataset=None, data_collator=None, tokenizer=None,
                 callbacks=None, **_unused):
        self.model, self.args = model, args
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.data_collator, self.tokenizer = data_collator, tokenizer
        self.callbacks = callbacks or []
        self.state = types.SimpleNamespace(epoch=0.0, global_step=0, log_history=[])
        self._dist_loss_scale_factor = dist.get_world_size() if getattr(args, "negatives_x_device", False) else 1
        self._writer = None

    # ------------------------------------------------------------------ reference hooks
    def is_world_process_zero(self) -> bool:
        return self.args.process_index == 0

    def _prepare_inputs(self, inputs):
        dev = self.args.device
        return [x.to(dev) if isinstance(x, torch.Tensor) else {k: v.to(dev, non_blocking=True) for k, v in x.items()}
                for x in inputs]

    def compute_loss(self, model, inputs, return_outputs=False):
        query, passage = inputs
        outputs = model(query=query, passage=passage)
        return (outputs.loss, outputs) if return_outputs else outputs.loss

    def get_train_dataloader(self) -> DataLoader:
        if self.train_dataset is None:
            raise ValueError("Trainer: training requires a train_dataset.")
        ds = self.train_dataset
        if self.args.world_size > 1:
            ds = _ShardByBatch(ds, self.args.per_device_train_batch_size, self.args.world_size, self.args.process_index)
        return DataLoader(ds, batch_size=self.args.per_device_train_batch_size, collate_fn=self.data_collator,
                          drop_last=False, num_workers=self.args.dataloader_num_workers,
                          pin_memory=self.args.dataloader_pin_memory)

    def _autocast(self):
        if self.args.bf16:
            return torch.autocast("cuda", dtype=torch.bfloat16)
        if self.args.fp16:
            return torch.autocast("cuda", dtype=torch.float16)
        return nullcontext()

    def training_step(self, model, inputs) -> torch.Tensor:
        model.train()
        inputs = self._prepare_inputs(inputs)
        with self._autocast():
            loss = self.compute_loss(model, inputs)
        if self.args.gradient_accumulation_steps > 1:
            loss = loss / self.args.gradient_accumulation_steps
        if self._scaler is not None:
            self._scaler.scale(loss).backward()
        else:
            loss.backward()
        return loss.detach() / self._dist_loss_scale_factor

    # ------------------------------------------------------------------ loop
    def _steps_per_epoch(self) -> Optional[int]:
        try:
            n = len(self.train_dataset)
        except TypeError:
            return None
        per_step = self.args.per_device_train_batch_size * self.args.world_size * self.args.gradient_accumulation_steps
        return max(1, math.ceil(n / per_step))

    def train(self):
        args = self.args
        device = args.device
        self.model.to(device)
        wrapped = self.model
        if args.world_size > 1:
            wrapped = torch.nn.parallel.DistributedDataParallel(
                self.model, device_ids=[device.index] if device.type == "cuda" else None,
                find_unused_parameters=True)  # BertPooler params never receive gradients (OpenMatch ignores it)
        decay = [p for n, p in self.model.named_parameters() if p.requires_grad and not any(
            k in n for k in ("bias", "LayerNorm.weight", "layer_norm.weight"))]
        no_decay = [p for n, p in self.model.named_parameters() if p.requires_grad and any(
            k in n for k in ("bias", "LayerNorm.weight", "layer_norm.weight"))]
        opt = torch.optim.AdamW([{"params": decay, "weight_decay": args.weight_decay},
                                 {"params": no_decay, "weight_decay": 0.0}], lr=args.learning_rate,
                                betas=(args.adam_beta1, args.adam_beta2), eps=args.adam_epsilon)
        per_epoch = self._steps_per_epoch()
        if args.max_steps > 0:
            total = args.max_steps
        elif per_epoch is not None:
            total = int(per_epoch * args.num_train_epochs)
        else:
            raise ValueError("set --max_steps for a dataset without a length")
        warmup = args.warmup_steps if args.warmup_steps > 0 else int(total * args.warmup_ratio)

        def lr_lambda(step):  # linear warm-up then linear decay (HF 'linear' schedule)
            if step < warmup:
                return step / max(1, warmup)
            return max(0.0, (total - step) / max(1, total - warmup))

        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda)
        self._scaler = torch.amp.GradScaler("cuda") if args.fp16 else None
        if self.is_world_process_zero() and args.logging_dir:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._writer = SummaryWriter(args.logging_dir)
            except Exception:  # tensorboard is optional
                self._writer = None
        step, epoch, running = 0, 0, 0.0
        while step < total:
            self.state.epoch = float(epoch)
            micro = 0
            for inputs in self.get_train_dataloader():
                sync = (micro + 1) % args.gradient_accumulation_steps == 0
                ctx = nullcontext() if (sync or args.world_size == 1) else wrapped.no_sync()
                with ctx:
                    running += float(self.training_step(wrapped, inputs))
                micro += 1
                if not sync:
                    continue
                if self._scaler is not None:
                    self._scaler.unscale_(opt)
                if args.max_grad_norm and args.max_grad_norm > 0:
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), args.max_grad_norm)
                if self._scaler is not None:
                    self._scaler.step(opt)
                    self._scaler.update()
                else:
                    opt.step()
                sched.step()
                opt.zero_grad(set_to_none=True)
                step += 1
                self.state.global_step = step
                if step % args.logging_steps == 0:
                    entry = {"step": step, "loss": running / args.logging_steps, "learning_rate": sched.get_last_lr()[0]}
                    self.state.log_history.append(entry)
                    if self.is_world_process_zero():
                        logger.info("%s", entry)
                        if self._writer:
                            self._writer.add_scalar("train/loss", entry["loss"], step)
                            self._writer.add_scalar("train/learning_rate", entry["learning_rate"], step)
                    running = 0.0
                if args.save_steps and step % args.save_steps == 0 and self.is_world_process_zero():
                    self._save(os.path.join(args.output_dir, "checkpoint-{}".format(step)))
                if step >= total:
                    break
            epoch += 1
        if self._writer:
            self._writer.close()
        return types.SimpleNamespace(global_step=step, training_loss=running)

    # ------------------------------------------------------------------ checkpoints
    def _save(self, output_dir: Optional[str] = None):
        output_dir = output_dir if output_dir is not None else self.args.output_dir
        os.makedirs(output_dir, exist_ok=True)
        logger.info("Saving model checkpoint to %s", output_dir)
        self.model.save(output_dir)
        if self.tokenizer is not None:
            self.tokenizer.save_pretrained(output_dir)
        torch.save(self.args, os.path.join(output_dir, TRAINING_ARGS_NAME))

    def save_model(self, output_dir: Optional[str] = None):
        if self.is_world_process_zero():
            self._save(output_dir)


def split_dense_inputs(model_input: dict, chunk_size: int):
    """{'query' | 'passage': {name: tensor [B, ...]}} -> list of the same structure over row chunks
    (reference dense_trainer.py:111-120)."""
    assert len(model_input) == 1
    key, tensors = next(iter(model_input.items()))
    names = list(tensors.keys())
    pieces = [tensors[n].split(chunk_size, dim=0) for n in names]
    return [{key: dict(zip(names, chunk))} for chunk in zip(*pieces)]


def get_dense_rep(x):
    return x.p_reps if x.q_reps is None else x.q_reps


class GCDenseTrainer(DRTrainer):
    """Gradient-cache training (reference dense_trainer.py:130-160, which delegates to the un-vendored
    ``grad_cache`` package).  Restated here from GradCache's published algorithm:
      1. encode every query / passage chunk WITHOUT autograd (remembering the RNG state of each chunk);
      2. evaluate the contrastive loss on the concatenated representations — the fused CUDA loss kernel returns
         d loss / d reps directly, which is exactly the "gradient cache";
      3. re-encode each chunk WITH autograd under its saved RNG state and back-propagate the surrogate
         <reps_chunk, cached_grad_chunk>.
    Peak activation memory is one chunk; the in-batch negatives span the whole (optionally cross-device) batch."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        from ..loss import DistributedContrastiveLoss, SimpleContrastiveLoss
        self.loss_fn = DistributedContrastiveLoss() if self.args.negatives_x_device else SimpleContrastiveLoss()
        self.chunk_sizes = [self.args.gc_q_chunk_size, self.args.gc_p_chunk_size]

    def _encode_chunks(self, model, chunks, with_grad, rng_states=None):
        reps, states = [], []
        for i, chunk in enumerate(chunks):
            if rng_states is not None:
                torch.cuda.set_rng_state(rng_states[i])
            states.append(torch.cuda.get_rng_state())
            with self._autocast(), (nullcontext() if with_grad else torch.no_grad()):
                reps.append(get_dense_rep(model(**chunk)))
        return reps, states

    def training_step(self, model, inputs) -> torch.Tensor:
        model.train()
        queries, passages = self._prepare_inputs(inputs)
        q_chunks = split_dense_inputs({"query": queries}, self.chunk_sizes[0])
        p_chunks = split_dense_inputs({"passage": passages}, self.chunk_sizes[1])
        # 1. representation pass, no graph (same torch encoder + dropout streams as pass 3, not the CUDA
        #    inference encoder)
        core = model.module if hasattr(model, "module") else model
        core.force_torch_path = True
        try:
            q_reps, q_rng = self._encode_chunks(model, q_chunks, with_grad=False)
            p_reps, p_rng = self._encode_chunks(model, p_chunks, with_grad=False)
        finally:
            core.force_torch_path = False
        # 2. loss + gradient cache w.r.t. the representations
        q_all = torch.cat(q_reps).float().requires_grad_()
        p_all = torch.cat(p_reps).float().requires_grad_()
        loss = self.loss_fn(q_all, p_all)
        if self.args.gradient_accumulation_steps > 1:
            loss = loss / self.args.gradient_accumulation_steps
        loss.backward()
        q_cache = q_all.grad.split(self.chunk_sizes[0])
        p_cache = p_all.grad.split(self.chunk_sizes[1])
        # 3. chunked forward-backward against the cached gradients (DDP all-reduces only on the last chunk)
        todo = [(c, g, q_rng[i]) for i, (c, g) in enumerate(zip(q_chunks, q_cache))] + \
               [(c, g, p_rng[i]) for i, (c, g) in enumerate(zip(p_chunks, p_cache))]
        saved = torch.cuda.get_rng_state()
        for j, (chunk, grad, state) in enumerate(todo):
            last = j == len(todo) - 1
            sync_ctx = nullcontext() if (last or not hasattr(model, "no_sync")) else model.no_sync()
            with sync_ctx:
                torch.cuda.set_rng_state(state)
                with self._autocast():
                    reps = get_dense_rep(model(**chunk))
                surrogate = (reps.float() * grad).sum()
                if self._scaler is not None:
                    self._scaler.scale(surrogate).backward()
                else:
                    surrogate.backward()
        torch.cuda.set_rng_state(saved)
        return loss.detach() / self._dist_loss_scale_factor
