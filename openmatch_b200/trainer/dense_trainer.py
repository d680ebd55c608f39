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
            while len(pending) < group:
                pending.append(pending[len(pending) % max(1, len(pending))])
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


class GCDenseTrainer(DRTrainer):
    """Gradient-cache training depends on the un-vendored ``grad_cache`` package (reference
    dense_trainer.py:20-24,130-160); it is outside the current hot-path scope."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("GradCache training needs the external `grad_cache` package and is not part of the "
                                  "B200 hot path yet; train with DRTrainer (--grad_cache False)")
