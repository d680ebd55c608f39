"""``python -m openmatch.driver.train_dr``: contrastive training of a dense retriever
(reference: ``src/openmatch/driver/train_dr.py``)."""
import logging
import os
import random

import numpy as np
import torch

from ..arguments import DataArguments, DRTrainingArguments as TrainingArguments, ModelArguments
from ..dataset import DRTrainDataset, QPCollator
from ..modeling import DRModel
from ..trainer import DRTrainer, GCDenseTrainer
from ._common import load_config, load_tokenizer, parse, setup_logging

logger = logging.getLogger(__name__)


def set_seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def main():
    model_args, data_args, training_args = parse((ModelArguments, DataArguments, TrainingArguments))
    out = training_args.output_dir
    if out and os.path.exists(out) and os.listdir(out) and training_args.do_train and not training_args.overwrite_output_dir:
        raise ValueError(f"Output directory ({out}) already exists and is not empty. Use --overwrite_output_dir to overcome.")
    setup_logging(training_args, logger)
    logger.info("Training/evaluation parameters %s", training_args)
    logger.info("MODEL parameters %s", model_args)
    set_seed(training_args.seed)
    config = load_config(model_args)
    tokenizer = load_tokenizer(model_args, use_fast=False)
    model = DRModel.build(model_args, data_args, training_args, config=config, cache_dir=model_args.cache_dir)
    train_dataset = DRTrainDataset(tokenizer, data_args, shuffle_seed=training_args.seed,
                                   cache_dir=data_args.data_cache_dir or model_args.cache_dir)
    trainer_cls = GCDenseTrainer if training_args.grad_cache else DRTrainer
    trainer = trainer_cls(model=model, args=training_args, tokenizer=tokenizer, train_dataset=train_dataset,
                          data_collator=QPCollator(tokenizer, max_p_len=data_args.p_max_len, max_q_len=data_args.q_max_len))
    train_dataset.trainer = trainer
    trainer.train()
    trainer.save_model()
    if trainer.is_world_process_zero():
        tokenizer.save_pretrained(training_args.output_dir)


if __name__ == "__main__":
    main()
