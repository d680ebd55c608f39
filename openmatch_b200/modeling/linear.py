"""Bias-free projection head, checkpoint-compatible with the reference (``src/openmatch/modeling/linear.py``):
``linear.pt`` holds the ``state_dict`` (key ``linear.weight``), ``head_config.json`` the two dimensions."""
import json
import logging
import os

import torch
import torch.nn as nn
from torch import Tensor

logger = logging.getLogger(__name__)


class LinearHead(nn.Module):
    def __init__(self, input_dim: int = 768, output_dim: int = 768):
        super().__init__()
        self.linear = nn.Linear(input_dim, output_dim, bias=False)
        self.config = {"input_dim": input_dim, "output_dim": output_dim}

    def forward(self, rep: Tensor = None):
        return self.linear(rep)

    @classmethod
    def load(cls, ckpt_dir: str):
        logger.info("Loading linear head from %s", ckpt_dir)
        with open(os.path.join(ckpt_dir, "head_config.json")) as f:
            head = cls(**json.load(f))
        head.load_state_dict(torch.load(os.path.join(ckpt_dir, "linear.pt"), map_location="cpu"))
        return head

    def save(self, save_path: str):
        torch.save(self.state_dict(), os.path.join(save_path, "linear.pt"))
        with open(os.path.join(save_path, "head_config.json"), "w") as f:
            json.dump(self.config, f, indent=4)
