"""GraphedTrainStep — single-process spelling of `train_step.B200TrainStep` kept for callers of the round-1 name:
`GraphedTrainStep(net, loss_fn, opt, img0, lab0)` == `B200TrainStep(net, loss_fn, opt, (img0, lab0), schedule='graph')`."""
from __future__ import annotations

from typing import Callable

import torch

from .optim import B200AdamW
from .train_step import B200TrainStep


class GraphedTrainStep(B200TrainStep):
    def __init__(self, net: torch.nn.Module, loss_fn: Callable, optimizer: B200AdamW, example_img: torch.Tensor,
                 example_lab: torch.Tensor, warmup: int = 3):
        super().__init__(net, loss_fn, optimizer, (example_img, example_lab), schedule="graph", warmup=warmup)
