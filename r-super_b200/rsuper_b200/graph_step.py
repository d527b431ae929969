"""GraphedTrainStep — the reference's train step (train_ddp.py:310-357) captured once into a CUDA graph and replayed.

Why: the step issues ~300 kernel launches; the host needs ~16 ms to enqueue them against ~22 ms of GPU time, so on hosts with
slow or shared cores the eager step is enqueue-bound and the two-stream schedule loses its lead.  Everything
on the path is capturable by construction: every launch goes to the current stream, no kernel allocates or synchronises, TMA
descriptors are encoded on the host at launch time with pointers that are static inside the graph's private pool, and the
optimizer's step-dependent scalars live in device memory (`B200AdamW(capturable=True)`), rewritten before each replay.

    step = GraphedTrainStep(net, lambda out, lab: lf.calculate_loss(out, lab, None, args, ...)['overall'], opt, img0, lab0)
    for img, lab in loader:
        loss = step(img, lab)              # H2D into the static inputs, one graph launch; loss is a device scalar

Restrictions (checked): single process (DDP's bucket hooks are not captured here), a `B200AdamW(capturable=True)` optimizer,
fixed input shapes, no host-side control flow on device values inside the loss (mask-only batches; the NaN check of
losses_foundation.py:1070 moves to the caller: `loss.item()` is NaN-checked by `GraphedTrainStep.check`).  STAGED: written
without GPU time left in round 1 — `tools/gpu_round2_first.sh` / `bench.py --cuda-graph` are its first run.
"""
from __future__ import annotations

from typing import Callable

import torch

from .optim import B200AdamW


class GraphedTrainStep:
    def __init__(self, net: torch.nn.Module, loss_fn: Callable, optimizer: B200AdamW, example_img: torch.Tensor,
                 example_lab: torch.Tensor, warmup: int = 3):
        if not isinstance(optimizer, B200AdamW) or not optimizer.capturable:
            raise ValueError("GraphedTrainStep needs a B200AdamW(capturable=True) optimizer")
        if not example_img.is_cuda:
            raise RuntimeError("GraphedTrainStep has no CPU path: inputs must live on a CUDA (sm_100a) device")
        if isinstance(net, torch.nn.parallel.DistributedDataParallel):
            raise NotImplementedError("GraphedTrainStep captures a single-process step (wrap the bare module)")
        self.net, self.loss_fn, self.opt = net, loss_fn, optimizer
        self.img = example_img.clone()
        self.lab = example_lab.clone()
        self.loss = torch.zeros((), dtype=torch.float32, device=example_img.device)
        # warm-up on a side stream (torch.cuda.graphs recipe).  At least one eager step is REQUIRED: it allocates everything that
        # must not be created during capture — packed-weight images and their job table, the gradients (their addresses go into
        # the optimizer's device table, uploaded from pinned memory), optimizer state, partials and the hyper-parameter block.
        # The captured body then starts with zero_grad(set_to_none=False): gradients keep those addresses for every replay.
        side = torch.cuda.Stream(device=example_img.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        self.warmup_steps = max(1, warmup)        # real optimizer steps already taken on (img, lab) = the example batch
        self.graph = torch.cuda.CUDAGraph()
        self.opt.prepare_step()
        from . import ops
        before = ops.LAUNCHES
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = ops.LAUNCHES - before   # kernels of ours inside one replay (bench.py's gpu_launches)
        # the capture itself executed nothing, but step() advanced the host counters: they describe the NEXT replay already
        self.opt._prepared = False
        self.opt.global_step -= 1
        for p in self.net.parameters():
            if p in self.opt.state and len(self.opt.state[p]):
                self.opt.state[p]["step"] -= 1

    def _body(self):
        self.opt.zero_grad(set_to_none=False)
        out = self.net(self.img)
        loss = self.loss_fn(out, self.lab)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def _eager(self):
        self.opt.prepare_step()
        self._body()

    def __call__(self, img: torch.Tensor, lab: torch.Tensor) -> torch.Tensor:
        """One train step.  img / lab may be pinned host tensors (async H2D into the static inputs) or device tensors."""
        self.img.copy_(img, non_blocking=True)
        self.lab.copy_(lab, non_blocking=True)
        self.opt.prepare_step()
        self.graph.replay()
        self.opt.finish_step()
        return self.loss

    @staticmethod
    def check(loss_value: float) -> float:
        if loss_value != loss_value:      # losses_foundation.py:1070-1071
            raise ValueError("loss is nan, propagating this can destroy the network weights, STOP!")
        return loss_value
