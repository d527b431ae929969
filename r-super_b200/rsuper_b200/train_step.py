"""B200TrainStep — the reference's train step (train_ddp.py:310-357) as ONE object: forward, calculate_loss, backward,
data-parallel gradient all-reduce (NCCL), clip + AdamW + EMA; captured once into a CUDA graph and replayed.

    step = B200TrainStep(net, loss_fn, opt, (img0, lab0), process_group=dist.group.WORLD)   # one per rank
    for img, lab in loader:
        loss = step(img, lab)              # H2D into the static inputs, one graph launch; loss is a device scalar

Why a graph: the step issues ~300 kernel launches; the host needs ~16 ms to enqueue them against ~21 ms of GPU time, so on
hosts with slow or shared cores (8 ranks on one box) the eager step is enqueue-bound, ranks straggle, and the two-stream
schedule (weight gradients beside the dgrad / InstanceNorm-backward chain, unet.set_side_stream) is only robust when the
kernel order is fixed by graph edges instead of by host timing.

What the reference does with DistributedDataParallel (train_ddp.py:661-671: bucketed all-reduce of fp32 gradients on NCCL)
is done here on ONE flat fp32 gradient buffer that every `p.grad` is a view of: the all-reduce is a single NCCL call on
that buffer (averaged, like DDP), captured inside the graph between backward and the fused optimizer.  Stock
`DistributedDataParallel(B200UNet(...))` keeps working — this object is the fast path.

Everything on the path is capturable by construction: every launch goes to the current stream, no kernel allocates or
synchronises, TMA descriptors are encoded on the host at capture time with pointers that are static inside the graph's
private pool, and the optimizer's step-dependent scalars live in device memory (`B200AdamW(capturable=True)`), rewritten
before each replay.  Restrictions (checked): `B200AdamW(capturable=True)`, fixed input shapes, no host-side control flow on
device values inside `loss_fn` for schedule='graph' (mask-only batches, or report batches through the device-side report
losses; the NaN check of losses_foundation.py:1070 moves to the caller: `B200TrainStep.check(loss.item())`).
schedule='eager' runs the same body launch by launch (any loss_fn).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from .optim import B200AdamW


class B200TrainStep:
    def __init__(self, net: torch.nn.Module, loss_fn: Callable, optimizer: B200AdamW, example_inputs: Sequence[Optional[torch.Tensor]],
                 *, schedule: str = "graph", process_group=None, side_stream: Optional[bool] = None, warmup: int = 3):
        if not isinstance(optimizer, B200AdamW) or not optimizer.capturable:
            raise ValueError("B200TrainStep needs a B200AdamW(capturable=True) optimizer")
        if schedule not in ("graph", "eager"):
            raise ValueError("schedule must be 'graph' or 'eager'")
        if isinstance(net, torch.nn.parallel.DistributedDataParallel):
            raise NotImplementedError("B200TrainStep does its own gradient all-reduce: pass the bare module and process_group=")
        from . import ops
        img = example_inputs[0]
        if not ops._on_device(img):
            raise RuntimeError("B200TrainStep has no CPU path: inputs must live on a CUDA (sm_100a) device")
        self.net, self.loss_fn, self.opt, self.schedule = net, loss_fn, optimizer, schedule
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.side_stream = (schedule == "graph") if side_stream is None else bool(side_stream)
        self.static = [None if t is None else t.clone() for t in example_inputs]
        self.loss = torch.zeros((), dtype=torch.float32, device=img.device)
        self._make_flat_grads()
        self.graph = None
        self.launches_per_step = None
        self.warmup_steps = 0
        if schedule == "graph":
            self._capture(max(1, warmup))

    # -- one flat gradient buffer; every p.grad is a 16-byte-aligned view of it ---------------------------------------------
    def _make_flat_grads(self):
        params = [p for p in self.net.parameters() if p.requires_grad]
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        dev = params[0].device
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.params = params

    def _allreduce(self):
        """DDP's gradient averaging (train_ddp.py:661-671) as one NCCL all-reduce of the flat buffer."""
        if self.world == 1:
            return
        import torch.distributed as dist
        if dist.get_backend(self.pg) == "nccl":
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG, group=self.pg)
        else:                                   # gloo (CPU tests of the host logic): no AVG
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
            self.flat_grad.mul_(1.0 / self.world)

    def _body(self):
        from . import unet as unet_mod
        prev = unet_mod.set_side_stream(self.side_stream)
        try:
            self.flat_grad.zero_()                 # p.grad are views: autograd accumulates in place
            out = self.net(self.static[0])
            loss = self.loss_fn(out, *self.static[1:])
            loss.backward()
            self._allreduce()
            self.opt.step()
            self.loss.copy_(loss.detach())
        finally:
            unet_mod.set_side_stream(prev)

    def _eager(self):
        self.opt.prepare_step()
        self._body()

    def _capture(self, warmup: int):
        from . import ops
        dev = self.loss.device
        # warm-up on a side stream (torch.cuda.graphs recipe).  At least one eager step is REQUIRED: it allocates everything
        # that must not be created during capture — packed-weight images and their job table, optimizer state and its device
        # table (uploaded from pinned memory), partials, the hyper-parameter block — and initialises the NCCL communicator.
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        self.warmup_steps = warmup                # real optimizer steps already taken on the example batch
        self.graph = torch.cuda.CUDAGraph()
        self.opt.prepare_step()
        before = ops.LAUNCHES
        # thread_local: other threads (NCCL watchdog, data loader pinning) may call CUDA while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._body()
        self.launches_per_step = ops.LAUNCHES - before   # kernels of ours inside one replay (bench.py's gpu_launches)
        # the capture itself executed nothing, but step() advanced the host counters: they describe the NEXT replay already
        self.opt._prepared = False
        self.opt.global_step -= 1
        for p in self.params:
            if p in self.opt.state and len(self.opt.state[p]):
                self.opt.state[p]["step"] -= 1

    def __call__(self, *inputs) -> torch.Tensor:
        """One train step.  inputs mirror example_inputs; pinned host tensors are copied asynchronously into the static inputs."""
        for dst, src in zip(self.static, inputs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        if self.graph is None:
            from . import ops
            before = ops.LAUNCHES
            self._eager()
            self.launches_per_step = ops.LAUNCHES - before
            return self.loss
        self.opt.prepare_step()
        self.graph.replay()
        self.opt.finish_step()
        return self.loss

    @staticmethod
    def check(loss_value: float) -> float:
        if loss_value != loss_value:      # losses_foundation.py:1070-1071
            raise ValueError("loss is nan, propagating this can destroy the network weights, STOP!")
        return loss_value
