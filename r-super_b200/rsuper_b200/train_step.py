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


def _production_order(names):
    """Parameter names in the order B200UNet's backward produces their gradients (block='BasicBlock': head, up4 .. up1,
    down4 .. down1, inc, stem) and the four buckets the engine reports (unet._Engine._stage_done): {head, up4} 0.3 M parameters,
    {up3 .. up1} 17 M, {down4, down3} 20 M, {down2, down1, inc, stem} 2.3 M — the two large buckets leave while the encoder /
    the full-resolution layers are still in backward, only the last 9 MB are exposed (measured at N = 8: profiles/r02_scale_*).
    conv1 / shortcut of a block are adjacent (their gradients come out of ONE merged weight-gradient GEMM).  Any other
    parameter set: given order, no buckets."""
    have = set(names)

    def block(pre, shortcut):
        out = [pre + "conv2.conv.weight", pre + "conv1.conv.weight"]
        if shortcut:
            out.append(pre + "shortcut.conv.weight")
        return out

    stages = [["outc.weight", "outc.bias"] + block("up4.conv.1.", False) + block("up4.conv.0.", True)]
    s1 = []
    for j in (3, 2, 1):
        s1 += block(f"up{j}.conv.1.", False) + block(f"up{j}.conv.0.", True)
    stages.append(s1)
    s2, s3 = [], []
    for l in (4, 3):
        s2 += block(f"down{l}.conv.2.", False) + block(f"down{l}.conv.1.", True)
    for l in (2, 1):
        s3 += block(f"down{l}.conv.2.", False) + block(f"down{l}.conv.1.", True)
    stages.append(s2)
    stages.append(s3 + block("inc.conv2.", False) + ["inc.conv1.weight"])
    order = [n for st in stages for n in st]
    if set(order) != have or len(order) != len(names):
        return list(names), None
    return order, stages


class _FlatGradSink:
    """What unet._Engine writes weight gradients into when B200TrainStep owns the gradients (see _Engine._dw / _stage_done)."""

    def __init__(self, step, by_name, offs, bounds):
        self.step, self.by_name, self.offs, self.bounds = step, by_name, offs, bounds
        self.works = []
        self.done = set()
        self.comm = None

    def _view(self, name):
        p = self.by_name.get(name)
        if p is None or p.grad is None:
            return None
        o = self.offs[name]
        v = self.step.flat_grad[o:o + p.numel()]
        return v if p.grad.data_ptr() == v.data_ptr() else None     # the caller may have replaced / dropped p.grad

    def owns(self, name) -> bool:
        return self._view(name) is not None

    def covers_all(self) -> bool:
        return all(self.owns(n) for n in self.by_name)

    def buffer(self, name):
        if isinstance(name, tuple):                                   # (conv1, shortcut): one [2c, cin, 3, 3, 3] tensor
            a, b = self._view(name[0]), self._view(name[1])
            if a is None or b is None or a.data_ptr() + a.numel() * 4 != b.data_ptr():
                return None
            pa = self.by_name[name[0]]
            o = self.offs[name[0]]
            return self.step.flat_grad[o:o + a.numel() + b.numel()].view(2 * pa.shape[0], *pa.shape[1:])
        v = self._view(name)
        return None if v is None else v.view_as(self.by_name[name])

    def begin(self):
        self.works, self.done = [], set()

    def stage_done(self, k, side_stream):
        """Bucket k of the flat buffer is fully enqueued (main stream + the weight-gradient side stream): start its all-reduce
        on the communication stream while backward goes on."""
        st = self.step
        if st.world == 1 or not self.covers_all():
            return
        lo, hi = self.bounds[k]
        bucket = st.flat_grad[lo:hi]
        if bucket.is_cuda:
            if self.comm is None:
                self.comm = torch.cuda.Stream(device=bucket.device)
            cur = torch.cuda.current_stream()
            self.comm.wait_stream(cur)
            if side_stream is not None:
                self.comm.wait_stream(side_stream)
            with torch.cuda.stream(self.comm):
                w = st._allreduce_range(bucket, async_op=True)
            self.works.append((w, bucket))
        else:
            st._allreduce_range(bucket)
        self.done.add(k)

    def flush(self) -> bool:
        """Join the in-flight buckets into the current stream; True when they covered the whole buffer."""
        for w, bucket in self.works:
            if w is not None:
                w.wait()
            if bucket.is_cuda:
                torch.cuda.current_stream().wait_stream(self.comm)
        ok = len(self.done) == len(self.bounds)
        self.works = []
        return ok


class _LossRead:
    __slots__ = ("buf", "ev")

    def __init__(self, buf, ev):
        self.buf, self.ev = buf, ev

    def get(self) -> float:
        self.ev.synchronize()
        return B200TrainStep.check(float(self.buf))


class B200TrainStep:
    def __init__(self, net: torch.nn.Module, loss_fn: Callable, optimizer: B200AdamW, example_inputs: Sequence[Optional[torch.Tensor]],
                 *, schedule: str = "graph", process_group=None, side_stream: Optional[bool] = None, warmup: int = 3):
        if not isinstance(optimizer, B200AdamW) or not optimizer.capturable:
            raise ValueError("B200TrainStep needs a B200AdamW(capturable=True) optimizer")
        if schedule not in ("graph", "split", "eager"):
            raise ValueError("schedule must be 'graph', 'split' or 'eager'")
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
        self.side_stream = (schedule in ("graph", "split")) if side_stream is None else bool(side_stream)
        self.static = [None if t is None else t.clone() for t in example_inputs]
        self.loss = torch.zeros((), dtype=torch.float32, device=img.device)
        self._make_flat_grads()
        self.graph = None
        self.launches_per_step = None
        self.warmup_steps = 0
        self.graph_fwd = self.graph_bwd = None
        if schedule == "graph":
            self._capture(max(1, warmup))
        elif schedule == "split":
            self._capture_split(max(1, warmup))

    # -- one flat gradient buffer; every p.grad is a 16-byte-aligned view of it ---------------------------------------------
    def _make_flat_grads(self):
        named = [(n, p) for n, p in self.net.named_parameters() if p.requires_grad]
        order, stages = _production_order([n for n, _ in named])
        by_name = dict(named)
        offs, total = {}, 0
        for n in order:
            offs[n] = total
            total += (by_name[n].numel() + 3) // 4 * 4
        dev = named[0][1].device
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for n in order:
            p = by_name[n]
            p.grad = self.flat_grad[offs[n]:offs[n] + p.numel()].view_as(p)
        self.params = [by_name[n] for n in order]
        self.sink = None
        if stages is not None:
            # stage k = [lo, hi) of the flat buffer, complete when the engine reports stage_done(k)
            bounds = []
            for names in stages:
                lo = offs[names[0]]
                hi = offs[names[-1]] + (by_name[names[-1]].numel() + 3) // 4 * 4
                bounds.append((lo, hi))
            self.sink = _FlatGradSink(self, by_name, offs, bounds)
            self.net.__dict__["_grad_sink"] = self.sink

    def _allreduce(self):
        """DDP's gradient averaging (train_ddp.py:661-671) on the flat buffer: the buckets the engine reported during backward
        are already in flight on the communication stream (the main stream joins them here); whatever is not covered by
        buckets goes in one all-reduce of the whole buffer."""
        if self.world == 1:
            return
        if self.sink is not None and self.sink.flush():
            return
        self._allreduce_range(self.flat_grad)

    def _allreduce_range(self, t, async_op=False):
        import torch.distributed as dist
        if dist.get_backend(self.pg) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=False)   # gloo (CPU tests of the host logic): no AVG
        t.mul_(1.0 / self.world)
        return w if async_op else None

    def _body(self):
        from . import medformer as mf_mod
        from . import unet as unet_mod
        prev = unet_mod.set_side_stream(self.side_stream)
        # B200MedFormer's Functions may run their weight gradients on a second stream only here: the gradients land in the flat
        # buffer's views (zeroed below, on the main stream, before the forward pass) and the stream is joined before the
        # all-reduce / optimizer read them
        prev_mf = mf_mod.set_side_stream(self.side_stream and self.flat_grad.is_cuda)
        try:
            if self.sink is None or not self.sink.covers_all():
                self.flat_grad.zero_()             # p.grad are views: autograd accumulates in place
            if self.sink is not None:
                self.sink.begin()
            out = self.net(self.static[0])
            loss = self.loss_fn(out, *self.static[1:])
            loss.backward()
            mf_mod.join_side_stream()
            self._allreduce()
            self.opt.step()
            self.loss.copy_(loss.detach())
        finally:
            unet_mod.set_side_stream(prev)
            mf_mod.set_side_stream(prev_mf)
            mf_mod.join_side_stream()

    def _capture_stream(self):
        import os
        if getattr(self, "_cap_stream", None) is None:
            prio = -1 if os.environ.get("RSB_CAPTURE_PRIORITY", "0") == "1" else 0
            self._cap_stream = torch.cuda.Stream(device=self.loss.device, priority=prio)
        return self._cap_stream

    def _eager(self):
        self.opt.prepare_step()
        self._body()

    # -- schedule='split': the network in two graphs, the loss launch by launch in between ------------------------------------
    # For losses with host control flow (the report-supervised Volume / Ball losses read a few scalars back per tumour, like
    # the reference): graph A = UNet forward, graph B = UNet backward + gradient all-reduce + optimizer; only
    # calculate_loss and its backward to d(logits) run eagerly.  The engine is driven directly (same kernels, same order as
    # through autograd); the saved activations live in graph A's memory pool, which graph B shares.
    def _engine(self):
        from .unet import _Engine
        net = self.net
        dtype = torch.bfloat16 if net.precision == "bf16" else torch.float32
        eng = net.__dict__.get("_engine")
        if eng is None or eng.dtype != dtype or eng.slope != float(net.negative_slope):
            eng = _Engine(net.base_ch, net.negative_slope, dtype, net.block, getattr(net, "up_mode", "trilinear"))
            net.__dict__["_engine"] = eng
        eng.sink = net.__dict__.get("_grad_sink") if (net.block == "BasicBlock" and eng.up_mode == "trilinear") else None
        return eng

    def _split_fwd(self):
        from . import unet as unet_mod
        prev = unet_mod.set_side_stream(self.side_stream)
        try:
            eng = self._engine()
            self._P = {k: v.detach() for k, v in self.net.named_parameters()}
            with torch.no_grad():
                self._logits, self._saved = eng.forward(self.static[0].float().contiguous(), self._P, self.net.num_classes, save=True)
        finally:
            unet_mod.set_side_stream(prev)

    def _split_loss(self):
        lg = self._logits.detach().requires_grad_(True)          # a leaf that shares the logits' storage
        out = {"segmentation": lg} if self.net.return_dict else lg
        loss = self.loss_fn(out, *self.static[1:])
        loss.backward()
        self._dlogits.copy_(lg.grad)
        self.loss.copy_(loss.detach())

    def _split_bwd(self):
        from . import unet as unet_mod
        prev = unet_mod.set_side_stream(self.side_stream)
        try:
            eng = self._engine()
            if self.sink is None or not self.sink.covers_all():
                self.flat_grad.zero_()
            if self.sink is not None:
                self.sink.begin()
            with torch.no_grad():
                G = eng.backward(self._saved, self._P, self._dlogits)
                for name, p in self.net.named_parameters():
                    if p.requires_grad and not (eng.sink is not None and eng.sink.owns(name)):
                        p.grad.add_(G[name])
            self._allreduce()
            self.opt.step()
        finally:
            unet_mod.set_side_stream(prev)

    def _eager_split(self):
        self.opt.prepare_step()
        self._split_fwd()
        self._split_loss()
        self._split_bwd()

    def _capture_split(self, warmup: int):
        from . import ops
        dev = self.loss.device
        img = self.static[0]
        self._dlogits = torch.zeros((img.shape[0], self.net.num_classes) + tuple(img.shape[2:]), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_split()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        self.warmup_steps = warmup
        before = ops.LAUNCHES
        self.graph_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fwd, stream=self._capture_stream(), capture_error_mode="thread_local"):
            self._split_fwd()
        n_fwd = ops.LAUNCHES - before
        self.graph_fwd.replay()                  # real logits / activations for the loss that seeds the second capture
        before = ops.LAUNCHES
        self._split_loss()
        self._loss_launches = ops.LAUNCHES - before
        self.opt.prepare_step()
        before = ops.LAUNCHES
        self.graph_bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_bwd, pool=self.graph_fwd.pool(), stream=self._capture_stream(), capture_error_mode="thread_local"):
            self._split_bwd()
        self.launches_per_step = n_fwd + self._loss_launches + (ops.LAUNCHES - before)
        self.opt._prepared = False
        self.opt.global_step -= 1
        for p in self.params:
            if p in self.opt.state and len(self.opt.state[p]):
                self.opt.state[p]["step"] -= 1

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
        # thread_local: other threads (NCCL watchdog, data loader pinning) may call CUDA while this thread captures.
        # The capture stream has a HIGHER priority than the engine's second stream (default priority): when a data gradient
        # (main chain) and a weight gradient (second stream) become ready together, the data gradient goes first and the
        # weight gradient then runs beside the HBM-bound InstanceNorm-backward pass that follows it.
        with torch.cuda.graph(self.graph, stream=self._capture_stream(), capture_error_mode="thread_local"):
            self._body()
        self.launches_per_step = ops.LAUNCHES - before   # kernels of ours inside one replay (bench.py's gpu_launches)
        # the capture itself executed nothing, but step() advanced the host counters: they describe the NEXT replay already
        self.opt._prepared = False
        self.opt.global_step -= 1
        for p in self.params:
            if p in self.opt.state and len(self.opt.state[p]):
                self.opt.state[p]["step"] -= 1

    # -- input prefetch: the next batch's host -> device copy runs on a copy stream beside the current step ----------------------
    def prefetch(self, *inputs) -> None:
        """Start the asynchronous upload of the NEXT batch (pinned host tensors, or device tensors) into staging buffers on a
        dedicated copy stream; the following `step()` call without arguments consumes it.  What a prefetching data loader does
        (the reference: DataLoader(pin_memory=True) + `.cuda(non_blocking=True)`, train_ddp.py:311-318): the copy engine works
        while the SMs run the current step, instead of in front of the next one."""
        dev = self.loss.device
        if getattr(self, "_stage", None) is None:
            self._stage = [None if t is None else torch.empty_like(t) for t in self.static]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_done = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._stage_free)            # the previous batch has left the staging buffers
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage, inputs):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self._copy_done.record(self._copy_stream)
        self._prefetched = True

    def loss_async(self):
        """Asynchronous read-back of the loss of the step that was just enqueued: a 4-byte copy into pinned host memory + an
        event.  `.get()` on the returned handle waits for THAT copy only, so the caller can enqueue step i + 1 before it
        looks at loss i (`loss.item()` would drain the stream, and the GPU would idle while the host prepares the next step)."""
        if getattr(self, "_rd", None) is None:
            self._rd = [(torch.empty((), dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._rd_i = -1
        self._rd_i = (self._rd_i + 1) % len(self._rd)
        buf, ev = self._rd[self._rd_i]
        buf.copy_(self.loss, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.loss.device))
        return _LossRead(buf, ev)

    def __call__(self, *inputs) -> torch.Tensor:
        """One train step.  inputs mirror example_inputs; pinned host tensors are copied asynchronously into the static inputs.
        Called without arguments after `prefetch(...)`: the staged batch is moved into the static inputs (device to device)."""
        if not inputs:
            if not getattr(self, "_prefetched", False):
                raise RuntimeError("B200TrainStep(): call prefetch(...) first, or pass the inputs")
            cur = torch.cuda.current_stream(self.loss.device)
            cur.wait_event(self._copy_done)
            for dst, src in zip(self.static, self._stage):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self._stage_free.record(cur)
            self._prefetched = False
        for dst, src in zip(self.static, inputs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        if self.graph_fwd is not None:
            self.opt.prepare_step()
            self.graph_fwd.replay()
            self._split_loss()
            self.graph_bwd.replay()
            self.opt.finish_step()
            return self.loss
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
