"""B200AdamW — the optimizer end of the reference train step as two multi-tensor kernel launches.

Mirrors what `train_epoch` does between `loss.backward()` and the next iteration
(rsuper_train/train_ddp.py:352-357):

    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)          # train_ddp.py:352
    optimizer.step()                                                # AdamW(eps=1e-5), training/utils.py:46-51
    update_ema_variables(net, ema_net, args.ema_alpha, step)        # training/utils.py:154-158

`B200AdamW` is a `torch.optim.Optimizer` (param_groups / state_dict / zero_grad / LR schedulers such as
`exp_lr_scheduler_with_warmup` keep working; the per-parameter state uses torch's AdamW keys `step`, `exp_avg`,
`exp_avg_sq`, so checkpoints are interchangeable with `training.utils.get_optimizer`'s AdamW), whose `step()` runs
`rsb_clip_adamw_ema_step` (csrc/train_glue.cu): gradient norm -> clip -> AdamW -> EMA in one pass over
(g, p, exp_avg, exp_avg_sq, ema).  No CPU / PyTorch fallback: CPU parameters raise.

Usage in place of the three reference lines above:

    opt = B200AdamW(net.parameters(), lr=args.base_lr, betas=args.betas, weight_decay=args.weight_decay,
                    max_norm=1.0, ema_params=list(ema_net.parameters()), ema_alpha=args.ema_alpha)
    ...
    loss.backward(); opt.step()          # opt.last_grad_norm = what clip_grad_norm_ would have returned
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Iterable, Optional, Sequence

import torch

from . import ops


class B200AdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-5, weight_decay: float = 1e-2,
                 max_norm: Optional[float] = None, ema_params: Optional[Sequence[torch.Tensor]] = None, ema_alpha: float = 0.99,
                 capturable: bool = False):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError(f"B200AdamW: invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_norm = max_norm
        self.ema_alpha = ema_alpha
        self.ema_params = list(ema_params) if ema_params is not None else None
        flat = [p for g in self.param_groups for p in g["params"]]
        if self.ema_params is not None and (len(self.ema_params) != len(flat) or
                                            any(e.shape != p.shape for e, p in zip(self.ema_params, flat))):
            raise ValueError("B200AdamW: ema_params must mirror the parameters one to one")
        self.global_step = 0          # update_ema_variables' global_step (train_ddp.py:308)
        self.last_grad_norm = None    # device scalar: total gradient norm before clipping
        self._tables = {}             # group index -> (pointer signature, device table, n_tensors, total_chunks)
        self._partials = None
        # capturable: the step-dependent scalars live in a device block that `prepare_step()` rewrites before every launch
        # (or CUDA-graph replay); `step()` then only launches — nothing step-dependent is baked into the captured kernels
        self.capturable = capturable
        self._hyper_dev = None
        self._prepared = False

    # -- device table of one param group --------------------------------------------------------------------------
    def _table(self, gi: int, group, ema_of):
        chunk = ops.lib().rsb_opt_chunk_elems()
        rows, sig = [], []
        begin = 0
        for p in group["params"]:
            if p.numel() == 0:
                continue
            if p.grad is None:
                # torch's AdamW skips a parameter without a gradient; update_ema_variables (training/utils.py:154-158) still
                # moves its EMA copy: an EMA-only row (g = NULL)
                e = ema_of.get(id(p))
                if e is not None:
                    if not (ops._on_device(p) and ops._on_device(e) and e.dtype == torch.float32 and e.is_contiguous()
                            and p.dtype == torch.float32 and p.is_contiguous()):
                        raise RuntimeError("B200AdamW: parameters / EMA tensors must be contiguous CUDA fp32")
                    rows.append((p.data_ptr(), 0, 0, 0, e.data_ptr(), p.numel(), begin))
                    sig.append(rows[-1][:5])
                    begin += (p.numel() + chunk - 1) // chunk
                continue
            if not ops._on_device(p):
                raise RuntimeError("B200AdamW has no CPU path: parameters must live on a CUDA (sm_100a) device")
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous() \
                    or p.grad.is_sparse:
                raise RuntimeError("B200AdamW: parameters and gradients must be dense contiguous fp32")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)  # host scalar tensor like torch's non-capturable AdamW
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            e = ema_of.get(id(p))
            if e is not None and not (ops._on_device(e) and e.dtype == torch.float32 and e.is_contiguous()):
                raise RuntimeError("B200AdamW: EMA tensors must be contiguous CUDA fp32")
            n = p.numel()
            rows.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         e.data_ptr() if e is not None else 0, n, begin))
            sig.append(rows[-1][:5])
            begin += (n + chunk - 1) // chunk
        if not any(r[1] for r in rows):
            return None
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == sig:
            return cached
        # 7 x int64 per row == struct RsbOptTensor (5 pointers, n, chunk_begin); uploaded only when a pointer changed
        dev = group["params"][0].device
        host = torch.tensor(rows, dtype=torch.int64)
        if dev.type == "cuda":
            host = host.pin_memory()
        table = host.to(dev, non_blocking=True)
        cached = (sig, table, len(rows), begin, host)   # keep the pinned source alive until the copy has run
        self._tables[gi] = cached
        return cached

    def _step_scalars(self, group):
        """(adam step t, EMA alpha) of the NEXT update of a group (t = torch's per-parameter `step` + 1; training/utils.py:156)."""
        def count(p):
            return int(self.state[p]["step"].item()) if (p in self.state and len(self.state[p])) else 0
        with_grad = [p for p in group["params"] if p.grad is not None and p.numel()]
        if with_grad:
            # the parameters this launch updates share one pair of bias corrections (a parameter that sat out some steps
            # and returns with a smaller count would need its own: rejected instead of silently mis-corrected)
            steps = {count(p) for p in with_grad}
            if len(steps) > 1:
                raise RuntimeError("B200AdamW: the parameters that receive a gradient in one step must share their step count "
                                   f"(found {sorted(steps)}): one launch carries one pair of bias corrections")
        else:
            # prepare_step() before backward (zero_grad(set_to_none=True) left no gradients yet): the count of the parameters
            # that have been stepping, i.e. the largest one
            steps = {max([count(p) for p in group["params"] if p.numel()] or [0])}
        t = (steps.pop() if steps else 0) + 1
        return t, min(1.0 - 1.0 / (self.global_step + 1), self.ema_alpha)

    @torch.no_grad()
    def prepare_step(self, global_step: Optional[int] = None):
        """capturable mode: upload the scalars of the next update (bias corrections for step t, the group's current lr,
        the EMA warm-up alpha) into the device block the kernel reads.  Call once before every `step()` / graph replay;
        after a replay call `finish_step()` so that the host-side counters follow."""
        if not self.capturable:
            raise RuntimeError("prepare_step() is for B200AdamW(capturable=True)")
        if global_step is not None:
            self.global_step = int(global_step)
        if len(self.param_groups) != 1:
            raise NotImplementedError("B200AdamW(capturable=True) supports one param group")
        group = self.param_groups[0]
        dev = group["params"][0].device
        n = ops.lib().rsb_opt_hyper_floats()
        t, alpha = self._step_scalars(group)
        host = torch.empty(n, dtype=torch.float32)
        b1, b2 = group["betas"]
        ops.check(ops.lib().rsb_opt_fill_hyper(C.c_void_p(host.data_ptr()), float(self.max_norm) if self.max_norm is not None else 0.0,
                                               float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                               t, float(alpha)), "opt_fill_hyper")
        if self._hyper_dev is None:
            self._hyper_dev = torch.empty(n, dtype=torch.float32, device=dev)
        self._hyper_dev.copy_(host)           # pageable source: the staging copy is synchronous for the host, stream-ordered on the device
        self._prepared = True

    def finish_step(self):
        """capturable mode, after a CUDA-graph replay of a captured `step()`: advance the host-side counters."""
        for group in self.param_groups:
            for p in group["params"]:
                if p in self.state and len(self.state[p]):
                    self.state[p]["step"] += 1
        self.global_step += 1
        self._prepared = False

    # -- checkpoints: update_ema_variables' global_step (train_ddp.py:308: i + epoch * len(loader)) travels with the state ----
    def state_dict(self):
        sd = super().state_dict()
        sd["b200_global_step"] = int(self.global_step)   # extra top-level key: torch.optim.AdamW.load_state_dict ignores it
        return sd

    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        gs = sd.pop("b200_global_step", None)
        super().load_state_dict(sd)
        if gs is None:
            # a checkpoint written by the reference's torch AdamW: the loop step equals the number of updates taken
            steps = [int(st["step"]) for st in self.state.values() if "step" in st]
            gs = max(steps) if steps else 0
        self.global_step = int(gs)
        for st in self.state.values():                   # host scalar steps like torch's non-capturable AdamW
            if "step" in st and isinstance(st["step"], torch.Tensor) and st["step"].device.type != "cpu":
                st["step"] = st["step"].detach().cpu()
        self._tables.clear()
        self._prepared = False

    @torch.no_grad()
    def step(self, closure=None, global_step: Optional[int] = None):
        """global_step: the loop step the reference hands to update_ema_variables (train_ddp.py:357); default = the
        optimizer's own counter (number of step() calls, persisted in state_dict())."""
        if global_step is not None:
            if self.capturable and self._prepared and int(global_step) != self.global_step:
                raise RuntimeError("B200AdamW(capturable=True): pass global_step to prepare_step(), not step()")
            self.global_step = int(global_step)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ema_of = {}
        if self.ema_params is not None:
            flat = [p for g in self.param_groups for p in g["params"]]
            ema_of = {id(p): e for p, e in zip(flat, self.ema_params)}
        if self.max_norm is not None and len(self.param_groups) != 1:
            raise NotImplementedError("B200AdamW: gradient clipping spans one param group (training/utils.py:29-33 builds one)")
        for gi, group in enumerate(self.param_groups):
            tab = self._table(gi, group, ema_of)
            if tab is None:
                continue
            _, table, n_tensors, total_chunks, _ = tab
            dev = table.device
            t, alpha = self._step_scalars(group)
            if self.capturable and not self._prepared:
                raise RuntimeError("B200AdamW(capturable=True): call prepare_step() before step()")
            if self._partials is None or self._partials.device != dev:
                self._partials = torch.empty(ops.lib().rsb_opt_max_blocks() + 1, dtype=torch.float32, device=dev)
            norm_out = self._partials[-1:]
            clip = float(self.max_norm) if self.max_norm is not None else 0.0
            b1, b2 = group["betas"]
            with torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext():
                ops._call("train_glue", 2 if clip > 0 else 1, 0.0, ops.lib().rsb_clip_adamw_ema_step, C.c_void_p(table.data_ptr()),
                          n_tensors, total_chunks, int(bool(ema_of)), ops._p(self._partials) if clip > 0 else None, ops._p(norm_out),
                          clip, float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), t,
                          float(alpha), ops._p(self._hyper_dev) if self.capturable else None, ops._stream(), what="clip_adamw_ema_step")
            self.last_grad_norm = norm_out if clip > 0 else None
            for p in group["params"]:
                if p.grad is not None and p.numel():
                    self.state[p]["step"] += 1
        self.global_step += 1
        self._prepared = False
        return loss
