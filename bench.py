#!/usr/bin/env python
"""bench.py — Mvoxels/s of the R-Super 3D segmentation TRAIN STEP on synthetic CT patches.

    python bench.py --gpus N --steps K --warmup W [--config cfg2|cfg3|cfg4|cfg5] [--precision bf16|fp32]
    python bench.py --impl reference --gpus N --steps K ...      # the reference's own modules on the host cores

One step = train_epoch's body (rsuper_train/train_ddp.py:310-357): zero_grad -> UNet forward -> calculate_loss -> backward
(data-parallel gradient all-reduce on NCCL for N > 1) -> clip_grad_norm_(1.0) -> AdamW(eps 1e-5, wd 0.05) -> EMA update,
through `rsuper_b200.train_step.B200TrainStep` (the whole step captured in a CUDA graph where the loss path has no host
control flow, launch by launch otherwise).

Workloads (BASELINE.json configs; reference UNet base 32, 5 levels, BasicBlock/IN/ReLU, batch 2 per GPU):
    cfg2 (default; configs[1])  2-class masked BCE + Dice on synthetic masks, 128^3
    cfg3 (configs[2])           + Volume / Ball loss on a mixed mask / report batch, 128^3
    cfg4 (configs[3])           PanTS-shaped: 3 classes, 96 x 192 x 192, mixed batch
    cfg5 (configs[4])           7 tumour channels + organ, 160^3, mixed batch

Prints ONE JSON line (rank 0).  `value` = whole-job Mvoxels/s with inputs resident in HBM (CUDA events around the K steps,
nothing else recorded, max over ranks); `e2e` = the same through the public API with pinned HOST buffers (H2D of every
input tensor — prefetched on a copy stream beside the previous step, B200TrainStep.prefetch — and D2H of the loss inside the timed region); `roofline` / `kernels` = a separate profiled pass of the same K
steps inside this script (every launch bracketed by CUDA events on its stream, one stream, so that a kernel's events time
that kernel alone): the dominant kernel (tcgen05 implicit-GEMM conv: fprop + dgrad launches) vs its algorithmic FLOPs and
the measured bf16 peak; `cpu_baseline` = the reference's own modules (oracle/_ref, staged by oracle/build_ref.py) timed on
this box's host cores on a bounded sample; `torch_gpu_baseline` = the same reference modules on this GPU with stock
PyTorch / cuDNN (fp32 as the reference trains, and bf16 autocast).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "r-super_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

METRIC = "Mvoxels/sec (train step) on 128^3 CT patches"
UNIT = "Mvoxels/s"
MFLOP_PER_VOXEL_STEP = 3.705  # 3 x 1.235 MFLOP/voxel (fprop + dgrad + wgrad), base-32 UNet
CLASSES = ["organ", "pancreatic_lesion"]

_TUMOURS = sorted(f"{o}_lesion" for o in ("adrenal", "bladder", "colon", "esophagus", "kidney", "liver", "spleen"))
CONFIGS = {
    "cfg2": dict(classes=CLASSES, shape=(128, 128, 128), kinds=("mask", "mask"), report=False, baseline="configs[1]",
                 loss="2-class masked BCE+Dice on synthetic masks"),
    "cfg3": dict(classes=CLASSES, shape=(128, 128, 128), kinds=("mask", "report"), report=True, baseline="configs[2]",
                 loss="2-class masked BCE+Dice + Volume/Ball loss on synthetic report targets, mixed mask/report batch"),
    "cfg4": dict(classes=["pancreas", "pancreatic_lesion", "veins"], shape=(96, 192, 192), kinds=("mask", "report"), report=True,
                 baseline="configs[3]", loss="PanTS-shaped 3-class BCE+Dice + Ball loss, mixed mask/report batch"),
    "cfg5": dict(classes=["organ"] + _TUMOURS, shape=(160, 160, 160), kinds=("report", "mask"), report=True, baseline="configs[4]",
                 loss="7-tumour + organ 8-class BCE+Dice + Ball loss, mixed mask/report batch"),
}


def workload_name(base: int, batch: int, size, cfg: str = "cfg2") -> str:
    c = CONFIGS[cfg]
    shape = (size,) * 3 if isinstance(size, int) else tuple(size)
    dims = f"{shape[0]}^3" if shape[0] == shape[1] == shape[2] else "x".join(str(s) for s in shape)
    return (f"reference UNet base{base} 5-level (BasicBlock/IN/ReLU) train step, batch {batch} x {dims} per GPU, "
            f"{c['loss']} (BASELINE.json {c['baseline']})")


def rank_seeds(rank: int):
    """(image seed, label seed, torch seed) of a rank: every rank trains on its own synthetic shard (weak scaling)."""
    return 1234 + rank, 4321 + rank, 1234 + rank


def job_voxels(world: int, batch: int, size) -> int:
    n = 1
    for s in ((size,) * 3 if isinstance(size, int) else size):
        n *= s
    return world * batch * n


def mvox_per_s(voxels: int, ms: float) -> float:
    return voxels / (ms * 1e-3) / 1e6


def profiled_traffic():
    """dram read + write bytes of ONE launch of the dominant kernel at the benchmark's dominant shape (32 -> 32 channels at
    2 x 128^3, plane-streaming kernel) from the committed ncu --set full capture (profiles/r02_kernels_ncu_summary.txt,
    written by tools/summarize_ncu.py roofline), or None."""
    path = os.path.join(ROOT, "profiles", "r02_kernels_ncu_summary.txt")
    if not os.path.exists(path):
        return None, None
    for ln in open(path):
        if "conv3_stream32_kernel" in ln:
            f = ln.split()
            # ... kernel name tokens | grid regs us dram_MB GB/s %HBM tensor% warps% | stalls
            for i, tok in enumerate(f):
                if tok == "148":
                    return float(f[i + 3]) * 1e6, os.path.basename(path)
    return None, os.path.basename(path)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# The reference's own implementation of the step (oracle/_ref: its unmodified modules, staged by oracle/build_ref.py; the
# oracle port when the staged copy is missing) — on the host cores (cpu_baseline, --impl reference) or on this GPU with
# stock PyTorch (torch_gpu_baseline).  This is the only place bench.py executes anything under oracle/.
# --------------------------------------------------------------------------------------------
def reference_step_factory(cfg: str, base: int, shape, kinds, device, autocast: bool = False):
    """One reference train step (order of train_ddp.py:310-357).  Returns (step() -> loss value, kind)."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle import build_ref
    c = CONFIGS[cfg]
    classes = c["classes"]
    kind = "port"
    os.chdir(tempfile.mkdtemp(prefix="rsb_ref_"))          # the reference's loss writes debug folders on its first calls
    largs = LR.default_args() if c["report"] else LR.default_args(report_volume_loss_basic=0.0)
    largs.model_genesis_pretrain = False
    batch = synth.make_batch(list(kinds), classes, tuple(shape), seed=12, device="cpu")
    if build_ref.available():
        unet_mod, lf, tu = build_ref.import_reference()
        from oracle.unet_ref import synthetic_state_dict
        net = unet_mod.UNet(1, base, num_classes=len(classes), scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5,
                            block="BasicBlock")
        net.load_state_dict(synthetic_state_dict(base, len(classes)))
        net = net.to(device)
        import copy
        ema_net = copy.deepcopy(net)
        from types import SimpleNamespace
        opt = tu.get_optimizer(SimpleNamespace(optimizer="adamw", momentum=0.9, base_lr=6e-4, betas=(0.9, 0.999), weight_decay=0.05), net)
        kind = "reference"
        forward = lambda x: {"segmentation": net(x)}
        loss_fn = lf.calculate_loss
        params = list(net.parameters())

        def ema_update(step):
            tu.update_ema_variables(net, ema_net, 0.99, step)
    else:
        from oracle.unet_ref import synthetic_state_dict, unet_forward
        sd = {k: torch.nn.Parameter(v.to(device)) for k, v in synthetic_state_dict(base, len(classes)).items()}
        params = list(sd.values())
        ema = [p.detach().clone() for p in params]
        opt = torch.optim.AdamW(params, lr=6e-4, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5)
        forward = lambda x: {"segmentation": unet_forward(x, sd)}
        loss_fn = LR.calculate_loss

        def ema_update(step):
            alpha = min(1 - 1 / (step + 1), 0.99)
            with torch.no_grad():
                for e, p in zip(ema, params):
                    e.mul_(alpha).add_(p.detach(), alpha=1 - alpha)
    x = batch["image"].to(device)
    if autocast:
        x = x.contiguous(memory_format=torch.channels_last_3d)
    lab = batch["label"].to(device).long()
    rep = c["report"]
    unk = batch["unk_channels"].to(device).float() if rep else None
    msk = batch["mask"].to(device).float() if rep else None
    vol = batch["volumes"].to(device) if rep else None
    dia = batch["diameters"].to(device) if rep else None
    state = {"step": 0}

    import contextlib
    import warnings

    def step():
        # the reference's loss prints / writes debug files on its first calls: keep stdout for the ONE JSON line
        with contextlib.redirect_stdout(sys.stderr), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            opt.zero_grad()
            with torch.autocast(torch.device(device).type, dtype=torch.bfloat16, enabled=autocast):
                out = forward(x)
            out = {"segmentation": out["segmentation"].float()}
            loss = loss_fn(model_output=out, label=lab, unk_voxels=unk, args=largs, matcher=None, chosen_segment_mask=msk,
                           tumor_volumes_report=vol, tumor_diameters=dia, classes=classes, input_tensor=x)
            loss["overall"].backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            ema_update(state["step"])
            state["step"] += 1
            return loss["overall"].item()

    return step, kind


def run_cpu_arm(args, as_reference_impl: bool):
    """The reference step on the host cores, on a bounded sample of the workload: the largest of (full batch, one sample,
    96^3 / 64^3 / 48^3 / 32^3 crops of one sample) whose warm-up + timed steps fit the time budget."""
    cfg = CONFIGS[args.config]
    torch.set_num_threads(os.cpu_count() or 1)
    nsteps = (args.steps + args.warmup) if as_reference_impl else 2
    budget = 220.0 if as_reference_impl else 22.0
    full = tuple(cfg["shape"])
    cal_shape = (32, 32, 32)
    step, kind = reference_step_factory(args.config, args.base, cal_shape, cfg["kinds"][:1], "cpu")
    step()
    t0 = time.perf_counter(); step(); per_vox = (time.perf_counter() - t0) / (32 ** 3)
    cands = [(len(cfg["kinds"]), full), (1, full)] + [(1, (s, s, s)) for s in (96, 64, 48, 32) if s < min(full)]
    nb, shp = cands[-1]
    for n_b, s in cands:
        if per_vox * n_b * s[0] * s[1] * s[2] * nsteps * 0.8 <= budget:   # larger patches run ~20 % more efficiently than 32^3
            nb, shp = n_b, s
            break
    step, kind = reference_step_factory(args.config, args.base, shp, cfg["kinds"][:nb], "cpu")
    warm = args.warmup if as_reference_impl else 1
    timed = args.steps if as_reference_impl else 1
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(timed):
        step()
    dt = (time.perf_counter() - t0) / timed
    vox = nb * shp[0] * shp[1] * shp[2]
    what = "the reference's own modules (oracle/_ref: UNet, calculate_loss, get_optimizer, update_ema_variables)" if kind == "reference" \
        else "oracle port of the reference step"
    whole = (nb == len(cfg["kinds"]) and shp == full)
    sample = (f"{what}, fp32, torch CPU, {torch.get_num_threads()} threads: "
              + ("the full per-GPU batch " if whole else "a bounded sample ")
              + f"{nb} x {'x'.join(map(str, shp))} of the {len(cfg['kinds'])} x {'x'.join(map(str, full))} workload, {warm} warm-up + {timed} timed steps")
    return dict(value=vox / dt / 1e6, unit=UNIT, cores=torch.get_num_threads(), kind=kind, sample=sample), dt * 1e3


def run_torch_gpu_baseline(args, dev):
    """The reference modules on THIS GPU with stock PyTorch / cuDNN: fp32 as the reference trains (`--amp` raises; torch's
    default cuDNN TF32 convolutions), and bf16 autocast + channels_last_3d (the most favourable stock setting)."""
    cfg = CONFIGS[args.config]
    out = {"unit": UNIT, "modes": {}}
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    cwd = os.getcwd()
    try:
        for name, autocast in (("fp32_tf32_convs (reference default)", False), ("bf16_autocast_channels_last", True)):
            torch.backends.cudnn.allow_tf32 = True
            step, kind = reference_step_factory(args.config, args.base, cfg["shape"], cfg["kinds"], dev, autocast=autocast)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 4
            e0.record()
            for _ in range(n):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out["modes"][name] = {"ms_per_step": ms, "value": mvox_per_s(job_voxels(1, len(cfg["kinds"]), cfg["shape"]), ms)}
            out["kind"] = kind
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev
        os.chdir(cwd)
    best = max(out["modes"].values(), key=lambda m: m["value"])
    out["value"] = best["value"]
    out["what"] = ("the reference's own UNet / calculate_loss / get_optimizer / update_ema_variables (oracle/_ref)" if out.get("kind") == "reference"
                   else "oracle port of the reference step") + " on this GPU with stock PyTorch + cuDNN, same batch; `value` = the faster mode"
    return out


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--base", type=int, default=32)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--schedule", default="auto", choices=["auto", "graph", "split", "eager"],
                    help="auto: CUDA-graph replay of the whole step when the loss path is capturable, else 'split' (network forward and "
                         "backward + all-reduce + optimizer as two graphs, the host-controlled report losses launch by launch in between)")
    ap.add_argument("--no-side-stream", action="store_true", help="weight gradients on the main stream (default: second stream inside the graph)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true")
    ap.add_argument("--packed-labels", action="store_true",
                    help="e2e leg uploads the masks in the reference's bit-packed on-disk format and unpacks them on the device")
    ap.add_argument("--no-allreduce", action="store_true",
                    help="diagnostic for N > 1: every rank steps on its own (no gradient all-reduce) — separates chip-to-chip / host "
                         "variation from communication cost in the max-over-ranks time; NOT a training configuration")
    ap.add_argument("--trace", default=None, help="write the per-launch CUDA-event timeline of the profiled pass to this file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    classes, shape, kinds = cfg["classes"], tuple(cfg["shape"]), list(cfg["kinds"])
    B = len(kinds)
    config = {"workload": workload_name(args.base, B, shape, args.config), "global_batch": B * world,
              "parallelism": f"dp{world}" if world > 1 else "single",
              "l2": "per-step working set (~3 GB of activations) >> 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb, ms = run_cpu_arm(args, as_reference_impl=True)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if args.warmup < 3:
        args.warmup = 3
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # NCCL collectives are captured into the step's CUDA graph
    from rsuper_b200 import synthetic as synth   # synthetic batches with the reference loader's contract + the loss hyper-parameters
    from rsuper_b200 import losses, ops          # (nothing under oracle/ is imported on this arm)
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.train_step import B200TrainStep
    from rsuper_b200.unet import B200UNet

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    seed_img, seed_lab, seed_torch = rank_seeds(rank)
    torch.manual_seed(seed_torch)
    net = B200UNet(1, args.base, num_classes=len(classes), precision=args.precision).to(dev)
    if world > 1:                       # DDP broadcasts rank 0's parameters at construction (train_ddp.py:661)
        for p in net.parameters():
            dist.broadcast(p.data, src=0)
    params = list(net.parameters())
    ema = [p.detach().clone() for p in params]
    opt = B200AdamW(params, lr=6e-4, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5, max_norm=1.0, ema_params=ema, ema_alpha=0.99,
                    capturable=True)
    report = cfg["report"]
    largs = synth.default_loss_args() if report else synth.default_loss_args(report_volume_loss_basic=0.0)
    largs.nan_check = False                 # a host sync: the loss value is NaN-checked after .item() instead (B200TrainStep.check)
    # synthetic batch (seeded per rank); host copies pinned for the e2e leg
    bt = synth.make_batch(kinds, classes, shape, seed=seed_lab, device="cpu")
    bt["image"] = synth.synthetic_image(B, *shape, seed=seed_img)
    keys = ["image", "label"] + (["unk_channels", "mask", "volumes", "diameters"] if report else [])
    host = {k: bt[k].contiguous().pin_memory() for k in keys}
    devb = {k: host[k].to(dev) for k in keys}

    if report:
        def loss_fn(out, lab, unk, msk, vol, dia):
            return losses.calculate_loss(out, lab, unk, largs, None, msk, vol, dia, classes)["overall"]
    else:
        def loss_fn(out, lab):
            return losses.calculate_loss(out, lab, None, largs, None, None, None, None, classes)["overall"]

    schedule = args.schedule
    if schedule == "auto":
        schedule = "graph" if losses.capturable(largs, report) else "split"
    step = B200TrainStep(net, loss_fn, opt, [devb[k] for k in keys], schedule=schedule, process_group=None if args.no_allreduce else pg,
                         side_stream=(False if args.no_side_stream else None), warmup=args.warmup)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dev_inputs = [devb[k] for k in keys]
    for _ in range(args.warmup):
        step(*dev_inputs)
    barrier()

    # ---- timed region 1: device-resident inputs, CUDA events around the K steps (no per-launch instrumentation) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up before the timed regions start
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(*dev_inputs)
    e1.record()
    barrier()
    launches = step.launches_per_step * args.steps
    ms_dev = e0.elapsed_time(e1) / args.steps

    # ---- timed region 2: end to end through the public API with host buffers ----
    # every step: H2D of that step's inputs (image, masks, report targets) from pinned host memory, the train step, D2H of the
    # loss (.item(), like train_ddp.py:363)
    host_inputs = [host[k] for k in keys]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_inputs)
    packed_keys = []
    if args.packed_labels:
        import numpy as np
        packed_keys = [k for k in ("label", "unk_channels", "mask") if k in host]
        packed = {k: torch.from_numpy(np.stack([synth.pack_masks(host[k][b]) for b in range(B)])).pin_memory() for k in packed_keys}
        h2d_bytes = sum(t.numel() * t.element_size() for k, t in host.items() if k not in packed_keys) + sum(t.numel() for t in packed.values())
    if not packed_keys:
        # untimed: first use of the prefetch path (staging buffers, copy stream, pinned-memory bookkeeping are created here)
        for _ in range(2):
            step.prefetch(*host_inputs)
            step()
            step.loss_async().get()
    barrier()
    t0 = time.perf_counter()
    lv = 0.0
    if packed_keys:
        for _ in range(args.steps):
            ins = [ops.unpack_masks(packed[k].to(dev, non_blocking=True), len(classes)) if k in packed_keys else host[k] for k in keys]
            lv = step(*ins).item()
    else:
        # the loop of a prefetching loader: while step i runs, the copy engine uploads batch i + 1 from pinned host memory
        # (B200TrainStep.prefetch: copy stream + staging buffers); K uploads, K steps and K loss read-backs inside the region
        # ... and every step's loss comes back through a 4-byte pinned copy + event (B200TrainStep.loss_async), looked at after
        # the NEXT step has been enqueued, so the GPU never waits for the host between two replays
        step.prefetch(*host_inputs)
        pending = None
        for i in range(args.steps):
            step()
            h = step.loss_async()
            if i + 1 < args.steps:
                step.prefetch(*host_inputs)
            if pending is not None:
                lv = pending.get()
            pending = h
        lv = pending.get()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop() if rank == 0 else None  # sampled every 50 ms across BOTH timed regions (same work)
    B200TrainStep.check(lv)

    # ---- profiled pass (not part of `value`): the same K steps launch by launch on ONE stream, every launch bracketed by
    # CUDA events, so that a kernel's events measure that kernel alone ----
    barrier()
    prev_side, step.side_stream = step.side_stream, False
    eager_body = step._eager_split if schedule == "split" else step._eager
    eager_body()
    barrier()
    ops.PROFILE = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        eager_body()
    p1.record()
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None
    ms_prof = p0.elapsed_time(p1) / args.steps
    step.side_stream = prev_side

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    per_rank = None
    if dist is not None:
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = [round(g[0].item(), 3) for g in gathered]      # every rank's own device time per step (the value uses the max)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    vox = job_voxels(world, B, shape)
    value = mvox_per_s(vox, ms_dev)
    e2e = mvox_per_s(vox, ms_e2e)

    if rank == 0:
        peaks = load_peaks()
        fam = {}
        for name, flops, a, b, _desc in prof:
            d = fam.setdefault(name, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b); d[1] += flops; d[2] += 1
        if args.trace:
            with open(args.trace, "w") as f:
                agg = {}
                for name, flops, a, b, desc in prof:
                    d = agg.setdefault(desc, [0.0, 0.0, 0])
                    d[0] += a.elapsed_time(b); d[1] += flops; d[2] += 1
                for desc, (ms, fl, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
                    tf = f"{fl / (ms * 1e-3) / 1e12:7.0f} TF/s" if fl > 0 else " " * 12
                    f.write(f"{ms / args.steps:8.3f} ms/step  {cnt / args.steps:5.1f} launches  {ms / cnt * 1e3:8.1f} us  {tf}  {desc}\n")
        kern = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] / args.steps,
                    "tflops": (v[1] / (v[0] * 1e-3) / 1e12) if v[0] > 0 and v[1] > 0 else None} for k, v in fam.items()}
        dom = fam.get("conv3_igemm", [0.0, 0.0, 0])
        achieved = dom[1] / (dom[0] * 1e-3) / 1e12 if dom[0] > 0 else None
        traffic, traffic_src = profiled_traffic()
        roof = {"bound": "tensor", "kernel": "conv3_fprop_kernel (TMA-fed tcgen05 implicit GEMM: the fprop + dgrad launches)",
                "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["tf_sustained"]) if achieved else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "traffic_note": "dram bytes of ONE launch at the dominant shape (32->32 channels, 2 x 128^3: 152.2 GFLOP, 537 MB algorithmic "
                                "+ 268 MB residual); `achieved` averages all launch shapes of the family",
                "algorithmic_flops_per_launch": dom[1] / dom[2] if dom[2] else None,
                "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)",
                "avg_launch_ms": dom[0] / dom[2] if dom[2] else None, "launches_per_step": dom[2] / args.steps,
                "share_of_step": (dom[0] / args.steps) / ms_prof if ms_prof else None,
                "timing": f"separate profiled pass of the same {args.steps} steps inside bench.py: launch by launch on one stream, every launch "
                          f"bracketed by CUDA events ({ms_prof:.2f} ms/step; `value` is the clean pass)"}
        cb = tgb = None
        if world == 1 and not args.no_torch_gpu_baseline:
            del step
            torch.cuda.empty_cache()
            try:
                tgb = run_torch_gpu_baseline(args, dev)
            except Exception as e:  # the baseline must never cost the measurement its JSON line
                tgb = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if not args.no_cpu_baseline and world == 1:
            cb, _ = run_cpu_arm(args, as_reference_impl=False)
        flop_roof_mvox = peaks["tf_sustained"] * 1e12 / (MFLOP_PER_VOXEL_STEP * 1e6) / 1e6 * world
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "bf16 operands / f32 storage", "data": "synthetic",
                "config": config,
                "impl_detail": {"schedule": {"graph": "CUDA graph replay of the whole step", "eager": "eager launches",
                                             "split": "two CUDA graphs (UNet forward | UNet backward + all-reduce + optimizer), "
                                                      "calculate_loss with its host-controlled report losses eager in between"}[schedule],
                                "side_stream": bool(prev_side), "precision": args.precision,
                                "optimizer": "B200AdamW (fused clip+AdamW+EMA kernel)",
                                "gradient_allreduce": (None if world == 1 else "DISABLED (--no-allreduce diagnostic)" if args.no_allreduce else
                                                       "NCCL all-reduce (AVG) of the flat fp32 gradient buffer in four buckets on a communication "
                                                       "stream while backward runs, inside the captured step"),
                                "ms_per_step_per_rank": per_rank,
                                "masks_h2d": "bit-packed (np.packbits) + device unpack" if args.packed_labels else "uint8"},
                "conv3d_flop_roofline_frac": value / flop_roof_mvox,
                "roofline": roof, "kernels": kern, "cpu_baseline": cb, "torch_gpu_baseline": tgb,
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if dist is not None:
        # Teardown with NCCL collectives captured in a live CUDA graph: release the graph first, then leave without
        # ProcessGroupNCCL's destructor (it can wait forever on communicators the captured kernels still reference; the JSON
        # line is already flushed and every rank has passed the barrier below).
        try:
            step.graph = None
        except NameError:
            pass
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
