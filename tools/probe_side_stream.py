"""NEXT-ROUND PROBE (not part of the product or the tests): why is the second stream box-dependent?

Runs the train step in three schedules and prints GPU time per step, host enqueue time per step and the spread over steps:
    serial      everything on the caller's stream (default)
    side        weight gradients + weight packing on a second stream (RSB_SIDE_STREAM=1 behaviour)
    side-lowpri the same with the main chain on a HIGH-priority stream (so that the dgrad on the critical path wins the SMs
                whenever both persistent kernels are pending)
Observed in round 1: `side` = 21.8-22.3 ms on four boxes, 27-30 ms on three others where `serial` was 22.7 ms.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch

from oracle import losses_ref as LR
from oracle import synth
from oracle.unet_ref import synthetic_image
from rsuper_b200 import losses
from rsuper_b200 import unet as unet_mod
from rsuper_b200.unet import B200UNet

dev = torch.device("cuda:0")
S, B = 128, 2
CLASSES = ["organ", "pancreatic_lesion"]
net = B200UNet(1, 32, num_classes=2).to(dev)
params = list(net.parameters())
opt = torch.optim.AdamW(params, lr=6e-4, weight_decay=0.05, eps=1e-5, fused=True)
img = synthetic_image(B, S, S, S, seed=1).to(dev)
lab = synth.make_batch(["mask"] * B, CLASSES, (S, S, S), seed=2)["label"].to(dev)
largs = LR.default_args(report_volume_loss_basic=0.0)
largs.nan_check = False


def step():
    opt.zero_grad(set_to_none=True)
    loss = losses.calculate_loss(net(img), lab, None, largs, None, None, None, None, CLASSES)
    loss["overall"].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()


def run(name, side, main_stream=None):
    unet_mod.set_side_stream(side)
    ctx = torch.cuda.stream(main_stream) if main_stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        per = []
        cpu = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            step()
            e1.record()
            cpu.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()
            per.append(e0.elapsed_time(e1))
    print(f"{name:12s} GPU ms/step (synchronised steps): min {min(per):.2f} median {sorted(per)[5]:.2f} max {max(per):.2f} | "
          f"host enqueue median {sorted(cpu)[5]:.2f} ms")


run("serial", False)
run("side", True)
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
run("side-lowpri", True, torch.cuda.Stream(device=dev, priority=hi))
print("nproc", os.cpu_count(), "loadavg", open("/proc/loadavg").read().strip())
