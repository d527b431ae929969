"""NEXT-ROUND PROBE (written without GPU time left; not part of the product or the tests).

BASELINE.md §3 "also reported": the reference step (oracle port: plain torch ops = cuDNN / ATen kernels) on ONE B200 with
stock PyTorch, as the fair same-box comparison for rsuper_b200's number:

    python tools/probe_torch_gpu_baseline.py            # fp32 (TF32 off, the parity setting), TF32 on, bf16 autocast

Prints Mvoxels/s of the train step (fwd + masked BCE/Dice + bwd + clip + fused AdamW + EMA) at BASELINE.json configs[1]
(base-32 UNet, batch 2 x 128^3, 2 classes).  channels_last_3d is used for the bf16 run (cuDNN's tensor-core layout).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import losses_ref as LR
from oracle import synth
from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward

dev = torch.device("cuda:0")
S, B, BASE = 128, 2, 32
CLASSES = ["organ", "pancreatic_lesion"]
img = synthetic_image(B, S, S, S, seed=1).to(dev)
lab = synth.make_batch(["mask"] * B, CLASSES, (S, S, S), seed=2)["label"].to(dev).long()
largs = LR.default_args(report_volume_loss_basic=0.0)


def run(name, tf32, autocast):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    sd = {k: torch.nn.Parameter(v) for k, v in synthetic_state_dict(BASE, len(CLASSES), device=dev).items()}
    params = list(sd.values())
    ema = [p.detach().clone() for p in params]
    opt = torch.optim.AdamW(params, lr=6e-4, weight_decay=0.05, eps=1e-5, fused=True)
    x = img.contiguous(memory_format=torch.channels_last_3d) if autocast else img

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits = unet_forward(x, sd)
        loss = LR.calculate_loss({"segmentation": logits.float()}, lab, None, largs, None, None, None, None, CLASSES)
        loss["overall"].backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        with torch.no_grad():
            torch._foreach_mul_(ema, 0.99)
            torch._foreach_add_(ema, [p.detach() for p in params], alpha=0.01)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"stock PyTorch {name:24s}: {ms:8.2f} ms/step  {B * S ** 3 / ms / 1e3:8.1f} Mvoxels/s   "
          f"(peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB)")


run("fp32 (TF32 off)", False, False)
run("fp32 storage, TF32 on", True, False)
run("bf16 autocast, NDHWC", True, True)
