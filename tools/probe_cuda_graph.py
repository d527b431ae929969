"""NEXT-ROUND PROBE (written without GPU time left; not part of the product or the tests).

Captures one whole train step (forward, loss WITHOUT the host-side NaN check, backward, clip, fused AdamW with
capturable=True, EMA) into a CUDA graph and compares replay time against eager execution:

    python tools/probe_cuda_graph.py [--side]     # --side: weight gradients on the second stream inside the capture

Why: the step issues ~300 launches and the host needs ~16 ms to enqueue them against ~22 ms of GPU time; on hosts with
slow / shared cores the eager step becomes enqueue-bound.  Everything the kernels need is capturable:
launches go to torch's current stream, tensor maps are encoded on the host at capture time (pointers are stable inside
the graph's private pool), no kernel allocates or synchronises.  Known obstacles to check when this runs:
  * losses.calculate_loss: the NaN check and any `.item()` are host syncs -> args.nan_check = False here;
  * PackPlan holds parameter pointers (stable) and its images are allocated before capture (warm-up does that);
  * the engine's zero pool / activations are allocated during capture -> they live in the graph pool (fine);
  * torch.nn.utils.clip_grad_norm_ with foreach=True is capturable when error_if_nonfinite=False.
"""
import argparse
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

ap = argparse.ArgumentParser()
ap.add_argument("--side", action="store_true")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda:0")
S, B = 128, 2
CLASSES = ["organ", "pancreatic_lesion"]
unet_mod.set_side_stream(args.side)
net = B200UNet(1, 32, num_classes=2).to(dev)
params = list(net.parameters())
opt = torch.optim.AdamW(params, lr=6e-4, weight_decay=0.05, eps=1e-5, fused=True, capturable=True)
ema = [p.detach().clone() for p in params]
img = synthetic_image(B, S, S, S, seed=1).to(dev)
lab = synth.make_batch(["mask"] * B, CLASSES, (S, S, S), seed=2)["label"].to(dev)
largs = LR.default_args(report_volume_loss_basic=0.0)
largs.nan_check = False
static_loss = torch.zeros((), device=dev)


def step():
    opt.zero_grad(set_to_none=False)   # grads must keep their addresses across replays
    out = net(img)
    loss = losses.calculate_loss(out, lab, None, largs, None, None, None, None, CLASSES)
    loss["overall"].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
    opt.step()
    with torch.no_grad():
        torch._foreach_mul_(ema, 0.99)
        torch._foreach_add_(ema, [p.detach() for p in params], alpha=0.01)
        static_loss.copy_(loss["overall"].detach())


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t_cpu = (time.perf_counter() - t0) * 1e3 / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu


# warm-up on a side stream (torch.cuda.graphs recipe), then eager timing
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
ms, cpu = timed(step, args.steps)
print(f"eager   : {ms:.2f} ms/step on the GPU, host enqueue {cpu:.2f} ms/step, loss {static_loss.item():.5f}")

g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
ms, cpu = timed(g.replay, args.steps)
print(f"replay  : {ms:.2f} ms/step on the GPU, host enqueue {cpu:.3f} ms/step, loss {static_loss.item():.5f}")
