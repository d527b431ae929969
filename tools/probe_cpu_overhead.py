"""CPU enqueue cost of one train step (the GPU must not starve on launches): python tools/probe_cpu_overhead.py"""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from oracle import losses_ref as LR
from oracle import synth
from oracle.unet_ref import synthetic_image
from rsuper_b200 import losses
from rsuper_b200.unet import B200UNet
dev = torch.device("cuda:0")
S, B = 128, 2
CLASSES = ["organ", "pancreatic_lesion"]
net = B200UNet(1, 32, num_classes=2).to(dev)
params = list(net.parameters())
opt = torch.optim.AdamW(params, lr=6e-4, weight_decay=0.05, eps=1e-5, fused=True)
ema = [p.detach().clone() for p in params]
img = synthetic_image(B, S, S, S, seed=1).to(dev)
lab = synth.make_batch(["mask"] * B, CLASSES, (S, S, S), seed=2)["label"].to(dev)
largs = LR.default_args(report_volume_loss_basic=0.0)

def step():
    opt.zero_grad(set_to_none=True)
    out = net(img)
    loss = losses.calculate_loss(out, lab, None, largs, None, None, None, None, CLASSES)
    loss["overall"].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()
    with torch.no_grad():
        torch._foreach_mul_(ema, 0.99)
        torch._foreach_add_(ema, [p.detach() for p in params], alpha=0.01)
    return loss["overall"]

for _ in range(3):
    step()
torch.cuda.synchronize()
# CPU-only cost: one step issued into an idle queue (GPU far behind), no sync inside
t0 = time.perf_counter(); step(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"CPU enqueue time of one step: {(t1 - t0) * 1e3:.2f} ms; until GPU done: {(t2 - t0) * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable(); step(); pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
