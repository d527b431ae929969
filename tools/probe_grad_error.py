"""Per-parameter gradient error of the parity mode (precision fp32) against the fp64 oracle, in network order.
    python tools/probe_grad_error.py [base] [side]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "r-super_b200")):
    sys.path.insert(0, p)
import torch

from oracle import losses_ref as LR
from oracle import synth
from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
from rsuper_b200 import losses
from rsuper_b200.unet import B200UNet

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
base = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
classes = ["organ", "pancreatic_lesion"]


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


sd = synthetic_state_dict(base, 2, device=dev)
x = synthetic_image(2, S, S, S, seed=5, device=dev)
batch = synth.make_batch(["mask", "mask"], classes, (S, S, S), seed=7, device=dev)
args = LR.default_args(report_volume_loss_basic=0.0)
net = B200UNet(1, base, num_classes=2, precision=precision).to(dev)
net.load_state_dict(sd)
out = net(x)
out["segmentation"].retain_grad()
loss = losses.calculate_loss(out, batch["label"], None, args, None, None, None, None, classes)
loss["overall"].backward()


def oracle(dtype):
    sdr = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    lg = unet_forward(x.to(dtype), sdr)
    lg.retain_grad()
    l = LR.calculate_loss({"segmentation": lg}, batch["label"].long(), None, args, None, None, None, None, classes)
    l["overall"].backward()
    return l["overall"].item(), {k: v.grad for k, v in sdr.items()}, lg.grad


l64, g64, dl64 = oracle(torch.float64)
l32, g32, dl32 = oracle(torch.float32)
print(f"base {base} {S}^3 {precision}: loss ours {loss['overall'].item():.7f} fp32 oracle {l32:.7f} fp64 {l64:.7f}")
for k, p in net.named_parameters():
    print(f"  {k:34s} ours {rel(p.grad, g64[k]):.3e}   fp32 oracle {rel(g32[k], g64[k]):.3e}   |g|max {g64[k].abs().max().item():.3e}")
