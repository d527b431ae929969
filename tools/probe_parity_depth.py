"""Where does the logits error of the parity mode come from, and how does it grow with the patch size?

For S in sizes: ours (precision fp32 = split 3 x bf16 products) and the oracle in fp32 (cuDNN, TF32 off) are both compared with
the oracle evaluated in fp64 on the same GPU, stage by stage (rel-to-max of the stage output), and every InstanceNorm
statistic our epilogues accumulated (sum, sumsq in fp32) is compared with the fp64 statistics of the very tensor it describes.
    python tools/probe_parity_depth.py [sizes...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "r-super_b200")):
    sys.path.insert(0, p)
import torch

from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
from rsuper_b200.unet import B200UNet, _Engine

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [32, 64, 128]
precision = "bf16" if "bf16" in sys.argv else "fp32"
base = 32


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def ncdhw(t):
    return t.permute(0, 4, 1, 2, 3)


for S in sizes:
    sd = synthetic_state_dict(base, 2, device=dev)
    x = synthetic_image(1, S, S, S, seed=6, device=dev)
    tr32, tr64 = {}, {}
    with torch.no_grad():
        want32 = unet_forward(x, sd, trace=tr32)
        want64 = unet_forward(x.double(), {k: v.double() for k, v in sd.items()}, trace=tr64)
        if precision == "bf16":
            tre = {}
            wante = unet_forward(x, sd, emulate=True, storage="bf16", trace=tre)
    net = B200UNet(1, base, num_classes=2, precision=precision).to(dev)
    net.load_state_dict(sd)
    eng = _Engine(base, 0.0, torch.float32 if precision == "fp32" else torch.bfloat16)
    P = {k: v.detach() for k, v in net.named_parameters()}
    with torch.no_grad():
        logits, Sv = eng.forward(x, P, 2, save=True)
    torch.cuda.synchronize()
    ours = {"inc": Sv["enc_out"][0]}
    for l in range(1, 5):
        ours[f"down{l}.2"] = Sv["enc_out"][l]
    for j, l in enumerate((3, 2, 1, 0), start=1):
        ours[f"up{j}.cat"] = Sv["cat"][l]
        ours[f"up{j}.0"] = Sv["saved"][8 + 2 * j][0]          # input of block up{j}.1 = output of block up{j}.0
    ours["up4.1"] = Sv["final"]
    print(f"==== S={S} precision={precision}: logits ours-vs-fp64 {rel(logits, want64):.3e}  oracle32-vs-fp64 {rel(want32, want64):.3e}  "
          f"ours-vs-oracle32 {rel(logits, want32):.3e}" + (f"  emul-oracle-vs-fp64 {rel(wante, want64):.3e}" if precision == "bf16" else ""))
    for k, a in ours.items():
        t = ncdhw(a.t).float()
        line = f"  {k:9s} ours-vs-fp64 {rel(t, tr64[k]):.3e}   oracle32-vs-fp64 {rel(tr32[k], tr64[k]):.3e}"
        if k.endswith(".cat"):
            nsk = {"up1.cat": 8, "up2.cat": 4, "up3.cat": 2, "up4.cat": 1}[k] * base
            line += f" (skip half {rel(t[:, :nsk], tr64[k][:, :nsk]):.2e}, upsampled half {rel(t[:, nsk:], tr64[k][:, nsk:]):.2e})"
        if a.st is not None:
            n = t[0, 0].numel()
            t64 = t.double()
            m64 = t64.mean(dim=(2, 3, 4))
            v64 = t64.var(dim=(2, 3, 4), unbiased=False)
            st = a.st.double()
            m = st[..., 0] / n
            v = st[..., 1] / n - m * m
            r64, r = torch.rsqrt(v64 + 1e-4), torch.rsqrt(v.clamp_min(0) + 1e-4)
            line += (f"   stats: mean abs err/std {((m - m64).abs() * r64).max().item():.2e}  rstd rel err {((r - r64).abs() / r64).max().item():.2e}"
                     f"  max mean^2/var {(m64 * m64 / v64).max().item():.1f}")
        print(line)
