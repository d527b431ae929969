"""Stage-by-stage comparison of the B200UNet engine against the (emulating) oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
import torch.nn.functional as F

from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
from rsuper_b200 import ops
from rsuper_b200.unet import _Engine

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def nc(t):
    return t.float().permute(0, 4, 1, 2, 3)


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def run(precision, base=8, C=2, S=32, N=2):
    sd = synthetic_state_dict(base, C, device=dev)
    x = synthetic_image(N, S, S, S, seed=4, device=dev)
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    eng = _Engine(base, 0.0, dt)
    with torch.no_grad():
        logits, Sv = eng.forward(x, sd, C, save=True)
        tr = {}
        ref = unet_forward(x, sd, emulate=True, storage=precision, trace=tr)
    saved, enc = Sv["saved"], Sv["enc_out"]
    print(f"=== precision {precision}")
    print("t0      ", rel(nc(saved[0][0].t), tr["t0"]))
    print("inc     ", rel(nc(enc[0].t), tr["inc"]))
    for l in range(1, 5):
        print(f"pool{l}   ", rel(nc(saved[2 * l - 1][0].t), tr[f"pool{l}"]))
        print(f"down{l}.1 ", rel(nc(saved[2 * l][0].t), tr[f"down{l}.1"]))
        print(f"down{l}.2 ", rel(nc(enc[l].t), tr[f"down{l}.2"]))
    for j in range(1, 5):
        print(f"up{j}.cat  ", rel(nc(saved[7 + 2 * j][0].t), tr[f"up{j}.cat"]))
        print(f"up{j}.0    ", rel(nc(saved[8 + 2 * j][0].t), tr[f"up{j}.0"]))
    print("final   ", rel(nc(Sv["final"].t), tr["up4.1"]))
    print("logits  ", rel(logits, ref))
    # statistics sanity: stored stats vs recomputed from the stored tensors
    for name, act in (("t0", saved[0][0]), ("inc", enc[0]), ("down1.1", saved[2][0]), ("up4.cat", saved[15][0])):
        t = act.t.float()
        n, c = t.shape[0], t.shape[4]
        v = t.reshape(n, -1, c).double()
        st = torch.stack([v.sum(1), (v * v).sum(1)], -1).float()
        print(f"stats[{name}] rel err vs recomputed-from-stored: {rel(act.st, st):.3e}")


def wgrad_case(N, D, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(dev)
    dy = torch.randn(N, D, H, W, Cout, generator=g).to(dev)
    xn = x.permute(0, 4, 1, 2, 3)
    a = F.relu(F.instance_norm(xn, eps=1e-4)).to(torch.bfloat16).float()
    wz = torch.zeros(Cout, Cin, 3, 3, 3, device=dev, requires_grad=True)
    (F.conv3d(a, wz, padding=1) * dy.permute(0, 4, 1, 2, 3).to(torch.bfloat16).float()).sum().backward()
    v = x.reshape(N, -1, Cin).double()
    st = torch.stack([v.sum(1), (v * v).sum(1)], -1).float()
    dw = torch.zeros(Cout, Cin, 3, 3, 3, device=dev)
    ops.conv3_wgrad(x, dy, dw, in_stats=st)
    err = (dw - wz.grad).abs()
    m = wz.grad.abs().max().item()
    print(f"wgrad N{N} D{D} H{H} W{W} Cin{Cin} Cout{Cout}: rel {err.max().item() / m:.3e}")
    e_co = err.amax(dim=(1, 2, 3, 4)) / m
    e_ci = err.amax(dim=(0, 2, 3, 4)) / m
    e_tap = err.reshape(Cout, Cin, 27).amax(dim=(0, 1)) / m
    print("  by co (blocks of 32):", [f"{e_co[i:i + 32].max().item():.1e}" for i in range(0, Cout, 32)])
    print("  by ci (blocks of 32):", [f"{e_ci[i:i + 32].max().item():.1e}" for i in range(0, Cin, 32)])
    print("  by tap:", [f"{v:.1e}" for v in e_tap.tolist()])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "fp32"):
        run("fp32")
    if what in ("all", "bf16"):
        run("bf16")
    if what in ("all", "wgrad"):
        wgrad_case(1, 2, 8, 8, 256, 320)
        wgrad_case(1, 2, 8, 8, 256, 128)
        wgrad_case(1, 2, 8, 8, 128, 320)
        wgrad_case(1, 4, 16, 16, 256, 320)
