"""Error structure of ONE split-precision conv (parity mode) against fp64, as a function of K = 27 * Cin:
is the floor the tensor pipe's accumulation (grows ~K, signed like the result = truncation) or the operand split?
    python tools/probe_accum_error.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "r-super_b200")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F

from rsuper_b200 import ops

torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)


def nc(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


def cl(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


for Cin, Cout in [(32, 32), (96, 32), (320, 64), (576, 64)]:
    for relu in (True, False):
        x = torch.randn(1, 8, 16, 16, Cin, generator=g).to(dev)
        if relu:
            x = x.relu()            # post-ReLU operands: all products' sign = sign(w): sums do not centre at 0
        w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
        ref64 = cl(F.conv3d(nc(x).double(), w.double(), padding=1))
        ref32 = cl(F.conv3d(nc(x), w, padding=1))
        res = {}
        for tag, split in (("3 products", True), ("6 products", 6)):
            y = torch.zeros(1, 8, 16, 16, Cout, device=dev)
            pieces = ops.norm_act(x, None, split=True if split is True else 3)
            kw = dict(a_lo=pieces[1])
            if split == 6:
                kw["a_lo2"] = pieces[2]
            ops.conv3_forward(pieces[0], ops.conv3_pack_weights(w, split=split), y, **kw)
            res[tag] = y.double()
        scale = ref64.abs().max()
        rms = ref64.pow(2).mean().sqrt()
        line = f"Cin {Cin:3d} Cout {Cout:2d} relu={int(relu)}  K={27 * Cin:5d}:"
        for tag, y in list(res.items()) + [("cuDNN fp32", ref32.double())]:
            e = y - ref64
            toward0 = (e * ref64.sign()).mean() / rms          # < 0: magnitudes shrink (truncation toward zero)
            line += f"  [{tag}] max/max {e.abs().max() / scale:.2e} rms/rms {e.pow(2).mean().sqrt() / rms:.2e} bias*sign {toward0:+.2e}"
        print(line)
