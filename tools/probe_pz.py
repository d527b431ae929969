"""fprop timing vs planes-per-item for the deep layers (B-traffic bound?): python tools/probe_pz.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from rsuper_b200 import ops
dev = "cuda"

def run(N, D, H, W, Cin, Cout, pz, mask=False):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
    y = torch.zeros(N, D, H, W, Cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16)
    ost = torch.zeros(N, Cout, 2, device=dev)
    wp = ops.conv3_pack_weights(w)
    kw = dict(res=r, out_stats=ost)
    if mask:
        kw = dict(mask_x=r, mask_stats=ops.channel_stats(r), bwd_sums=torch.zeros(N, Cout, 2, device=dev))
    try:
        for _ in range(3):
            ops.conv3_forward(x, wp, y, planes_per_item=pz, **kw)
    except RuntimeError as e:
        print(f"{D}^3 {Cin}->{Cout} pz={pz}: {str(e)[-80:]}")
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv3_forward(x, wp, y, planes_per_item=pz, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{D}^3 {Cin}->{Cout} pz={pz}{' mask' if mask else ''}: {ms * 1e3:.1f} us {2.0 * 27 * Cin * Cout * N * D * H * W / ms / 1e9:.0f} TF/s")

for (d, ci, co) in ((32, 128, 128), (32, 384, 256), (32, 256, 384), (16, 256, 256), (16, 576, 512), (16, 512, 576), (8, 320, 320), (8, 640, 256), (64, 64, 64), (64, 128, 192), (64, 192, 128)):
    for pz in (0, 1, 2, 4):
        run(2, d, d, d, ci, co, pz)
