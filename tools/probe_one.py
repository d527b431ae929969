"""One conv configuration, timed (and suitable for ncu source-level captures):
python tools/probe_one.py N D H W Cin Cout [res] [mask] [nostats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from rsuper_b200 import ops
N, D, H, W, Cin, Cout = [int(v) for v in sys.argv[1:7]]
res, mask, nostats = "res" in sys.argv, "mask" in sys.argv, "nostats" in sys.argv
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
y = torch.zeros(N, D, H, W, Cout, dtype=torch.bfloat16, device=dev)
r = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16) if (res or mask) else None
ost = None if nostats else torch.zeros(N, Cout, 2, device=dev)
wp = ops.conv3_pack_weights(w)
kw = dict(res=r, out_stats=ost)
if mask:
    mst = ops.channel_stats(r)
    kw = dict(mask_x=r, mask_stats=mst, bwd_sums=torch.zeros(N, Cout, 2, device=dev))
for _ in range(3):
    ops.conv3_forward(x, wp, y, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv3_forward(x, wp, y, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{' '.join(sys.argv[1:])}: {ms:.3f} ms {2.0 * 27 * Cin * Cout * N * D * H * W / ms / 1e9:.0f} TF/s")
