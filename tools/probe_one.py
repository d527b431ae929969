"""One conv launch for ncu source-level captures: python tools/probe_one.py N D H W Cin Cout [res] [mask]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from rsuper_b200 import ops
N, D, H, W, Cin, Cout = [int(v) for v in sys.argv[1:7]]
res = "res" in sys.argv
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
y = torch.zeros(N, D, H, W, Cout, dtype=torch.bfloat16, device=dev)
r = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16) if res else None
ost = torch.zeros(N, Cout, 2, device=dev)
wp = ops.conv3_pack_weights(w)
for _ in range(3):
    ops.conv3_forward(x, wp, y, res=r, out_stats=ost)
torch.cuda.synchronize()
print("done")
