"""One wgrad configuration, timed (also the target of ncu source-level captures): python tools/probe_wgrad_one.py N D H W Cin Cout"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from rsuper_b200 import ops
N, D, H, W, Cin, Cout = [int(v) for v in sys.argv[1:7]]
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
dy = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16)
dw = torch.zeros(Cout, Cin, 3, 3, 3, device=dev)
for _ in range(3):
    ops.conv3_wgrad(x, dy, dw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv3_wgrad(x, dy, dw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"wgrad {' '.join(sys.argv[1:])}: {ms:.3f} ms {2.0 * 27 * Cin * Cout * N * D * H * W / ms / 1e9:.0f} TF/s")
