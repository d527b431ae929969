"""Level-0 satellite kernels in isolation (timed with CUDA events; also the target of a light ncu capture):
python tools/probe_sat.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from rsuper_b200 import ops
dev = "cuda"
N, S, C = 2, 128, 32
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)

def rnd(*shape, dtype=bf):
    return torch.randn(*shape, generator=g).to(dev).to(dtype)

ONLY = sys.argv[1] if len(sys.argv) > 1 else ""


def timed(name, fn, bytes_moved, reps=5):
    if ONLY and ONLY not in name:
        return
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} {ms * 1e3:8.1f} us   {bytes_moved / ms / 1e6:7.0f} GB/s (algorithmic bytes {bytes_moved / 1e6:.0f} MB)")

x32 = rnd(N, S, S, S, C)
y16 = torch.empty(N, S // 2, S // 2, S // 2, C, dtype=bf, device=dev)
st = torch.zeros(N, C, 2, device=dev)
E = x32.numel() * 2
timed("maxpool2_forward 32@128^3", lambda: ops.maxpool2_forward(x32, y16, st), E + E // 8)
dy16 = rnd(N, S // 2, S // 2, S // 2, C)
dsk = rnd(N, S, S, S, C)
dx32 = torch.empty_like(x32)
timed("maxpool2_backward 32@128^3", lambda: ops.maxpool2_backward(x32, dy16, dx32, dskip=dsk), 3 * E + E // 8)
x64 = rnd(N, S // 2, S // 2, S // 2, 64)
cat = torch.empty(N, S, S, S, 96, dtype=bf, device=dev)
stc = torch.zeros(N, 96, 2, device=dev)
up = cat[..., 32:]
E64 = N * S ** 3 * 64 * 2
timed("upsample_forward 64 -> 128^3", lambda: ops.upsample_forward(x64, up, stc[:, 32:]), E64 + E64 // 8)
dcat = rnd(N, S, S, S, 96)
dxu = torch.empty_like(x64)
timed("upsample_backward separable", lambda: ops.upsample_backward(dcat[..., 32:], dxu), E64 + E64 // 8)
timed("upsample_backward 3-D gather", lambda: ops.upsample_backward(dcat[..., 32:], dxu, two_pass=False), E64 + E64 // 8)
img = torch.randn(N, 1, S, S, S, generator=g).to(dev)
w0 = (torch.randn(C, 1, 3, 3, 3, generator=g) / 5).to(dev)
t0 = torch.empty(N, S, S, S, C, dtype=bf, device=dev)
timed("stem_conv_forward", lambda: ops.stem_conv_forward(img, w0, t0, st), E + img.numel() * 4)
dws = torch.empty_like(w0)
timed("stem_conv_wgrad", lambda: ops.stem_conv_wgrad(img, x32, dws), E + img.numel() * 4)
logits = torch.empty(N, 2, S, S, S, device=dev)
hw, hb = torch.randn(2, C, generator=g).to(dev), torch.zeros(2, device=dev)
timed("head_forward", lambda: ops.head_forward(x32, hw, hb, logits), E + logits.numel() * 4)
dwh, dbh = torch.empty_like(hw), torch.empty_like(hb)
timed("head_backward", lambda: ops.head_backward(x32, hw, logits, dx32, dwh, dbh), 2 * E + logits.numel() * 4)
stx = ops.channel_stats(x32)
timed("norm_act 32@128^3", lambda: ops.norm_act(x32, stx), 2 * E)
sums = torch.zeros(N, C, 2, device=dev)
timed("instnorm_backward_apply (+add)", lambda: ops.instnorm_backward_apply(dsk, x32, stx, sums, dx32, add=x32), 4 * E)
timed("channel_stats 32@128^3", lambda: ops.channel_stats(x32), E)
