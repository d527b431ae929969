"""Per-role pipeline wait breakdown of the conv3 kernel (uses rsb_debug_set_timing_buffer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import ctypes as C
import torch
from rsuper_b200 import ops
from rsuper_b200._lib import lib
dev = "cuda"

def stats_of(x):
    n, c = x.shape[0], x.shape[4]
    v = x.reshape(n, -1, c).double()
    return torch.stack([v.sum(1), (v * v).sum(1)], -1).float()

def run(N, D, H, W, Cin, Cout, res=True, pz=0):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
    y = torch.zeros(N, D, H, W, Cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16) if res else None
    ost = torch.zeros(N, Cout, 2, device=dev)
    wp = ops.conv3_pack_weights(w)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.conv3_forward(x, wp, y, res=r, out_stats=ost, planes_per_item=pz)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        ops.conv3_forward(x, wp, y, res=r, out_stats=ost, planes_per_item=pz)
    f1.record()
    torch.cuda.synchronize()
    ms_plain = f0.elapsed_time(f1) / 5
    lib().rsb_debug_set_timing_buffer(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3_forward(x, wp, y, res=r, out_stats=ost, planes_per_item=pz)
    e1.record()
    torch.cuda.synchronize()
    lib().rsb_debug_set_timing_buffer(None)
    ms = e0.elapsed_time(e1)
    b = buf.view(148, 16).double()
    m = b.mean(0)
    fl = 2.0 * 27 * Cin * Cout * N * D * H * W
    print(f"N{N} {D}x{H}x{W} {Cin}->{Cout} pz={pz}: {ms_plain:.3f} ms {fl / ms_plain / 1e9:.0f} TF/s (instrumented {ms:.3f}) | items/CTA {m[4]:.1f} | "
          f"MMA thread total {m[0]:.0f} cyc: wait a_full {100 * m[1] / m[0]:.0f}% b_full {100 * m[2] / m[0]:.0f}% acc_empty {100 * m[3] / m[0]:.0f}% "
          f"issue/other {100 * (m[0] - m[1] - m[2] - m[3]) / m[0]:.0f}% | "
          f"epilogue wait acc_full {100 * m[8] / m[7]:.0f}% | cyc/item {m[0] / m[4]:.0f}")

if __name__ == "__main__":
    run(2, 128, 128, 128, 32, 32)
    run(2, 128, 128, 128, 32, 32, pz=2)
    run(2, 128, 128, 128, 32, 32, res=False)
    run(2, 128, 128, 128, 96, 64, res=False)
    run(2, 128, 128, 128, 64, 32)
    run(2, 64, 64, 64, 64, 64)
    run(2, 64, 64, 64, 192, 128, res=False)
    run(2, 64, 64, 64, 128, 192, res=False)
    run(2, 32, 32, 32, 128, 128)
    run(2, 16, 16, 16, 256, 256)
    run(2, 16, 16, 16, 576, 512, res=False)
    run(2, 8, 8, 8, 320, 320)
