"""Who waits on whom in the 32-channel conv (fprop / residual / dgrad-mask epilogues): per-role pipeline waits of the generic
kernel (rsb_debug_set_timing_buffer) next to the production (compile-time N tile) timings.
    python tools/probe_epilogue.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import ctypes as C
import torch
from rsuper_b200 import ops
from rsuper_b200._lib import lib
dev = "cuda"


def run(N, D, H, W, Cin, Cout, mode, pz=0):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(dev)
    y = torch.zeros(N, D, H, W, Cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16)
    ost = torch.zeros(N, Cout, 2, device=dev)
    wp = ops.conv3_pack_weights(w)
    if mode == "plain":
        kw = dict()
    elif mode == "stats":
        kw = dict(out_stats=ost)
    elif mode == "res":
        kw = dict(res=r, out_stats=ost)
    else:
        kw = dict(mask_x=r, mask_stats=ops.channel_stats(r), bwd_sums=torch.zeros(N, Cout, 2, device=dev))
    kw["planes_per_item"] = pz
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.conv3_forward(x, wp, y, **kw)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        ops.conv3_forward(x, wp, y, **kw)
    f1.record()
    torch.cuda.synchronize()
    ms_plain = f0.elapsed_time(f1) / 5
    lib().rsb_debug_set_timing_buffer(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3_forward(x, wp, y, **kw)
    e1.record()
    torch.cuda.synchronize()
    lib().rsb_debug_set_timing_buffer(None)
    ms = e0.elapsed_time(e1)
    m = buf.view(148, 16).double().mean(0)
    fl = 2.0 * 27 * Cin * Cout * N * D * H * W
    print(f"{Cin}->{Cout} {N}x{D}x{H}x{W} {mode:5s} pz={pz}: {ms_plain * 1e3:7.1f} us {fl / ms_plain / 1e9:5.0f} TF/s (generic+hooks {ms * 1e3:7.1f} us) | items/CTA {m[4]:.1f} | "
          f"MMA warp {m[0] / m[4]:.0f} cyc/item: wait a_full {100 * m[1] / m[0]:.0f}% b_full {100 * m[2] / m[0]:.0f}% acc_empty {100 * m[3] / m[0]:.0f}% "
          f"issue {100 * (m[0] - m[1] - m[2] - m[3]) / m[0]:.0f}% | epilogue {m[7] / m[4]:.0f} cyc/item, waits acc_full {100 * m[8] / m[7]:.0f}%")


if __name__ == "__main__":
    for mode in ("plain", "stats", "res", "mask"):
        run(2, 128, 128, 128, 32, 32, mode)
    for mode in ("plain", "mask"):
        run(2, 128, 128, 128, 32, 32, mode, pz=2)
    for mode in ("stats", "res", "mask"):
        run(2, 64, 64, 64, 64, 64, mode)
    run(2, 128, 128, 128, 64, 96, "mask")
    run(2, 128, 128, 128, 96, 64, "stats")
