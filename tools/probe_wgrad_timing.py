import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import ctypes as C
import torch
from rsuper_b200 import ops
from rsuper_b200._lib import lib
dev = "cuda"

def run(N, D, H, W, Cin, Cout, max_ctas=0):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    dy = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(torch.bfloat16)
    dw = torch.zeros(Cout, Cin, 3, 3, 3, device=dev)
    buf = torch.zeros(1024 * 16, dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.conv3_wgrad(x, dy, dw, max_ctas=max_ctas)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        ops.conv3_wgrad(x, dy, dw, max_ctas=max_ctas)
    f1.record()
    torch.cuda.synchronize()
    ms_plain = f0.elapsed_time(f1) / 5
    lib().rsb_debug_set_wgrad_timing_buffer(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv3_wgrad(x, dy, dw, max_ctas=max_ctas); e1.record()
    torch.cuda.synchronize()
    lib().rsb_debug_set_wgrad_timing_buffer(None)
    ms = e0.elapsed_time(e1)
    b = buf.view(1024, 16).double()
    b = b[b[:, 0] > 0]
    m = b.mean(0)
    fl = 2.0 * 27 * Cin * Cout * N * D * H * W
    print(f"[max_ctas {max_ctas}] wgrad N{N} {D}x{H}x{W} {Cin}->{Cout}: {ms_plain:.3f} ms {fl / ms_plain / 1e9:.0f} TF/s (instrumented {ms:.3f} ms) | CTAs {b.shape[0]} steps/CTA {m[2]:.0f} cyc/step {m[0] / m[2]:.0f} | "
          f"MMA warp waits: a-plane {100 * m[1] / m[0]:.0f}% dy {100 * m[3] / m[0]:.0f}% | SM clock during loop {m[0] / m[4] * 1e3:.0f} MHz")

if __name__ == "__main__":
    print("RSB_WGRAD_DEBUG_MODE =", os.environ.get("RSB_WGRAD_DEBUG_MODE"))
    run(2, 128, 128, 128, 32, 32)
    run(2, 128, 128, 128, 96, 64)
    run(2, 64, 64, 64, 64, 64)
    run(2, 64, 64, 64, 32, 128)
    for mc in (16, 37, 74):
        run(2, 128, 128, 128, 32, 32, mc)
    if os.environ.get("RSB_WGRAD_DEBUG_MODE"):
        sys.exit(0)
    run(2, 64, 64, 64, 192, 128)
    run(2, 32, 32, 32, 128, 128)
    run(2, 32, 32, 32, 384, 256)
    run(2, 16, 16, 16, 256, 256)
    run(2, 16, 16, 16, 576, 512)
    run(2, 8, 8, 8, 320, 320)
