"""Bring-up probe for the tcgen05 conv3 kernel: run one named case, print diagnostics.

Usage: python tools/probe_conv3.py <case> ; cases are independent processes so that a trap in one
does not poison the CUDA context of the next.
"""
import os
import sys
import time

os.environ.setdefault("RSB_LOADER_LAX", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))

import torch
import torch.nn.functional as F

from rsuper_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def ref_conv(x_cl, w, in_norm, slope, res_cl):
    """x_cl [N,D,H,W,C] fp32 (already storage-rounded); returns conv output NDHWC fp32 and the activated input."""
    x = x_cl.permute(0, 4, 1, 2, 3).contiguous()
    if in_norm:
        mean = x.mean(dim=(2, 3, 4), keepdim=True)
        var = x.var(dim=(2, 3, 4), keepdim=True, unbiased=False)
        a = (x - mean) * torch.rsqrt(var + 1e-4)
        a = torch.where(a > 0, a, a * slope)
    else:
        a = x
    a = bf16r(a)
    y = F.conv3d(a, bf16r(w), padding=1)
    y = y.permute(0, 2, 3, 4, 1).contiguous()
    if res_cl is not None:
        y = y + res_cl
    return y


def stats_of(x_cl):
    n, c = x_cl.shape[0], x_cl.shape[4]
    v = x_cl.reshape(n, -1, c).double()
    return torch.stack([v.sum(1), (v * v).sum(1)], dim=-1).float()


def run_case(name, N, D, H, W, Cin, Cout, *, dtype=torch.bfloat16, in_norm=False, res=False, stats=False,
             weights="random", pz=0, nt=0, x_extra=0, y_extra=0, slope=0.0, iters=0):
    g = torch.Generator(device="cpu").manual_seed(1234)
    xfull = torch.randn(N, D, H, W, Cin + x_extra, generator=g).to(dev)
    xfull = xfull.to(dtype)
    x = xfull[..., :Cin] if x_extra else xfull
    if weights == "center":
        w = torch.zeros(Cout, Cin, 3, 3, 3)
        for i in range(min(Cout, Cin)):
            w[i, i, 1, 1, 1] = 1.0
    elif weights.startswith("tap"):
        kd, kh, kw = [int(ch) for ch in weights[3:]]
        w = torch.zeros(Cout, Cin, 3, 3, 3)
        for i in range(min(Cout, Cin)):
            w[i, i, kd, kh, kw] = 1.0
    else:
        w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (1.0 / (27 * Cin) ** 0.5)
    w = w.to(dev)
    yfull = torch.zeros(N, D, H, W, Cout + y_extra, dtype=dtype, device=dev)
    y = yfull[..., :Cout] if y_extra else yfull
    res_t = None
    if res:
        res_t = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(dtype)
    in_stats = stats_of(x.float()) if in_norm else None
    out_stats = torch.zeros(N, Cout, 2, device=dev) if stats else None
    wp = ops.conv3_pack_weights(w.contiguous())
    torch.cuda.synchronize()
    t0 = time.time()
    ops.conv3_forward(x, wp, y, in_stats=in_stats, slope=slope, res=res_t, out_stats=out_stats,
                      planes_per_item=pz, n_tile=nt)
    torch.cuda.synchronize()
    t1 = time.time()
    yr = ref_conv(x.float(), w, in_norm, slope, res_t.float() if res else None)
    got = y.float()
    err = (got - yr).abs()
    denom = yr.abs().max().item() + 1e-12
    print(f"[{name}] N{N} D{D} H{H} W{W} Cin{Cin} Cout{Cout} {dtype} pz={pz} nt={nt} "
          f"max_abs_err={err.max().item():.4e} rel_to_max={err.max().item() / denom:.4e} "
          f"mean_abs_err={err.mean().item():.4e} ref_max={denom:.3f} first_call_ms={(t1 - t0) * 1e3:.1f}")
    if err.max().item() / denom > 2e-2:
        bad = (err > 1e-2 * denom).nonzero()
        print(f"   mismatches: {bad.shape[0]} of {err.numel()}; first few idx (n,z,y,x,c):")
        for b in bad[:12].tolist():
            print("    ", b, "got", got[tuple(b)].item(), "ref", yr[tuple(b)].item())
        # per-plane / per-channel error pattern helps diagnose descriptor problems
        print("   err by z:", [f"{v:.2e}" for v in err.amax(dim=(0, 2, 3, 4)).tolist()][:16])
        print("   err by y:", [f"{v:.2e}" for v in err.amax(dim=(0, 1, 3, 4)).tolist()][:20])
        print("   err by x:", [f"{v:.2e}" for v in err.amax(dim=(0, 1, 2, 4)).tolist()][:20])
        print("   err by c:", [f"{v:.2e}" for v in err.amax(dim=(0, 1, 2, 3)).tolist()][:32])
    if y_extra:
        assert yfull[..., Cout:].abs().max().item() == 0, "kernel wrote outside its channel slice"
    if stats:
        sr = stats_of(yr)
        serr = (out_stats - sr).abs().max().item() / (sr.abs().max().item() + 1e-9)
        print(f"   out_stats rel err {serr:.3e}")
    if iters:
        for _ in range(3):
            ops.conv3_forward(x, wp, y, in_stats=in_stats, slope=slope, res=res_t, out_stats=out_stats,
                              planes_per_item=pz, n_tile=nt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv3_forward(x, wp, y, in_stats=in_stats, slope=slope, res=res_t, out_stats=out_stats,
                              planes_per_item=pz, n_tile=nt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * 27 * Cin * Cout * N * D * H * W
        print(f"   time {ms:.3f} ms  -> {flops / ms / 1e9:.1f} TFLOP/s")


def run_dgrad_case(name, N, D, H, W, Cin, Cout, dtype=torch.bfloat16, slope=0.0):
    """dgrad + mask epilogue vs autograd of conv(act(instnorm(x)))."""
    g = torch.Generator(device="cpu").manual_seed(99)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(dev).to(dtype)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (1.0 / (27 * Cin) ** 0.5)).to(dev)
    dy = torch.randn(N, D, H, W, Cout, generator=g).to(dev).to(dtype)
    # reference: da = conv_transpose(dy, w); g = da * act'(xhat); S1 = sum g; S2 = sum g*xhat
    xf = x.float().permute(0, 4, 1, 2, 3)
    mean = xf.mean(dim=(2, 3, 4), keepdim=True)
    var = xf.var(dim=(2, 3, 4), keepdim=True, unbiased=False)
    xhat = (xf - mean) * torch.rsqrt(var + 1e-4)
    da = F.conv_transpose3d(dy.float().permute(0, 4, 1, 2, 3), bf16r(w), padding=1)
    gref = torch.where(xhat > 0, da, da * slope)
    s1 = gref.sum(dim=(2, 3, 4))
    s2 = (gref * xhat).sum(dim=(2, 3, 4))
    gref = gref.permute(0, 2, 3, 4, 1).contiguous()
    wp = ops.conv3_pack_weights(w.contiguous(), transpose_flip=True)
    gout = torch.zeros(N, D, H, W, Cin, dtype=dtype, device=dev)
    sums = torch.zeros(N, Cin, 2, device=dev)
    ops.conv3_forward(dy, wp, gout, mask_x=x, mask_stats=stats_of(x.float()), bwd_sums=sums, slope=slope)
    torch.cuda.synchronize()
    err = (gout.float() - gref).abs().max().item() / (gref.abs().max().item() + 1e-9)
    e1 = (sums[..., 0] - s1).abs().max().item() / (s1.abs().max().item() + 1e-9)
    e2 = (sums[..., 1] - s2).abs().max().item() / (s2.abs().max().item() + 1e-9)
    print(f"[{name}] dgrad rel_err={err:.3e} S1 rel={e1:.3e} S2 rel={e2:.3e}")


CASES = {
    "center16": lambda: run_case("center16", 1, 4, 16, 8, 16, 16, weights="center", pz=1),
    "tap000": lambda: run_case("tap000", 1, 4, 16, 8, 16, 16, weights="tap000", pz=1),
    "tap212": lambda: run_case("tap212", 1, 4, 16, 8, 16, 16, weights="tap212", pz=1),
    "rand16": lambda: run_case("rand16", 1, 4, 16, 8, 16, 16, pz=1),
    "rand32": lambda: run_case("rand32", 2, 8, 32, 16, 32, 32, pz=2),
    "rand32_pz4": lambda: run_case("rand32_pz4", 2, 8, 32, 16, 32, 32, pz=4),
    "norm_res_stats": lambda: run_case("norm_res_stats", 2, 8, 32, 16, 32, 32, in_norm=True, res=True, stats=True),
    "ragged": lambda: run_case("ragged", 1, 5, 20, 12, 24, 40, in_norm=True, stats=True),
    "c96_64": lambda: run_case("c96_64", 1, 8, 32, 32, 96, 64, in_norm=True, stats=True, x_extra=0, y_extra=8),
    "c64_128": lambda: run_case("c64_128", 1, 8, 16, 16, 64, 128, in_norm=True, res=True, stats=True),
    "c256_320": lambda: run_case("c256_320", 1, 4, 8, 8, 256, 320, in_norm=True, stats=True),
    "c576_256": lambda: run_case("c576_256", 1, 4, 16, 16, 576, 256, in_norm=True),
    "fp32": lambda: run_case("fp32", 1, 8, 32, 16, 32, 32, dtype=torch.float32, in_norm=True, res=True, stats=True),
    "lrelu": lambda: run_case("lrelu", 1, 8, 32, 16, 32, 32, in_norm=True, slope=0.01),
    "slice": lambda: run_case("slice", 1, 8, 32, 16, 32, 32, in_norm=True, x_extra=64, y_extra=64),
    "dgrad": lambda: run_dgrad_case("dgrad", 2, 8, 32, 16, 32, 64),
    "perf32": lambda: run_case("perf32", 2, 128, 128, 128, 32, 32, in_norm=True, res=True, stats=True, iters=5),
    "perf32_pz2": lambda: run_case("perf32_pz2", 2, 128, 128, 128, 32, 32, in_norm=True, res=True, stats=True, iters=5, pz=2),
    "perf96_64": lambda: run_case("perf96_64", 2, 128, 128, 128, 96, 64, in_norm=True, stats=True, iters=5),
    "perf64": lambda: run_case("perf64", 2, 64, 64, 64, 64, 64, in_norm=True, res=True, stats=True, iters=5),
    "perf128": lambda: run_case("perf128", 2, 32, 32, 32, 128, 128, in_norm=True, res=True, stats=True, iters=5),
    "perf256": lambda: run_case("perf256", 2, 16, 16, 16, 256, 256, in_norm=True, res=True, stats=True, iters=5),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        CASES[nm]()
