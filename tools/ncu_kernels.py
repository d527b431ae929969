"""One launch of every kernel family of the train step at its benchmark shape (BASELINE.json configs[1]: 2 x 128^3, base 32),
as the target of ONE `ncu --set full` capture (each launch is replayed ~40 times by ncu, so nothing is called twice):

    ncu --set full --clock-control none --import-source on -o gpurun_out/r02_kernels python tools/ncu_kernels.py
    python tools/summarize_ncu.py roofline gpurun_out/r02_kernels.ncu-rep profiles/r02_kernels_ncu_summary.txt
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from rsuper_b200 import ops
dev = "cuda"
N, S, C = 2, 128, 32
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)


def rnd(*shape, dtype=bf):
    return torch.randn(*shape, generator=g).to(dev).to(dtype)


which = set(sys.argv[1:])


def want(tag):
    return not which or tag in which


x32 = rnd(N, S, S, S, C)
stx = ops.channel_stats(x32)
dx32 = torch.empty_like(x32)
dsk = rnd(N, S, S, S, C)
st = torch.zeros(N, C, 2, device=dev)
if want("sat"):
    # ---- HBM-bound satellites (level 0) ----
    ops.norm_act(x32, stx)
    sums = torch.zeros(N, C, 2, device=dev)
    ops.instnorm_backward_apply(dsk, x32, stx, sums, dx32, add=x32)
    y16 = torch.empty(N, S // 2, S // 2, S // 2, C, dtype=bf, device=dev)
    ops.maxpool2_forward(x32, y16, st)
    ops.maxpool2_backward(x32, rnd(N, S // 2, S // 2, S // 2, C), dx32, dskip=dsk)
    x64 = rnd(N, S // 2, S // 2, S // 2, 64)
    cat = torch.empty(N, S, S, S, 96, dtype=bf, device=dev)
    stc = torch.zeros(N, 96, 2, device=dev)
    ops.upsample_forward(x64, cat[..., 32:], stc[:, 32:])
    dcat = rnd(N, S, S, S, 96)
    ops.upsample_backward(dcat[..., 32:], torch.empty_like(x64))
    img = torch.randn(N, 1, S, S, S, generator=g).to(dev)
    w0 = (torch.randn(C, 1, 3, 3, 3, generator=g) / 5).to(dev)
    ops.stem_conv_forward(img, w0, torch.empty(N, S, S, S, C, dtype=bf, device=dev), st)
    ops.stem_conv_wgrad(img, x32, torch.empty_like(w0))
    logits = torch.empty(N, 2, S, S, S, device=dev)
    hw, hb = torch.randn(2, C, generator=g).to(dev), torch.zeros(2, device=dev)
    ops.head_forward(x32, hw, hb, logits)
    ops.head_backward(x32, hw, torch.randn(N, 2, S, S, S, generator=g).to(dev), dx32, torch.empty_like(hw), torch.empty_like(hb))
    # ---- losses ----
    lab = (torch.rand(N, 2, S, S, S, generator=g) < 0.2).to(torch.uint8).to(dev)
    lg = torch.randn(N, 2, S, S, S, generator=g).to(dev)
    sl = ops.seg_loss_forward(lg, lab)
    ops.seg_loss_backward(sl, torch.ones(2, device=dev), torch.empty_like(lg))
    ops.dilate_ball(lab[:, 1].contiguous(), 7)
    prob = torch.rand(S, S, S, generator=g).to(dev) * (torch.rand(S, S, S, generator=g) < 0.3).to(dev)
    from rsuper_b200 import report_losses as RL
    g1d, wtab, reach, g_host, w_host = RL._gauss_ball_sep(31, 1.5, torch.device(dev))
    ops.ball_correlate_argmax_sep(prob.contiguous(), g1d, wtab, reach, g_host, w_host)
if want("opt"):
    # ---- optimizer end: clip + AdamW + EMA over the 40.56 M parameters of the base-32 UNet; weight packing ----
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.unet import B200UNet, _Engine
    net = B200UNet(1, 32, num_classes=2).to(dev)
    params = list(net.parameters())
    for p in params:
        p.grad = torch.randn_like(p) * 1e-3
    opt = B200AdamW(params, lr=6e-4, weight_decay=0.05, max_norm=1.0, ema_params=[p.detach().clone() for p in params])
    opt.step()
    eng = _Engine(32, 0.0, bf)
    eng.prepare({k: v.detach() for k, v in net.named_parameters()})
if want("conv"):
    # ---- tensor-core kernels ----
    w = (torch.randn(32, 32, 3, 3, 3, generator=g) / 30).to(dev)
    wp = ops.conv3_pack_weights(w)
    y = torch.empty_like(x32)
    ops.conv3_forward(x32, wp, y, out_stats=st)                                   # plane-streaming kernel: fprop + statistics
    ops.conv3_forward(x32, wp, y, res=dsk, out_stats=st)                          # ... + residual
    ops.conv3_forward(x32, ops.conv3_pack_weights(w, True), y, mask_x=dsk, mask_stats=stx, bwd_sums=torch.zeros(N, C, 2, device=dev))  # dgrad
    ops.conv3_forward(x32, wp, y, out_stats=st, planes_per_item=4)                # item-based kernel on the same layer
    a96 = rnd(N, S, S, S, 96)
    w96 = (torch.randn(64, 96, 3, 3, 3, generator=g) / 50).to(dev)
    y64 = torch.empty(N, S, S, S, 64, dtype=bf, device=dev)
    ops.conv3_forward(a96, ops.conv3_pack_weights(w96), y64, out_stats=torch.zeros(N, 64, 2, device=dev))
    ops.conv3_wgrad(x32, dsk, torch.empty(32, 32, 3, 3, 3, device=dev))
    ops.conv3_wgrad(a96, y64, torch.empty(64, 96, 3, 3, 3, device=dev))
    a64 = rnd(N, 64, 64, 64, 64)
    w64 = (torch.randn(64, 64, 3, 3, 3, generator=g) / 40).to(dev)
    ops.conv3_forward(a64, ops.conv3_pack_weights(w64), torch.empty_like(a64), out_stats=torch.zeros(N, 64, 2, device=dev))
torch.cuda.synchronize()
print("done")

if "mf" in which:
    # ---- MedFormer voxel-side kernels (csrc/medformer.cu) at the shapes of the yaml configuration, 1 x 128^3 ----
    a = rnd(1, 64, 64, 64, 256)                       # PatchMerging of down1: 8 * 32 channels at 64^3
    wdw = (torch.randn(256, 1, 3, 3, 3, generator=g) / 5).to(dev)
    ops.dwconv3(a, wdw)
    ops.dwconv3_wgrad(a, rnd(1, 64, 64, 64, 256))
    h = rnd(1, 32, 32, 32, 512)                       # MBConv of down2 / up2: 4 * 128 channels at 32^3
    wd2 = (torch.randn(512, 1, 3, 3, 3, generator=g) / 5).to(dev)
    ops.dwconv3(h, wd2)
    ops.dwconv3_wgrad(h, rnd(1, 32, 32, 32, 512))
    s = torch.rand(1, 512, generator=g).to(dev)
    ops.scale_channels(h, s)
    ops.channel_dot(h, h)
    qv = rnd(1, 32, 32, 32, 256)                      # attention of down2 / up2: 128 channels, 4 heads of 32
    mq = torch.randn(1, 4, 27, 32, generator=g).to(dev)
    mv = torch.randn(1, 4, 27, 32, generator=g).to(dev)
    fo, mo, ms = ops.biattention_forward(qv, mq, mv, 4)
    ops.biattention_backward(qv, mq, mv, ms, rnd(1, 32, 32, 32, 128), torch.randn(1, 4, 27, 32, generator=g).to(dev),
                             torch.randn(1, 4, 27, generator=g).to(dev), 4)
    feat, logit = rnd(1, 32, 32, 32, 128), rnd(1, 32, 32, 32, 32)
    smap, ms2 = ops.softmax_pool_forward(feat, logit, 27)
    ops.softmax_pool_backward(feat, logit, ms2, torch.randn(1, 128, 27, generator=g).to(dev), torch.randn(1, 27, generator=g).to(dev))
    torch.cuda.synchronize()
