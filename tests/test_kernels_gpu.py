"""GPU parity tests of every kernel family, called through the C-ABI (ctypes), against the oracle /
plain torch fp32 references on the same seeded inputs.  Tolerances are written next to each check:
bit-exact for byte/index work, operand-rounding-aware bounds for the bf16 tensor-core paths."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def stats_of(x_cl, pitch=None, c0=0):
    """(sum, sumsq) laid out with the activation's channel pitch."""
    n, c = x_cl.shape[0], x_cl.shape[4]
    v = x_cl.reshape(n, -1, c).double()
    st = torch.stack([v.sum(1), (v * v).sum(1)], dim=-1).float()
    if pitch is None:
        return st
    full = torch.zeros(n, pitch, 2, device=x_cl.device)
    full[:, c0:c0 + c] = st
    return full[:, c0:c0 + c]


def cl(t):  # NCDHW -> NDHWC
    return t.permute(0, 2, 3, 4, 1).contiguous()


def nc(t):  # NDHWC -> NCDHW
    return t.permute(0, 4, 1, 2, 3).contiguous()


def act_ref(x_nc, slope):
    h = F.instance_norm(x_nc, eps=1e-4)
    return F.leaky_relu(h, slope) if slope else F.relu(h)


CONV_CASES = [
    # N, D, H, W, Cin, Cout, dtype, norm, res, slope, pz
    (1, 4, 16, 8, 16, 16, torch.bfloat16, False, False, 0.0, 1),
    (2, 8, 32, 16, 32, 32, torch.bfloat16, True, True, 0.0, 4),
    (1, 5, 20, 12, 24, 40, torch.bfloat16, True, False, 0.0, 0),   # ragged tiles, padded channels
    (1, 8, 16, 16, 96, 64, torch.bfloat16, True, False, 0.01, 2),  # LeakyReLU variant
    (1, 4, 8, 8, 256, 320, torch.bfloat16, True, True, 0.0, 0),
    (1, 2, 16, 16, 576, 256, torch.bfloat16, True, False, 0.0, 0),
    (1, 8, 32, 16, 32, 32, torch.float32, True, True, 0.0, 0),
    (1, 3, 7, 5, 8, 8, torch.float32, True, True, 0.0, 0),          # smaller than one tile
    (2, 6, 16, 16, 64, 96, torch.bfloat16, True, True, 0.0, 0),     # D not a multiple of PZ, N tile 96
    (1, 9, 24, 24, 32, 192, torch.bfloat16, True, False, 0.0, 2),   # two N tiles of 96, ragged z block
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3_forward(cuda_dev, case):
    from rsuper_b200 import ops
    N, D, H, W, Cin, Cout, dt, norm, res, slope, pz = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(cuda_dev).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(cuda_dev)
    r = torch.randn(N, D, H, W, Cout, generator=g).to(cuda_dev).to(dt) if res else None
    y = torch.zeros(N, D, H, W, Cout, dtype=dt, device=cuda_dev)
    ost = torch.zeros(N, Cout, 2, device=cuda_dev)
    a_op = ops.norm_act(x, stats_of(x.float()) if norm else None, slope=slope)
    ops.conv3_forward(a_op, ops.conv3_pack_weights(w), y, res=r, out_stats=ost, planes_per_item=pz)
    # identical bf16 operands (the kernel's own operand tensor), fp32 accumulation in TMEM
    ref = cl(F.conv3d(nc(a_op.float()), bf16r(w), padding=1))
    if res:
        ref = ref + r.float()
    # fp32 storage: only accumulation-order noise; bf16 storage: one bf16 rounding of the output (2^-9)
    tol = 2e-5 if dt == torch.float32 else 4e-3
    assert rel(y.float(), ref) <= tol
    assert rel(ost, stats_of(ref)) <= 1e-4  # statistics come from the fp32 accumulators
    # and the operand itself is bf16(act(instance_norm(x))) up to single-ulp rounding flips
    a_ref = act_ref(nc(x.float()), slope) if norm else nc(x.float())
    assert rel(a_op.float(), cl(a_ref)) <= 2.0 ** -8


def test_conv3_split_precision_matches_fp32(cuda_dev):
    """3-pass split product (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo) with fp32 storage: fp32-level parity with
    F.conv3d on UNROUNDED operands (north-star bar for logits is 1e-3 relative; a single conv is ~1e-5)."""
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(21)
    for (N, D, H, W, Cin, Cout) in [(1, 6, 16, 16, 32, 32), (1, 4, 16, 8, 96, 64), (1, 3, 8, 8, 40, 136)]:
        x = torch.randn(N, D, H, W, Cin, generator=g).to(cuda_dev)
        w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(cuda_dev)
        y = torch.zeros(N, D, H, W, Cout, device=cuda_dev)
        hi, lo = ops.norm_act(x, stats_of(x), split=True)
        ops.conv3_forward(hi, ops.conv3_pack_weights(w, split=True), y, a_lo=lo)
        ref = cl(F.conv3d(act_ref(nc(x), 0.0), w, padding=1))
        assert rel(y, ref) <= 3e-5
        # three pieces per operand, six products: the operands are then exact to ~2^-24, but the result is not better —
        # the floor (~1e-5 of the output range) is the tensor pipe's truncating fp32 accumulation, which is why
        # precision='fp32' stops at two pieces
        hi, lo, lo2 = ops.norm_act(x, stats_of(x), split=3)
        assert rel((hi.double() + lo.double() + lo2.double()).float(), cl(act_ref(nc(x), 0.0))) <= 2e-6
        ops.conv3_forward(hi, ops.conv3_pack_weights(w, split=6), y, a_lo=lo, a_lo2=lo2)
        ref64 = cl(F.conv3d(act_ref(nc(x), 0.0).double(), w.double(), padding=1))
        assert rel(y.double(), ref64) <= 3e-5


def test_conv3_channel_slices_and_stats_pitch(cuda_dev):
    """Reads a channel slice of a wider buffer and writes into a slice (free torch.cat)."""
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(2)
    xb = torch.randn(1, 4, 16, 16, 96, generator=g).to(cuda_dev).to(torch.bfloat16)
    yb = torch.zeros(1, 4, 16, 16, 64, dtype=torch.bfloat16, device=cuda_dev)
    ab = torch.zeros_like(xb)
    x, y = xb[..., 32:64], yb[..., 16:48]
    w = (torch.randn(32, 32, 3, 3, 3, generator=g) / 30).to(cuda_dev)
    ist = stats_of(x.float(), pitch=96, c0=32)
    ostb = torch.zeros(1, 64, 2, device=cuda_dev)
    a_op = ops.norm_act(x, ist, out=ab[..., 32:64])   # operand written into (and TMA-read from) a channel slice
    ops.conv3_forward(a_op, ops.conv3_pack_weights(w), y, out_stats=ostb[:, 16:48])
    ref = cl(F.conv3d(nc(a_op.float()), bf16r(w), padding=1))
    assert rel(y.float(), ref) <= 4e-3
    assert yb[..., :16].abs().max() == 0 and yb[..., 48:].abs().max() == 0
    assert ab[..., :32].abs().max() == 0 and ab[..., 64:].abs().max() == 0
    assert rel(ostb[:, 16:48], stats_of(ref)) <= 1e-4
    assert ostb[:, :16].abs().max() == 0 and ostb[:, 48:].abs().max() == 0


@pytest.mark.parametrize("shape", [(2, 8, 32, 16, 32, 64), (1, 4, 16, 8, 64, 32), (1, 3, 7, 5, 16, 8), (1, 4, 16, 16, 96, 64)])
def test_conv3_dgrad_with_norm_backward_sums(cuda_dev, shape):
    from rsuper_b200 import ops
    N, D, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(cuda_dev)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(cuda_dev)
    dy = torch.randn(N, D, H, W, Cout, generator=g).to(cuda_dev)
    xhat = F.instance_norm(nc(x), eps=1e-4)
    da = F.conv_transpose3d(bf16r(nc(dy)), bf16r(w), padding=1)
    gref = torch.where(xhat > 0, da, torch.zeros_like(da))
    gout = torch.zeros(N, D, H, W, Cin, device=cuda_dev)
    sums = torch.zeros(N, Cin, 2, device=cuda_dev)
    ops.conv3_forward(ops.norm_act(dy), ops.conv3_pack_weights(w, True), gout, mask_x=x, mask_stats=stats_of(x), bwd_sums=sums)
    assert rel(gout, cl(gref)) <= 2e-5
    assert rel(sums[..., 0], gref.sum(dim=(2, 3, 4))) <= 1e-4
    assert rel(sums[..., 1], (gref * xhat).sum(dim=(2, 3, 4))) <= 1e-4


WG_CASES = [(2, 8, 32, 16, 32, 32), (1, 6, 16, 8, 96, 32), (1, 4, 16, 16, 64, 64), (1, 4, 8, 8, 128, 128),
            (1, 2, 8, 8, 256, 320), (1, 5, 20, 12, 24, 40), (1, 3, 7, 5, 8, 16), (1, 4, 16, 16, 192, 128)]


@pytest.mark.parametrize("shape", WG_CASES)
def test_conv3_wgrad(cuda_dev, shape):
    from rsuper_b200 import ops
    N, D, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, D, H, W, Cin, generator=g).to(cuda_dev)
    dy = torch.randn(N, D, H, W, Cout, generator=g).to(cuda_dev).to(torch.bfloat16)
    a_op = ops.norm_act(x, stats_of(x))
    # the operand producer itself: bf16(relu(instance_norm(x))) up to rare rounding flips (1e-7 differences in
    # the normalisation crossing a bf16 rounding boundary = one bf16 ulp)
    a_ref = act_ref(nc(x), 0.0)
    assert rel(cl(a_ref), a_op.float()) <= 2.0 ** -8
    wz = torch.zeros(Cout, Cin, 3, 3, 3, device=cuda_dev, requires_grad=True)
    (F.conv3d(nc(a_op.float()), wz, padding=1) * nc(dy.float())).sum().backward()
    dw = torch.full((Cout, Cin, 3, 3, 3), 7.0, device=cuda_dev)
    ops.conv3_wgrad(a_op, dy, dw)
    # identical bf16 operands, fp32 accumulation: only summation-order noise remains
    assert rel(dw, wz.grad) <= 2e-5
    ops.conv3_wgrad(a_op, dy, dw, accumulate=True)
    assert rel(dw, 2 * wz.grad) <= 2e-5


def test_norm_act_split_and_slices(cuda_dev):
    """hi + lo reproduces the fp32 operand to ~2^-17; channel-slice views (pitch != C) work; LeakyReLU slope."""
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(14)
    xb = torch.randn(2, 4, 8, 8, 48, generator=g).to(cuda_dev)
    x = xb[..., 16:40]
    st = stats_of(x, pitch=48, c0=16)
    hi, lo = ops.norm_act(x, st, slope=0.01, split=True)
    ref = cl(act_ref(nc(x.contiguous()), 0.01))
    assert rel(hi.float() + lo.float(), ref) <= 2.0 ** -15
    assert rel(hi.float(), ref) <= 2.0 ** -8
    c16 = ops.norm_act(x.to(torch.bfloat16))      # plain cast of a bf16 view is the identity
    assert torch.equal(c16, x.to(torch.bfloat16))


def test_stem_and_head(cuda_dev):
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(5)
    N, D, H, W, Co, C = 2, 12, 20, 24, 32, 3
    x = torch.randn(N, 1, D, H, W, generator=g).to(cuda_dev)
    w = (torch.randn(Co, 1, 3, 3, 3, generator=g) / 5).to(cuda_dev)
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 4e-3)):
        y = torch.zeros(N, D, H, W, Co, dtype=dt, device=cuda_dev)
        st = torch.zeros(N, Co, 2, device=cuda_dev)
        ops.stem_conv_forward(x, w, y, st)
        ref = cl(F.conv3d(x, w, padding=1))
        assert rel(y.float(), ref) <= tol
        assert rel(st, stats_of(ref)) <= 1e-4
    dy = torch.randn(N, D, H, W, Co, generator=g).to(cuda_dev)
    wz = w.clone().requires_grad_(True)
    (F.conv3d(x, wz, padding=1) * nc(dy)).sum().backward()
    dw = torch.zeros_like(w)
    ops.stem_conv_wgrad(x, dy, dw)
    assert rel(dw, wz.grad) <= 1e-4
    # head: C <= 8 takes the coalesced (voxel, channel-group) kernels, C = 10 the one-voxel-per-thread ones
    for C in (3, 10):
        f = torch.randn(N, D, H, W, Co, generator=g).to(cuda_dev)
        hw = (torch.randn(C, Co, generator=g) / 6).to(cuda_dev).requires_grad_(True)
        hb = torch.randn(C, generator=g).to(cuda_dev).requires_grad_(True)
        fz = f.clone().requires_grad_(True)
        ref = F.conv3d(nc(fz), hw[:, :, None, None, None], hb)
        logits = torch.zeros(N, C, D, H, W, device=cuda_dev)
        ops.head_forward(f, hw.detach(), hb.detach(), logits)
        assert rel(logits, ref) <= 1e-5
        dl = torch.randn(N, C, D, H, W, generator=g).to(cuda_dev)
        ref.backward(dl)
        dx = torch.zeros_like(f)
        dw_, db_ = torch.zeros(C, Co, device=cuda_dev), torch.zeros(C, device=cuda_dev)
        ops.head_backward(f, hw.detach(), dl, dx, dw_, db_)
        assert rel(dx, fz.grad) <= 1e-5
        assert rel(dw_, hw.grad) <= 1e-4 and rel(db_, hb.grad) <= 1e-4


def test_maxpool_forward_backward_with_ties(cuda_dev):
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(6)
    N, D, H, W, C = 2, 8, 12, 16, 24
    # coarse values => many exact ties inside pooling windows: first maximum (d,h,w order) must win
    x = torch.randint(-2, 3, (N, D, H, W, C), generator=g).float().to(cuda_dev)
    y = torch.zeros(N, D // 2, H // 2, W // 2, C, device=cuda_dev)
    st = torch.zeros(N, C, 2, device=cuda_dev)
    ops.maxpool2_forward(x, y, st)
    xz = nc(x).requires_grad_(True)
    ref = F.max_pool3d(xz, 2)
    assert torch.equal(y, cl(ref))
    assert rel(st, stats_of(cl(ref))) <= 1e-5
    dy = torch.randn(N, D // 2, H // 2, W // 2, C, generator=g).to(cuda_dev)
    dskip = torch.randn(N, D, H, W, C, generator=g).to(cuda_dev)
    ref.backward(nc(dy))
    dx = torch.zeros_like(x)
    ops.maxpool2_backward(x, dy, dx, dskip=dskip)
    assert torch.equal(dx, cl(xz.grad) + dskip)  # bit-exact routing


@pytest.mark.parametrize("dims", [((4, 4, 4), (8, 8, 8)), ((2, 2, 2), (4, 4, 4)), ((8, 6, 4), (16, 12, 8)),
                                  ((5, 4, 3), (9, 8, 7)), ((32, 16, 16), (64, 32, 32)), ((1, 3, 3), (2, 6, 6))])
def test_upsample_trilinear(cuda_dev, dims):
    from rsuper_b200 import ops
    (di, hi, wi), (do, ho, wo) = dims
    g = torch.Generator().manual_seed(7)
    N, C = 2, 16
    x = torch.randn(N, di, hi, wi, C, generator=g).to(cuda_dev)
    xz = nc(x).requires_grad_(True)
    ref = F.interpolate(xz, size=(do, ho, wo), mode="trilinear", align_corners=True)
    y = torch.zeros(N, do, ho, wo, C, device=cuda_dev)
    st = torch.zeros(N, C, 2, device=cuda_dev)
    ops.upsample_forward(x, y, st)
    assert rel(y, cl(ref)) <= 1e-5
    assert rel(st, stats_of(cl(ref))) <= 1e-4
    dy = torch.randn(N, do, ho, wo, C, generator=g).to(cuda_dev)
    ref.backward(nc(dy))
    for two_pass in (True, False):   # z-adjoint into a scratch + (y, x) gather  |  single-pass 3-D gather
        dx = torch.full_like(x, float("nan"))
        ops.upsample_backward(dy, dx, two_pass=two_pass)
        assert rel(dx, cl(xz.grad)) <= 1e-5


def test_instnorm_backward_apply(cuda_dev):
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(8)
    N, D, H, W, C = 2, 6, 10, 12, 16
    x = torch.randn(N, D, H, W, C, generator=g).to(cuda_dev) * 2 + 0.5
    gup = torch.randn(N, D, H, W, C, generator=g).to(cuda_dev)
    add = torch.randn(N, D, H, W, C, generator=g).to(cuda_dev)
    xz = nc(x).requires_grad_(True)
    xhat = F.instance_norm(xz, eps=1e-4)
    xhat.backward(nc(gup))
    sums = torch.stack([nc(gup).sum(dim=(2, 3, 4)), (nc(gup) * xhat.detach()).sum(dim=(2, 3, 4))], dim=-1).contiguous()
    dx = torch.zeros_like(x)
    ops.instnorm_backward_apply(gup, x, stats_of(x), sums, dx, add=add)
    assert rel(dx, cl(xz.grad) + add) <= 2e-5


def test_layout_and_stats(cuda_dev):
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(9)
    src = torch.randn(2, 24, 5, 6, 7, generator=g).to(cuda_dev)
    for dt in (torch.float32, torch.bfloat16):
        dst = torch.zeros(2, 5, 6, 7, 24, dtype=dt, device=cuda_dev)
        ops.ncdhw_to_ndhwc(src, dst)
        assert torch.equal(dst.float(), cl(src).to(dt).float())
        back = torch.zeros_like(src)
        ops.ndhwc_to_ncdhw(dst, back)
        assert torch.equal(back, src.to(dt).float())
        st = ops.channel_stats(dst)
        assert rel(st, stats_of(dst.float())) <= 1e-5


def test_seg_loss_forward_backward_vs_oracle(cuda_dev):
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import ops
    shp = (16, 24, 32)
    cls3 = ["liver", "liver_lesion", "pancreas"]
    lg = synth.synthetic_logits(2, 3, shp, seed=2, device=cuda_dev)
    b3 = synth.make_batch(["mask", "report"], cls3, shp, seed=11, device=cuda_dev)
    known = LR.get_known_voxels(b3["unk_channels"].float())
    for cw in (None, torch.tensor([[1.0, 2.0, 0.5], [0.25, 1.0, 3.0]], device=cuda_dev)):
        lz = lg.clone().requires_grad_(True)
        ref = LR.seg_loss(lz, b3["label"].float(), known, None if cw is None else cw[:, :, None, None, None])
        ref.backward()
        st = ops.seg_loss_forward(lg, b3["label"], known.to(torch.uint8), cw)
        assert abs(st.loss_out[0].item() - ref.item()) <= 1e-5  # north star: loss within 1e-5
        dl = torch.zeros_like(lg)
        ops.seg_loss_backward(st, torch.ones(2, device=cuda_dev), dl)
        assert rel(dl, lz.grad) <= 1e-4
    # known=None path (mask-only batches, train_ddp.py:268-271)
    lz = lg.clone().requires_grad_(True)
    ref = LR.seg_loss(lz, b3["label"].float(), torch.ones_like(lg))
    st = ops.seg_loss_forward(lg, b3["label"], None, None)
    assert abs(st.loss_out[0].item() - ref.item()) <= 1e-5


def test_seg_loss_matches_golden(cuda_dev, golden):
    """Against the REAL reference's recorded values (tests/golden)."""
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import ops
    shp = (16, 24, 32)
    lg = synth.synthetic_logits(2, 3, shp, seed=2, device=cuda_dev)
    b3 = synth.make_batch(["mask", "report"], ["liver", "liver_lesion", "pancreas"], shp, seed=11, device=cuda_dev)
    known = ops.dilate_ball(b3["unk_channels"], 5).logical_not().to(torch.uint8)
    st = ops.seg_loss_forward(lg, b3["label"], known, None)
    assert abs(st.loss_out[1].item() - float(golden["seg_bce"])) <= 1e-5
    assert abs(st.loss_out[2].item() - float(golden["seg_dice"])) <= 1e-5
    dl = torch.zeros_like(lg)
    ops.seg_loss_backward(st, torch.ones(2, device=cuda_dev), dl)
    np.testing.assert_allclose(dl.cpu().numpy()[:, :, ::4, ::4, ::4], golden["seg_grad"], rtol=2e-3, atol=1e-9)


def test_dilation_bit_exact(cuda_dev, golden):
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import ops
    vol = synth.make_batch(["report", "mask"], ["organ", "pancreatic_lesion"], (24, 28, 20), seed=9,
                           device=cuda_dev)["unk_channels"]
    for k in (1, 3, 5, 7, 9, 15, 31):
        got = ops.dilate_ball(vol, k)
        assert np.array_equal(np.packbits(got.cpu().numpy().astype(bool).reshape(-1)), golden[f"dilate_{k}"]), k
        assert torch.equal(got.float(), LR.dilate_volume(vol.float(), k))


@pytest.mark.parametrize("split", [False, True])
def test_batched_weight_packing_matches_per_layer_packing(split):
    """ONE launch over a job table == the per-layer packer, bit for bit: plain, merged rows (conv1 || shortcut) and
    the flipped / transposed dgrad images."""
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(7)
    dev = "cuda"
    w1 = torch.randn(40, 24, 3, 3, 3, generator=g).to(dev)
    wsc = torch.randn(40, 24, 3, 3, 3, generator=g).to(dev)
    w2 = torch.randn(96, 64, 3, 3, 3, generator=g).to(dev)
    w3 = torch.randn(320, 32, 3, 3, 3, generator=g).to(dev)
    jobs = [("a", w1, wsc, False), ("aT", w1, wsc, True), ("b", w2, None, False), ("bT", w2, None, True), ("c", w3, None, True)]
    plan = ops.PackPlan(jobs, split=split)
    for img in plan.images.values():
        img.fill_(0x5A)
    plan.refresh()
    wcat = torch.cat([w1, wsc], 0).contiguous()
    want = {"a": ops.conv3_pack_weights(wcat, False, split=split), "aT": ops.conv3_pack_weights(wcat, True, split=split),
            "b": ops.conv3_pack_weights(w2, False, split=split), "bT": ops.conv3_pack_weights(w2, True, split=split),
            "c": ops.conv3_pack_weights(w3, True, split=split)}
    for k, ref in want.items():
        assert plan.images[k].shape == ref.shape and torch.equal(plan.images[k], ref), k
    # values are re-read on refresh (the optimizer updates parameters in place)
    w2.mul_(0.5)
    plan.refresh()
    assert torch.equal(plan.images["b"], ops.conv3_pack_weights(w2, False, split=split))
    assert plan.ptr_key == ops.PackPlan.pointer_key([(w1, wsc), (w1, wsc), (w2, None), (w2, None), (w3, None)])


STREAM_CASES = [
    # N, D, H, W, Cin, Cout, storage dtype, mode ('res' | 'plain' | 'mask'), CTAs
    (2, 40, 20, 12, 32, 32, torch.bfloat16, "res", 4),     # ragged last z-chunk (32 + 8), ragged y / x tiles, sample change inside a CTA
    (1, 70, 16, 8, 32, 32, torch.bfloat16, "plain", 1),    # one CTA walks 3 chunks: the 16-slot accumulator ring wraps inside units
    (1, 33, 16, 16, 24, 24, torch.float32, "res", 3),      # padded channels (Cout 24 -> N tile 32), fp32 storage, a 1-plane chunk
    (2, 32, 16, 16, 16, 32, torch.bfloat16, "mask", 5),    # dgrad epilogue, Cin = 16 (one k-step)
    (1, 64, 32, 16, 32, 32, torch.float32, "mask", 2),
]


@pytest.mark.parametrize("case", STREAM_CASES)
def test_conv3_plane_streaming_kernel(cuda_dev, case, monkeypatch):
    """The plane-streaming kernel of the 32-channel layers (conv3_stream.cu) against F.conv3d on identical bf16
    operands; max_ctas forces it at test sizes (it needs >= 3 units per CTA)."""
    from rsuper_b200 import ops
    monkeypatch.delenv("RSB_FPROP_STREAM", raising=False)   # the default for eligible layers (RSB_FPROP_STREAM=0 disables it)
    N, D, H, W, Cin, Cout, dt, mode, ctas = case
    g = torch.Generator().manual_seed(31)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(cuda_dev)
    if mode != "mask":
        x = torch.randn(N, D, H, W, Cin, generator=g).to(cuda_dev).to(dt)
        r = torch.randn(N, D, H, W, Cout, generator=g).to(cuda_dev).to(dt) if mode == "res" else None
        y = torch.full((N, D, H, W, Cout), float("nan"), dtype=dt, device=cuda_dev)
        ost = torch.zeros(N, Cout, 2, device=cuda_dev)
        a_op = ops.norm_act(x, stats_of(x.float()))
        ops.conv3_forward(a_op, ops.conv3_pack_weights(w), y, res=r, out_stats=ost, max_ctas=ctas)
        ref = cl(F.conv3d(nc(a_op.float()), bf16r(w), padding=1))
        if r is not None:
            ref = ref + r.float()
        assert rel(y.float(), ref) <= (2e-5 if dt == torch.float32 else 4e-3)
        assert rel(ost, stats_of(ref)) <= 1e-4
        # and it is the same result as the item-based kernel (planes_per_item given => not eligible for streaming)
        y2 = torch.zeros_like(y)
        ops.conv3_forward(a_op, ops.conv3_pack_weights(w), y2, res=r, planes_per_item=2)
        assert rel(y.float(), y2.float()) <= (2e-5 if dt == torch.float32 else 4e-3)
    else:
        # dgrad: effective conv has Cin' = Cout, Cout' = Cin
        x = torch.randn(N, D, H, W, Cin, generator=g).to(cuda_dev).to(dt)
        dy = torch.randn(N, D, H, W, Cout, generator=g).to(cuda_dev)
        xhat = F.instance_norm(nc(x.float()), eps=1e-4)
        da = F.conv_transpose3d(bf16r(nc(dy)), bf16r(w), padding=1)
        gref = torch.where(xhat > 0, da, torch.zeros_like(da))
        gout = torch.full((N, D, H, W, Cin), float("nan"), dtype=dt, device=cuda_dev)
        sums = torch.zeros(N, Cin, 2, device=cuda_dev)
        ops.conv3_forward(ops.norm_act(dy), ops.conv3_pack_weights(w, True), gout, mask_x=x, mask_stats=stats_of(x.float()), bwd_sums=sums,
                          max_ctas=ctas)
        if dt == torch.float32:
            assert rel(gout, cl(gref)) <= 2e-5
        else:
            # bf16 x: a handful of voxels sit within rounding of the ReLU threshold; compare where |xhat| is clear of it
            clear = (xhat.abs() > 1e-2)
            assert rel(torch.where(clear, nc(gout.float()), gref), gref) <= 4e-3
        assert rel(sums[..., 0], gref.sum(dim=(2, 3, 4))) <= 2e-3
        assert rel(sums[..., 1], (gref * xhat).sum(dim=(2, 3, 4))) <= 2e-3
