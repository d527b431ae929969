"""CPU host-logic tests: the product's Python side (engine, losses, optimizer, inference, batch assembly) executed end to end
in the GPU-less container with every kernel launch replaced by a recorder (tests/dryrun.py).  What is checked is the launch
SEQUENCE the GPU box would receive — kernel names, counts, geometry and which optional pointers are set — for every block
type and precision mode, i.e. the branches that otherwise only run on hardware."""
from collections import Counter
from types import SimpleNamespace

import pytest
import torch

from dryrun import recording
from oracle import synth
from oracle.unet_ref import synthetic_image, synthetic_state_dict

CLASSES = ["organ", "pancreatic_lesion"]


def _step(rec, block, precision, base=8, side=32, batch=1):
    from rsuper_b200 import losses
    from rsuper_b200 import unet as U
    net = U.B200UNet(1, base, num_classes=2, block=block, precision=precision)
    net.load_state_dict(synthetic_state_dict(base, 2, block=block))
    names, params = zip(*net.named_parameters())
    eng = U._Engine(base, 0.0, torch.bfloat16 if precision == "bf16" else torch.float32, block)
    x = synthetic_image(batch, side, side, side, seed=1)
    out = U._UNetFunction.apply(x, eng, names, 2, True, *params)       # B200UNet.forward minus its CUDA check
    n_fwd = len(rec.calls)
    lab = synth.make_batch(["mask"] * batch, CLASSES, (side,) * 3, seed=2)["label"]
    losses.seg_loss(out, lab).backward()
    return net, out, n_fwd


def _convs(calls, name):
    return [a[0] for n, a in calls if n == name]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_basicblock_train_step_launch_sequence(monkeypatch, precision):
    with recording(monkeypatch) as rec:
        net, out, n_fwd = _step(rec, "BasicBlock", precision)
    assert out.shape == (1, 2, 32, 32, 32) and out.dtype == torch.float32
    cnt = Counter(n for n, _ in rec.calls)
    # 34 tensor-core convs forward (18 BasicBlocks: conv1 || shortcut merged where present), as many dgrads and wgrads
    nw = 3 if precision == "fp32" else 1      # parity mode: three split products per weight gradient (hi*hi, then two accumulating)
    assert cnt["rsb_conv3_forward"] == 68 and cnt["rsb_conv3_wgrad"] == 34 * nw
    assert cnt["rsb_conv3_pack_weights_batched"] == 1 and cnt["rsb_stem_conv_forward"] == 1 and cnt["rsb_stem_conv_wgrad"] == 1
    assert cnt["rsb_head_forward"] == 1 and cnt["rsb_head_backward"] == 1
    assert cnt["rsb_maxpool2_forward"] == cnt["rsb_maxpool2_backward"] == 4
    assert cnt["rsb_upsample_trilinear_forward"] == cnt["rsb_upsample_trilinear_backward"] == 4
    assert cnt["rsb_seg_loss_forward"] == cnt["rsb_seg_loss_backward"] == 1
    fwd = _convs(rec.calls[:n_fwd], "rsb_conv3_forward")
    bwd = _convs(rec.calls[n_fwd:], "rsb_conv3_forward")
    wg = _convs(rec.calls, "rsb_conv3_wgrad")
    assert len(fwd) == len(bwd) == 34 and len(wg) == 34 * nw
    assert sum(a["accumulate"] == 0 for a in wg) == 34              # exactly one overwriting launch per weight tensor
    wg = [a for a in wg if a["accumulate"] == 0]
    # every forward conv has its data gradient (Cin/Cout swapped, same volume) and its weight gradient
    key = lambda a, swap=False: (a["N"], a["D"], a["H"], a["W"]) + ((a["Cout"], a["Cin"]) if swap else (a["Cin"], a["Cout"]))
    assert sorted(key(a) for a in fwd) == sorted(key(a, swap=True) for a in bwd) == sorted(key(a) for a in wg)
    for a in fwd:
        assert a["a"] == "p" and a["w_packed"] == "p" and a["y"] == "p" and a["mask_x"] is None and a["bwd_sums"] is None
        assert (a["a_lo"] == "p") == (precision == "fp32")          # split-precision operands only in the parity mode
        assert a["dtype"] == (0 if precision == "bf16" else 1) and abs(a["eps"] - 1e-4) < 1e-10 and a["slope"] == 0.0
        assert a["D"] % 2 == 0 and a["Cin"] % 8 == 0 and a["Cout"] % 8 == 0
    # dgrads into a pre-activation carry the act'/InstanceNorm-backward epilogue
    assert sum(a["mask_x"] == "p" and a["bwd_sums"] == "p" for a in bwd) >= 30
    # total forward FLOPs of the recorded convs == the closed form of SURVEY §8d for this net (base 8 at 32^3)
    flops = sum(2 * 27 * a["Cin"] * a["Cout"] * a["N"] * a["D"] * a["H"] * a["W"] for a in fwd)
    assert flops > 0 and all(p.grad is not None and p.grad.shape == p.shape for p in net.parameters())


@pytest.mark.parametrize("block", ["BasicBlock", "SingleConv"])
def test_pack_plan_is_built_once_across_forwards(monkeypatch, block):
    """The packed-weight plan (images + device job table, uploaded with a host->device copy) is keyed on the parameter
    storage: a second forward with the same parameters must reuse it (a rebuild per step is a hidden H2D copy that also
    breaks CUDA-graph capture), a re-allocated parameter must rebuild it."""
    from rsuper_b200 import ops
    from rsuper_b200 import unet as U
    built = []
    real_init = ops.PackPlan.__init__
    monkeypatch.setattr(ops.PackPlan, "__init__", lambda self, *a, **k: (built.append(1), real_init(self, *a, **k))[1])
    with recording(monkeypatch):
        net = U.B200UNet(1, 8, num_classes=2, block=block)
        names, params = zip(*net.named_parameters())
        eng = U._Engine(8, 0.0, torch.bfloat16, block)
        x = synthetic_image(1, 32, 32, 32, seed=1)
        for _ in range(3):
            U._UNetFunction.apply(x, eng, names, 2, True, *params)
        assert len(built) == 1
        w = dict(net.named_parameters())["down2.conv.1.conv1.conv.weight" if block == "BasicBlock" else "down2.conv.1.conv.conv.weight"]
        w.data = w.data.clone()                                  # re-allocated storage (e.g. load_state_dict(assign=True))
        names, params = zip(*net.named_parameters())
        U._UNetFunction.apply(x, eng, names, 2, True, *params)
        assert len(built) == 2


def test_singleconv_train_step_launch_sequence(monkeypatch):
    with recording(monkeypatch) as rec:
        net, out, n_fwd = _step(rec, "SingleConv", "bf16")
    cnt = Counter(n for n, _ in rec.calls)
    n_convs = sum(1 for k, _ in net.named_parameters() if k.endswith("conv.conv.weight"))
    assert len(_convs(rec.calls[:n_fwd], "rsb_conv3_forward")) == n_convs == cnt["rsb_conv3_wgrad"]
    assert cnt["rsb_act_backward_stats"] >= n_convs - 1
    assert all(p.grad is not None for p in net.parameters())


def test_flop_count_of_the_recorded_convs_matches_survey(monkeypatch):
    """the base-32 UNet costs 1.235 MFLOP per voxel forward, independent of the patch size — recomputed from the
    launches the engine actually issues (plus the CUDA-core stem and head)."""
    with recording(monkeypatch) as rec:
        _step(rec, "BasicBlock", "bf16", base=32, side=32)
        fwd = _convs(rec.calls, "rsb_conv3_forward")[:34]
    flops = sum(2 * 27 * a["Cin"] * a["Cout"] * a["N"] * a["D"] * a["H"] * a["W"] for a in fwd)
    vox = 32 ** 3
    flops += 2 * 27 * 1 * 32 * vox + 2 * 32 * 2 * vox          # stem 1 -> 32 and the 1x1x1 head 32 -> 2
    assert abs(flops / vox / 1e6 - 1.235) < 0.002


def test_fused_optimizer_launch_and_state(monkeypatch):
    from rsuper_b200.optim import B200AdamW
    with recording(monkeypatch) as rec:
        ps = [torch.nn.Parameter(torch.zeros(5000)), torch.nn.Parameter(torch.zeros(3, 7)), torch.nn.Parameter(torch.zeros(0))]
        ema = [p.detach().clone() for p in ps]
        opt = B200AdamW(ps, lr=6e-4, weight_decay=0.05, max_norm=1.0, ema_params=ema)
        for _ in range(3):
            for p in ps:
                p.grad = torch.ones_like(p)
            opt.step()
    assert [n for n, _ in rec.calls] == ["rsb_clip_adamw_ema_step"] * 3
    for i, (_, a) in enumerate(rec.calls):
        # (table, n_tensors, total_chunks, has_ema, partials, norm_out, max_norm, lr, b1, b2, eps, wd, step, ema_alpha, hyper_device, stream)
        assert a[0] == "p" and a[1] == 2 and a[2] == 2 + 1 and a[3] == 1 and a[4] == "p" and a[5] == "p"
        assert a[6:12] == (1.0, 6e-4, 0.9, 0.999, 1e-5, 0.05) and a[12] == i + 1
        assert a[13] == pytest.approx(min(1 - 1 / (i + 1), 0.99)) and a[14] is None and len(a) == 16
    assert float(opt.state[ps[0]]["step"]) == 3 and opt.global_step == 3
    table = opt._tables[0][1]
    assert table.shape == (2, 7) and table[0, 5] == 5000 and table[1, 5] == 21 and table[0, 6] == 0 and table[1, 6] == 2
    assert table[0, 0] == ps[0].data_ptr() and table[0, 1] == ps[0].grad.data_ptr() and table[0, 4] == ema[0].data_ptr()
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    with pytest.raises(ValueError):
        B200AdamW(ps, ema_params=ema[:1])
    with pytest.raises(ValueError):
        B200AdamW(ps, lr=-1.0)


def test_sliding_window_visits_the_reference_windows(monkeypatch):
    """Window placement of rsuper_b200.inference == the oracle's (pinned on the real inference_sliding_window), including
    the padded small-volume case and gated-out windows (pred pointer NULL)."""
    from oracle.inference_ref import window_starts
    from rsuper_b200.inference import inference_sliding_window

    class Net(torch.nn.Module):
        def forward(self, x):
            return {"segmentation": [x.repeat(1, 3, 1, 1, 1), None]}

    gate = torch.zeros(1, 1, 40, 48, 56)
    gate[:, :, 4:20, 8:30, 10:20] = 1
    for shp, win, g in (((40, 48, 56), (16, 16, 32), None), ((12, 20, 16), (16, 16, 32), None), ((40, 48, 56), (16, 16, 32), gate)):
        with recording(monkeypatch) as rec:
            out = inference_sliding_window(Net(), torch.zeros((1, 1) + shp), SimpleNamespace(window_size=list(win), classes=3),
                                           pancreas=g, keep_on_device=True)
        assert out.shape == (1, 3) + shp
        pshape = tuple(max(s, w) for s, w in zip(shp, win))
        want = [(d0, h0, w0) for d0, _ in window_starts(pshape[0], win[0]) for h0, _ in window_starts(pshape[1], win[1])
                for w0, _ in window_starts(pshape[2], win[2])]
        acc = [a for n, a in rec.calls if n == "rsb_sigmoid_window_accumulate"]
        assert [a[11:14] for a in acc] == want and all(a[3:8] == (1, 3) + pshape and a[8:11] == win for a in acc)
        skipped = [a[0] is None for a in acc]
        if g is None:
            assert not any(skipped)
        else:
            expect = [bool(gate[:, :, d0:d0 + win[0], h0:h0 + win[1], w0:w0 + win[2]].sum() == 0) for d0, h0, w0 in want]
            assert skipped == expect and any(skipped) and not all(skipped)
        assert [n for n, _ in rec.calls][-1] == "rsb_blend_finalize"


def test_batch_assembly_shapes_and_packed_upload(monkeypatch):
    from rsuper_b200 import batch as B
    classes = ["organ"] + [f"c{i}_lesion" for i in range(10)]                    # 11 classes -> 2 byte planes
    ref = synth.make_batch(["mask", "report"], classes, (8, 12, 16), seed=3)
    with recording(monkeypatch) as rec:
        got = B.assemble_batch([ref["image"][b, 0].numpy() for b in range(2)], [synth.pack_masks(ref["label"][b]) for b in range(2)],
                               len(classes), "cpu", unk_packed=[None, synth.pack_masks(ref["unk_channels"][1])],
                               chosen_packed=None, volumes=[None, ref["volumes"][1].numpy()], diameters=None)
    assert [n for n, _ in rec.calls] == ["rsb_unpack_masks"] * 2
    assert all(a[2:6] == (2, 11, 8 * 12 * 16, 0) for _, a in rec.calls)
    assert got["image"].shape == (2, 1, 8, 12, 16) and got["label"].shape == (2, 11, 8, 12, 16) and got["label"].dtype == torch.uint8
    assert got["mask"].shape == got["label"].shape and torch.equal(got["volumes"], ref["volumes"]) and got["diameters"].shape == (2, 10, 3)
    with pytest.raises(ValueError):
        with recording(monkeypatch):
            B.assemble_batch([ref["image"][0, 0].numpy()], [], len(classes), "cpu")
    monkeypatch.undo()                                 # the real device check again: CPU devices are refused
    with pytest.raises(RuntimeError, match="no CPU path"):
        B.assemble_batch([ref["image"][0, 0].numpy()], [synth.pack_masks(ref["label"][0])], len(classes), "cpu")


def test_intensity_augmentation_launches_and_rng_stream(monkeypatch):
    """rsuper_b200.augment consumes np.random / torch's generator exactly like the loader block (same gates, same draws as
    the oracle's draws_like_reference, which is pinned on the real reference) and issues the expected kernels per op."""
    import numpy as np
    from oracle import augment_ref as AR
    from rsuper_b200 import augment as A
    x = torch.zeros(1, 1, 8, 10, 12)
    with recording(monkeypatch) as rec:
        for seed in range(12):
            np.random.seed(seed); torch.manual_seed(100 + seed)
            want = AR.draws_like_reference(x.shape)
            end_np, end_t = np.random.random(), torch.rand(1).item()
            del rec.calls[:]
            np.random.seed(seed); torch.manual_seed(100 + seed)
            A.online_intensity_augmentation(x)
            assert (np.random.random(), torch.rand(1).item()) == (end_np, end_t), seed     # both generators advanced identically
            names = [n for n, _ in rec.calls]
            expect = []
            if "multiply" in want: expect += ["rsb_aug_affine"]
            if "additive" in want: expect += ["rsb_aug_affine"]
            if "gamma" in want: expect += ["rsb_aug_stats", "rsb_aug_gamma", "rsb_aug_stats", "rsb_aug_renorm"]
            if "contrast" in want: expect += ["rsb_aug_stats", "rsb_aug_contrast"]
            if "blur" in want: expect += ["rsb_aug_blur_axis"] * 3
            if "noise" in want: expect += ["rsb_aug_affine"]
            assert names == expect, (seed, names, expect)
            scal = [a for n, a in rec.calls if n == "rsb_aug_affine"]
            if "multiply" in want:
                assert scal[0][3] == pytest.approx(want["multiply"], rel=1e-7) and scal[0][4] == 1
            if "noise" in want:
                assert scal[-1][7] == "p" and scal[-1][8] == pytest.approx(want["noise"][0])
            if "blur" in want:
                blur = [a for n, a in rec.calls if n == "rsb_aug_blur_axis"]
                assert [a[6] for a in blur] == [0, 1, 2] and blur[0][8] == 2 * int(np.ceil(3 * want["blur"])) + 1
    taps = A.gaussian_kernel_1d(1.0)
    assert len(taps) == 7 and abs(sum(taps) - 1) < 1e-12 and taps[3] == max(taps) and taps[0] == pytest.approx(taps[6])
    with pytest.raises(NotImplementedError):
        with recording(monkeypatch):
            A.gamma(torch.zeros(1, 2, 4, 4, 4))


def test_public_dice_loss_wrapper_shapes_and_backward_scale(monkeypatch):
    """DiceLossMultiClass copy-out wrapper: 3-D / 4-D / 5-D inputs reach the kernel as [B, C, V]; class weights are reduced
    to [B, C]; backward passes a (0, grad) scale pair (Dice term only)."""
    from rsuper_b200 import losses
    with recording(monkeypatch) as rec:
        lg = torch.zeros(2, 3, 4, 6, 8, requires_grad=True)
        m = torch.ones(2, 3, 4, 6, 8)
        cw = torch.ones(2, 3, 4, 6, 8) * 2
        losses.DiceLossMultiClass(lg, m, m, class_weights=cw).backward()
        losses.DiceLossMultiClass(lg[0, 0], m[0, 0], m[0, 0])
        losses.DiceLossMultiClass(lg[0], m[0], m[0])
    f = [a[0] for n, a in rec.calls if n == "rsb_seg_loss_forward"]
    assert [(a["B"], a["C"], a["V"]) for a in f] == [(2, 3, 192), (1, 1, 192), (1, 3, 192)]
    assert f[0]["class_weights"] == "p" and f[1]["class_weights"] is None and f[0]["known"] == "p"
    assert [n for n, _ in rec.calls].count("rsb_seg_loss_backward") == 1 and lg.grad is not None
    with pytest.raises(AssertionError):
        losses.DiceLossMultiClass(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 2, 4, 4, 4))
