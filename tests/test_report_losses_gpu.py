"""GPU parity of the report-supervised losses (Volume loss, Ball loss, calculate_loss dicts) through the
reference-facing API of rsuper_b200 against the oracle (CPU, fp32) on the same seeded inputs, and against the
golden values recorded from the REAL reference (tests/golden/reference_outputs.npz).

Tolerances: discrete structures (ball centre, pseudo / dilated / penalize masks) bit-exact; losses 1e-5 (north star);
gradients 1e-4 relative to their maximum."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def pack(t):
    return np.packbits(t.detach().cpu().numpy().astype(bool).reshape(-1))


def _seg_inputs():
    from oracle import synth
    shp = (16, 24, 32)
    cls3 = ["liver", "liver_lesion", "pancreas"]
    lg = synth.synthetic_logits(2, 3, shp, seed=2)
    b3 = synth.make_batch(["mask", "report"], cls3, shp, seed=11)
    return lg, b3, cls3


def _ball_inputs(shape=(32, 32, 32), seed=21, kinds=("report", "mask"), lseed=6):
    from oracle import synth
    cls2 = ["organ", "pancreatic_lesion"]
    bb = synth.make_batch(list(kinds), cls2, shape, seed=seed)
    lgb = synth.synthetic_logits(len(kinds), 2, shape, seed=lseed, scale=2.0)
    return bb, lgb, cls2


def test_volume_loss(cuda_dev, golden):
    from oracle import losses_ref as LR
    from rsuper_b200 import report_losses as RL
    lg, b3, cls3 = _seg_inputs()
    ref_in = lg.clone().requires_grad_(True)
    ref = LR.volume_loss_basic(ref_in, b3["mask"].float(), b3["volumes"], b3["label"].float(), b3["unk_channels"].float(), cls3,
                               tolerance=0.2)["dice_volume_loss"]
    ref.backward()
    x = lg.to(cuda_dev).requires_grad_(True)
    out = RL.volume_loss_basic(x, b3["mask"].to(cuda_dev), b3["volumes"].to(cuda_dev), b3["label"].to(cuda_dev),
                               b3["unk_channels"].to(cuda_dev), cls3, tolerance=0.2)["dice_volume_loss"]
    out.backward()
    assert abs(out.item() - ref.item()) <= 1e-6 and abs(out.item() - float(golden["volume_loss"])) <= 1e-6
    assert rel(x.grad.cpu(), ref_in.grad) <= 1e-4
    assert abs(x.grad.double().abs().sum().item() / float(golden["volume_grad_sum"]) - 1) <= 1e-4
    # class weights and the larger tolerance branch
    cw = torch.tensor([[1.0, 2.0, 0.5], [0.25, 1.5, 3.0]])
    r2 = LR.volume_loss_basic(lg, b3["mask"].float(), b3["volumes"], b3["label"].float(), b3["unk_channels"].float(), cls3,
                              tolerance=0.1, class_weights=cw[:, :, None, None, None])["dice_volume_loss"]
    o2 = RL.volume_loss_basic(lg.to(cuda_dev), b3["mask"].to(cuda_dev), b3["volumes"].to(cuda_dev), b3["label"].to(cuda_dev),
                              b3["unk_channels"].to(cuda_dev), cls3, tolerance=0.1, class_weights=cw.to(cuda_dev))["dice_volume_loss"]
    assert abs(o2.item() - r2.item()) <= 1e-6


def test_isolate_tumor_bit_exact(cuda_dev, golden):
    from oracle import losses_ref as LR
    from rsuper_b200 import report_losses as RL
    bb, lgb, _ = _ball_inputs()
    prob = torch.sigmoid(lgb[0, 1]) * LR.dilate_volume(bb["mask"][0, 1].float(), 31)
    dia, volm = float(golden["isolate_args"][0]), float(golden["isolate_args"][1])
    m, ms, mb = RL.isolate_tumor(prob.to(cuda_dev).contiguous(), dia, True, 1.5, volm, diameter_margin=0.2, volume_margin=0.2)
    assert np.array_equal(pack(m), golden["isolate_mask"])
    assert np.array_equal(pack(ms), golden["isolate_small"])
    assert np.array_equal(pack(mb), golden["isolate_big"])
    # second configuration: ball clipped by the volume border, several margins
    for seed, dia2 in ((5, 9.0), (8, 14.0)):
        bb2, lg2, _ = _ball_inputs(shape=(32, 48, 40), seed=seed, lseed=seed)
        p2 = torch.sigmoid(lg2[0, 1]) * LR.dilate_volume(bb2["mask"][0, 1].float(), 31)
        vol2 = 4.0 / 3.0 * np.pi * (dia2 / 2) ** 3
        ref = LR.isolate_tumor(p2, dia2, True, 1.5, vol2, diameter_margin=0.5, volume_margin=0.5)
        got = RL.isolate_tumor(p2.to(cuda_dev).contiguous(), dia2, True, 1.5, vol2, diameter_margin=0.5, volume_margin=0.5)
        for a, b in zip(got, ref):
            assert np.array_equal(pack(a), pack(b)), (seed, dia2)


def test_gwrp_weights(cuda_dev, golden):
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import ops
    xl = synth.synthetic_logits(1, 1, (12, 12, 12), seed=4)[0, 0]
    xv = torch.sigmoid(xl)
    pm = (xv > 0.6)
    n = int(pm.sum())
    ref = LR.gwrp_weights(xv * pm.float() + pm.float(), N=pm.float().sum(), c=0.5, hard_cutoff=True)
    np.testing.assert_allclose(ref.numpy(), golden["gwrp_weights"], rtol=1e-5, atol=1e-9)
    x = xl.to(cuda_dev).contiguous()
    pseudo = pm.to(torch.uint8).to(cuda_dev).reshape(-1).contiguous()
    wmap = torch.zeros(x.numel(), device=cuda_dev)
    cand = torch.empty((n, 2), dtype=torch.int32, device=cuda_dev)
    n_cand = torch.empty(1, dtype=torch.int32, device=cuda_dev)
    ops.ball_candidates(x, pseudo, 1, (0, 0, 0), 0, 0.0, cand, n_cand, None)
    ops.ball_rank_gwrp(cand, n_cand, n, 0.5, wmap)
    assert int(n_cand.item()) == n
    np.testing.assert_allclose(wmap.cpu().numpy().reshape(12, 12, 12) / n, golden["gwrp_weights"], rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("cfg", [dict(), dict(shape=(32, 48, 32), seed=33, kinds=("mask", "report", "report"), lseed=9)])
def test_ball_loss(cuda_dev, golden, cfg):
    from oracle import losses_ref as LR
    from rsuper_b200 import report_losses as RL
    bb, lgb, cls2 = _ball_inputs(**cfg)
    ref_in = lgb.clone().requires_grad_(True)
    dbg_ref, dbg = {}, {}
    ref = LR.ball_loss(ref_in, bb["label"].float(), bb["unk_channels"].float(), bb["mask"].float(), bb["volumes"], bb["diameters"],
                       cls2, apply_dice_loss=True, diameter_margin=0.2, volume_margin=0.2, debug=dbg_ref)
    (ref["ball_loss_bce"] + 0.5 * ref["ball_loss_dice"]).backward()
    x = lgb.to(cuda_dev).requires_grad_(True)
    out = RL.ball_loss(x, bb["label"].to(cuda_dev), bb["unk_channels"].to(cuda_dev), bb["mask"].to(cuda_dev), bb["volumes"].to(cuda_dev),
                       bb["diameters"].to(cuda_dev), cls2, apply_dice_loss=True, diameter_margin=0.2, volume_margin=0.2, debug=dbg)
    (out["ball_loss_bce"] + 0.5 * out["ball_loss_dice"]).backward()
    for k in ("pseudo", "dilated", "penalize"):
        assert len(dbg[k]) == len(dbg_ref[k])
        for a, b in zip(dbg[k], dbg_ref[k]):
            assert np.array_equal(pack(a), pack(b)), k
    assert abs(out["ball_loss_bce"].item() - ref["ball_loss_bce"].item()) <= 1e-5
    assert abs(out["ball_loss_dice"].item() - ref["ball_loss_dice"].item()) <= 1e-5
    assert rel(x.grad.cpu(), ref_in.grad) <= 1e-4
    if not cfg:
        assert abs(out["ball_loss_bce"].item() - float(golden["ball_loss_bce"])) <= 1e-5
        assert abs(out["ball_loss_dice"].item() - float(golden["ball_loss_dice"])) <= 1e-5


def test_calculate_loss_dicts(cuda_dev, golden):
    from oracle import losses_ref as LR
    from rsuper_b200 import losses
    bb, lgb, cls2 = _ball_inputs()
    dev = cuda_dev
    for tag, lossname, deep in (("ball_dice_last_deep", "ball_dice_last", True), ("dice", "dice", False),
                                ("ball", "ball", False), ("both", "ball_dice_both", False)):
        a = LR.default_args(loss=lossname)
        lg = lgb.to(dev).clone().requires_grad_(True)     # clone: on a CPU device (emulated run) .to() would alias lgb
        mo = {"segmentation": [lg, lg * 0.5 + 0.1]} if deep else {"segmentation": lg}
        res = losses.calculate_loss(mo, bb["label"].long().to(dev), bb["unk_channels"].float().to(dev), a, None, bb["mask"].float().to(dev),
                                    bb["volumes"].to(dev), bb["diameters"].to(dev), cls2, input_tensor=bb["image"].to(dev))
        assert sorted(res.keys()) == list(golden[f"calc::{tag}::keys"]), tag
        for k, v in res.items():
            assert abs(v.item() - float(golden[f"calc::{tag}::{k}"])) <= 1e-5, (tag, k, v.item(), float(golden[f"calc::{tag}::{k}"]))
        res["overall"].backward()
        assert abs(lg.grad.double().abs().sum().item() / float(golden[f"calc::{tag}::grad_sum"]) - 1) <= 1e-4, tag
    # inconsistent report batch raises like the reference (:864-869)
    with pytest.raises(ValueError):
        losses.calculate_loss({"segmentation": lgb.to(dev).requires_grad_(True)}, bb["label"].long().to(dev),
                              torch.zeros_like(bb["unk_channels"]).float().to(dev), LR.default_args(loss="ball"), None,
                              bb["mask"].float().to(dev), bb["volumes"].to(dev), bb["diameters"].to(dev), cls2)



@pytest.mark.parametrize("shape,diameter", [((32, 32, 32), 9), ((24, 40, 56), 15), ((64, 64, 64), 31), ((17, 23, 37), 5), ((8, 8, 8), 3)])
def test_separable_ball_correlation_matches_tap_list(cuda_dev, shape, diameter):
    """The rows -> discs -> planes decomposition of the truncated Gaussian ball (rsb_ball_correlate_argmax_sep) against the
    direct tap-list kernel (the definition): same first maximum, scores equal up to the dropped normalisation 1 / Z and fp32
    re-association; ragged shapes, ball larger than the volume, empty rows, all-zero volume."""
    from rsuper_b200 import ops
    from rsuper_b200 import report_losses as RL
    g = torch.Generator().manual_seed(diameter)
    x = torch.rand(shape, generator=g)
    blob = torch.zeros(shape)
    blob[shape[0] // 4:shape[0] // 4 * 3, shape[1] // 3:shape[1] // 3 * 2, 2:shape[2] - 3] = 1      # empty rows around a block
    for vol in (x * blob, x, torch.zeros(shape)):
        xi = vol.to(cuda_dev).contiguous()
        taps, support, khalf = RL._gauss_ball_taps(diameter, True, 1.5, cuda_dev)
        g1d, wtab, reach, g_host, w_host = RL._gauss_ball_sep(diameter, 1.5, cuda_dev)
        k_tap = int(ops.ball_correlate_argmax(xi, taps, khalf).item())
        k_sep = int(ops.ball_correlate_argmax_sep(xi, g1d, wtab, reach, g_host, w_host).item())
        assert k_sep == int(ops.ball_correlate_argmax_sep(xi, g1d, wtab, reach).item())      # tiled and per-row disc stages agree bit for bit
        assert (k_tap & 0xFFFFFFFF) == (k_sep & 0xFFFFFFFF), (shape, diameter)
        s_tap = np.array([k_tap >> 32], dtype=np.uint32).view(np.float32)[0]
        s_sep = np.array([k_sep >> 32], dtype=np.uint32).view(np.float32)[0]
        # tap weights are normalised to sum 1; the separable factors are not: Z = sum of g(dz) g(dy) g(dx) over the ball
        w = taps.cpu().numpy()[:, 3].copy().view(np.float32)
        gz = g1d.cpu().numpy()
        off = np.abs(taps.cpu().numpy()[:, :3])
        Z = float((gz[off[:, 0]] * gz[off[:, 1]] * gz[off[:, 2]]).sum(dtype=np.float64))
        assert abs(s_sep / Z - s_tap) <= 2e-5 * max(s_tap, 1e-12) + 1e-12
        np.testing.assert_allclose(gz[off[:, 0]] * gz[off[:, 1]] * gz[off[:, 2]] / Z, w, rtol=2e-5)


@pytest.mark.parametrize("shape", [(16, 24, 32), (9, 13, 70), (8, 8, 31), (40, 33, 64)])
@pytest.mark.parametrize("k", [1, 3, 5, 7, 11, 31])
def test_dilation_tile_kernel_edge_shapes(cuda_dev, shape, k):
    """The bit-sliced tile kernel against the oracle's conv-based dilation on ragged volumes (W not a multiple of 8 / 32,
    tiles cut by every border), several volumes per launch incl. an empty one."""
    from oracle import losses_ref as LR
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(k)
    vols = (torch.rand((3,) + shape, generator=g) < 0.01).to(torch.uint8)
    vols[1] = 0
    vols[2, 0, 0, 0] = 1
    vols[2, -1, -1, -1] = 1
    got = ops.dilate_ball(vols.to(cuda_dev), k).cpu()
    ref = torch.stack([LR.dilate_volume(v.float(), k) for v in vols]).to(torch.uint8)
    assert torch.equal(got, ref)


MERGE_CLASSES = ["liver", "liver_lesion_1", "liver_lesion_2", "pancreatic_cyst", "kidney_cyst_lesion"]


def _merge_inputs(tag):
    from oracle import synth
    lg = synth.synthetic_logits(2, len(MERGE_CLASSES), (16, 24, 32), seed=4)
    bt = synth.make_batch(["report", "mask"], MERGE_CLASSES, (16, 24, 32), seed={"merged": 20, "single": 23}[tag])
    return lg, bt


def _golden_merge():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_merge.npz"))


@pytest.mark.parametrize("tag", ["merged", "single"])
def test_lesion_group_max_merge(cuda_dev, tag):
    """Lesion groups that merge several channels of one organ (get_lesion_channels, losses_foundation.py:204-248): a
    two-channel group ('liver_lesion_1' + 'liver_lesion_2') and a channel that joins two groups ('kidney_cyst_lesion') —
    Volume and Ball loss values and the per-channel gradient mass against values recorded from the REAL reference
    (tests/golden/make_golden_merge.py) and against the oracle element by element (the gradient of a merged row goes to the
    member that attained the maximum)."""
    from oracle import losses_ref as LR
    from rsuper_b200 import report_losses as RL
    gold = _golden_merge()
    lg, bt = _merge_inputs(tag)
    assert RL.lesion_channels(MERGE_CLASSES) == [[1, 2], [3], [4], [4]] and list(gold["names"]) == ["liver_lesion", "pancreas_cyst",
                                                                                                    "kidney_cyst_lesion", "kidney_cyst"]
    dev = {k: v.to(cuda_dev) for k, v in bt.items()}
    # volume loss
    xr = lg.detach().clone().requires_grad_(True)
    ref = LR.volume_loss_basic(xr, bt["mask"].float(), bt["volumes"], bt["label"].float(), bt["unk_channels"].float(), MERGE_CLASSES,
                               tolerance=0.2)["dice_volume_loss"]
    ref.backward()
    x = lg.detach().clone().to(cuda_dev).requires_grad_(True)
    out = RL.volume_loss_basic(x, dev["mask"], dev["volumes"], dev["label"], dev["unk_channels"], MERGE_CLASSES, tolerance=0.2)["dice_volume_loss"]
    out.backward()
    assert abs(out.item() - float(gold[f"{tag}_volume_loss"])) <= 1e-6 and abs(out.item() - ref.item()) <= 1e-6
    assert rel(x.grad.cpu(), xr.grad) <= 1e-4
    np.testing.assert_allclose(x.grad.double().abs().sum(dim=(0, 2, 3, 4)).cpu().numpy(), gold[f"{tag}_volume_grad_per_channel"], rtol=1e-4, atol=1e-9)
    if tag == "merged":
        assert x.grad[:, 1].abs().sum() > 0 and x.grad[:, 2].abs().sum() > 0          # both members of the group receive gradient
    # ball loss
    x2r = lg.detach().clone().requires_grad_(True)
    rb = LR.ball_loss(x2r, bt["label"].float(), bt["unk_channels"].float(), bt["mask"].float(), bt["volumes"], bt["diameters"], MERGE_CLASSES,
                      apply_dice_loss=True, diameter_margin=0.2, volume_margin=0.2)
    (rb["ball_loss_bce"] + rb["ball_loss_dice"]).backward()
    x2 = lg.detach().clone().to(cuda_dev).requires_grad_(True)
    gb = RL.ball_loss(x2, dev["label"], dev["unk_channels"], dev["mask"], dev["volumes"], dev["diameters"], MERGE_CLASSES,
                      apply_dice_loss=True, diameter_margin=0.2, volume_margin=0.2)
    (gb["ball_loss_bce"] + gb["ball_loss_dice"]).backward()
    for k in ("ball_loss_bce", "ball_loss_dice"):
        assert abs(gb[k].item() - float(gold[f"{tag}_{k}"])) <= 1e-5 * max(1.0, abs(float(gold[f"{tag}_{k}"]))), k
        assert abs(gb[k].item() - rb[k].item()) <= 1e-5 * max(1.0, abs(rb[k].item())), k
    assert rel(x2.grad.cpu(), x2r.grad) <= 1e-4
    np.testing.assert_allclose(x2.grad.double().abs().sum(dim=(0, 2, 3, 4)).cpu().numpy(), gold[f"{tag}_ball_grad_per_channel"], rtol=1e-4, atol=1e-9)
