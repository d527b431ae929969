"""Generate golden vectors by RUNNING THE REAL REFERENCE (read-only /root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
It writes small .npz fixtures next to this file; the tests never import the reference.  Inputs are
regenerated from formulas (oracle/synth.py, oracle/unet_ref.synthetic_*), so only OUTPUTS are stored.

Import recipe: stub nibabel / matplotlib / SimpleITK, register bare `model` and
`model.dim3` packages to bypass model/dim3/__init__.py (needs monai/timm/mmcv), run from a scratch
cwd because the loss writes debug folders on its first calls.
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/rsuper_train"
sys.path.insert(0, ROOT)

from oracle import losses_ref as LR  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.unet_ref import synthetic_image, synthetic_state_dict  # noqa: E402


# (tag, class names, patch shape, sample kinds): scaled-down versions of BASELINE.json configs[3] (PanTS-shaped, 3 classes,
# non-cubic) and configs[4] (7 tumour channels + 1 organ channel), mixed mask / report batches
CONFIG_LOSS_CASES = [
    ("cfg4", ["pancreas", "pancreatic_lesion", "veins"], (24, 48, 32), ("mask", "report")),
    ("cfg5", ["organ"] + sorted(f"{o}_lesion" for o in ("adrenal", "bladder", "colon", "esophagus", "kidney", "liver", "spleen")),
     (32, 32, 32), ("report", "mask")),
]


def import_reference():
    for name in ("nibabel", "matplotlib", "matplotlib.pyplot", "SimpleITK"):
        m = types.ModuleType(name)
        if name == "nibabel":
            m.Nifti1Image = lambda *a, **k: None
            m.save = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    sys.path.insert(0, REF)
    for pkg, sub in (("model", "model"), ("model.dim3", "model/dim3")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[pkg] = m
    unet = importlib.import_module("model.dim3.unet")
    lf = importlib.import_module("training.losses_foundation")
    return unet, lf


# (tag, volume shape, window): ragged last windows, volume smaller than the window (padding path), exact multiples
SLIDING_CASES = [("ragged", (40, 48, 56), (16, 16, 32)), ("small", (12, 20, 16), (16, 16, 32)), ("exact", (32, 32, 32), (16, 16, 16))]


def sliding_test_net():
    """A tiny deterministic 'network' for the window logic: 3 output channels from fixed 3x3x3 filters (so window borders
    matter, like with a real conv net)."""
    w = torch.frac(torch.sin(torch.arange(3 * 27, dtype=torch.float64) * 12.9898) * 43758.5453).float().reshape(3, 1, 3, 3, 3) - 0.5
    net = torch.nn.Conv3d(1, 3, 3, padding=1, bias=True)
    with torch.no_grad():
        net.weight.copy_(w)
        net.bias.copy_(torch.tensor([0.1, -0.2, 0.05]))
    return net.eval()


def reference_medformer(cfg, num_classes):
    """The real MedFormer (model/dim3/medformer.py) in the yaml configuration (config/abdomenatlas_ufo/medformer_3d.yaml)
    scaled by cfg; import_reference() must have run."""
    mf = importlib.import_module("model.dim3.medformer")
    return mf.MedFormer(1, num_classes, base_chan=cfg["base_chan"], map_size=cfg["map_size"], conv_block="BasicBlock",
                        conv_num=cfg["conv_num"], trans_num=cfg["trans_num"], chan_num=cfg["chan_num"], num_heads=cfg["num_heads"],
                        fusion_depth=cfg["fusion_depth"], fusion_dim=cfg["fusion_dim"], fusion_heads=cfg["fusion_heads"],
                        expansion=cfg["expansion"], proj_type="depthwise", norm="in", act="relu", kernel_size=[[3, 3, 3]] * 5,
                        scale=[[2, 2, 2]] * 4, aux_loss=cfg["aux_loss"])


def pack(t: torch.Tensor) -> np.ndarray:
    return np.packbits(t.detach().cpu().numpy().astype(bool).reshape(-1))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    unet_mod, lf = import_reference()
    os.chdir(tempfile.mkdtemp(prefix="rsb_golden_"))
    out = {}

    # ---------------- UNet forward / backward (cfg1: base 8, 32^3, C=2) ----------------
    base, C, S = 8, 2, 32
    net = unet_mod.UNet(1, base, num_classes=C, scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5,
                        block="BasicBlock")
    sd = synthetic_state_dict(base, C)
    assert list(sd.keys()) == [k for k, _ in net.named_parameters()], "state-dict contract changed"
    net.load_state_dict(sd, strict=True)
    x = synthetic_image(1, S, S, S, seed=3)
    logits = net(x)
    out["unet_logits"] = logits.detach().numpy()
    classes = ["organ", "pancreatic_lesion"]
    batch = synth.make_batch(["mask"], classes, (S, S, S), seed=5)
    args = LR.default_args(report_volume_loss_basic=0.0)
    loss = lf.calculate_loss(model_output={"segmentation": logits}, label=batch["label"].long(), unk_voxels=None,
                             args=args, matcher=None, chosen_segment_mask=None, tumor_volumes_report=None,
                             tumor_diameters=None, classes=classes)
    loss["overall"].backward()
    out["unet_loss_overall"] = np.float32(loss["overall"].item())
    out["unet_loss_segmentation"] = np.float32(loss["segmentation"].item())
    out["unet_grad_norms"] = np.array([p.grad.norm().item() for _, p in net.named_parameters()], dtype=np.float64)
    for k in ("outc.weight", "outc.bias", "inc.conv1.weight", "down4.conv.2.conv2.conv.weight"):
        g = dict(net.named_parameters())[k].grad
        out["unet_grad::" + k] = g.detach().numpy() if g.numel() < 4096 else g.detach().reshape(-1)[:4096].numpy()

    # ---------------- UNet(block='SingleConv') forward / backward ----------------
    net1 = unet_mod.UNet(1, base, num_classes=C, scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5, block="SingleConv")
    sd1 = synthetic_state_dict(base, C, block="SingleConv")
    assert list(sd1.keys()) == [k for k, _ in net1.named_parameters()], "SingleConv state-dict contract changed"
    net1.load_state_dict(sd1, strict=True)
    logits1 = net1(x)
    out["unet_single_logits"] = logits1.detach().numpy()
    loss1 = lf.calculate_loss(model_output={"segmentation": logits1}, label=batch["label"].long(), unk_voxels=None,
                              args=args, matcher=None, chosen_segment_mask=None, tumor_volumes_report=None,
                              tumor_diameters=None, classes=classes)
    loss1["overall"].backward()
    out["unet_single_loss_overall"] = np.float32(loss1["overall"].item())
    out["unet_single_grad_norms"] = np.array([p.grad.norm().item() for _, p in net1.named_parameters()], dtype=np.float64)

    # ---------------- MedFormer forward / deep-supervision loss / backward (oracle groundwork) ----------------
    from oracle.medformer_ref import SMALL_CFG, fill_like
    mnet = reference_medformer(SMALL_CFG, C)
    msd = fill_like([(k, tuple(v.shape)) for k, v in mnet.named_parameters()])
    mnet.load_state_dict(msd, strict=True)
    out["medformer_param_names"] = np.array([k for k, _ in mnet.named_parameters()])
    out["medformer_param_shapes"] = np.array([",".join(str(d) for d in v.shape) for _, v in mnet.named_parameters()])
    mo = mnet(x)
    out["medformer_logits"] = mo["segmentation"][0].detach().numpy()[:, :, ::2, ::2, ::2].copy()
    out["medformer_aux"] = mo["segmentation"][1].detach().numpy()[:, :, ::2, ::2, ::2].copy()
    out["medformer_logits_abs_sum"] = np.float64(mo["segmentation"][0].double().abs().sum().item())
    mloss = lf.calculate_loss(model_output=mo, label=batch["label"].long(), unk_voxels=None, args=args, matcher=None,
                              chosen_segment_mask=None, tumor_volumes_report=None, tumor_diameters=None, classes=classes)
    mloss["overall"].backward()
    out["medformer_loss_overall"] = np.float32(mloss["overall"].item())
    out["medformer_grad_norms"] = np.array([p.grad.norm().item() for _, p in mnet.named_parameters()], dtype=np.float64)

    # ---------------- sliding-window inference (SURVEY §8f N3 groundwork): the real inference3d.inference_sliding_window ----------------
    inf3d = importlib.import_module("inference.inference3d")
    tiny = sliding_test_net()
    for tag, shp, win in SLIDING_CASES:
        vol = synthetic_image(1, *shp, seed=9)
        a = types.SimpleNamespace(window_size=list(win), classes=3)
        full = inf3d.inference_sliding_window(tiny, vol, a)
        out[f"sliding_{tag}"] = full.numpy()[:, :, 1::3, 1::3, 1::3].copy()       # every 3rd voxel + a checksum of all of them
        out[f"sliding_{tag}_sum"] = np.float64(full.double().sum().item())
    gate = torch.zeros(1, 1, 40, 48, 56)
    gate[:, :, 4:20, 8:30, 10:20] = 1
    a = types.SimpleNamespace(window_size=[16, 16, 32], classes=3)
    full = inf3d.inference_sliding_window(tiny, synthetic_image(1, 40, 48, 56, seed=9), a, pancreas=gate)
    out["sliding_gated"] = full.numpy()[:, :, 1::3, 1::3, 1::3].copy()
    out["sliding_gated_sum"] = np.float64(full.double().sum().item())

    # ---------------- structuring elements & dilation ----------------
    for d in (1, 3, 5, 7, 11):
        out[f"ball_{d}"] = pack(lf.create_ball_kernel(d))
        out[f"ball_{d}_size"] = np.int32(lf.create_ball_kernel(d).shape[0])
    g = lf.create_ball_kernel(9, gaussian=True, gaussian_std=1.5)
    out["ball_gauss_9"] = g.numpy()
    vol = synth.make_batch(["report", "mask"], classes, (24, 28, 20), seed=9)["unk_channels"].float()
    for k in (1, 3, 5, 7, 9, 15, 31):
        out[f"dilate_{k}"] = pack(lf.dilate_volume(vol, k))
    out["known_voxels"] = pack(lf.get_known_voxels(vol, vol, sanity=False))

    # ---------------- segmentation loss (with gradient) ----------------
    shp = (16, 24, 32)
    lg = synth.synthetic_logits(2, 3, shp, seed=2).requires_grad_(True)
    cls3 = ["liver", "liver_lesion", "pancreas"]
    b3 = synth.make_batch(["mask", "report"], cls3, shp, seed=11)
    known = lf.get_known_voxels(b3["label"], b3["unk_channels"].float(), sanity=False)
    l_bce = (torch.nn.functional.binary_cross_entropy_with_logits(lg, b3["label"].float(), reduction="none") * known).mean()
    l_dice = lf.DiceLossMultiClass(lg, b3["label"].float(), known, sigmoid=True)
    (l_bce + l_dice).backward()
    out["seg_bce"], out["seg_dice"] = np.float32(l_bce.item()), np.float32(l_dice.item())
    out["seg_grad"] = lg.grad.detach().numpy()[:, :, ::4, ::4, ::4].copy()
    out["seg_grad_sum"] = np.float64(lg.grad.double().abs().sum().item())
    cw = torch.tensor([[1.0, 2.0, 0.5], [0.25, 1.0, 3.0]])
    lg2 = lg.detach().clone().requires_grad_(True)
    cw5 = cw[:, :, None, None, None]
    l2 = (torch.nn.functional.binary_cross_entropy_with_logits(lg2, b3["label"].float(), reduction="none", weight=cw5)
          * known).mean() + lf.DiceLossMultiClass(lg2, b3["label"].float(), known, sigmoid=True, class_weights=cw5)
    l2.backward()
    out["seg_cw_loss"] = np.float32(l2.item())
    out["seg_cw_grad_sum"] = np.float64(lg2.grad.double().abs().sum().item())

    # ---------------- volume loss ----------------
    lg3 = lg.detach().clone().requires_grad_(True)
    vl = lf.volume_loss_basic(lg3, b3["mask"].float(), b3["volumes"], b3["label"].float(), b3["unk_channels"].float(),
                              cls3, tolerance=0.2)["dice_volume_loss"]
    vl.backward()
    out["volume_loss"] = np.float32(vl.item())
    out["volume_grad_sum"] = np.float64(lg3.grad.double().abs().sum().item())
    xs = torch.tensor([[0.0, 50.0, 900.0, 1000.0, 5000.0]])
    ys = torch.tensor([[0.0, 80.0, 1000.0, 1000.0, 1000.0]])
    out["dice_volume_curve"] = lf.dice_based_volume_loss(xs, ys, tolerance=0.2, E=500).numpy()

    # ---------------- GWRP, isolate_tumor, ball loss ----------------
    xv = torch.sigmoid(synth.synthetic_logits(1, 1, (12, 12, 12), seed=4)[0, 0])
    pm = (xv > 0.6).float()
    out["gwrp_weights"] = lf.GlobalWeightedRankPooling(xv * pm + pm, N=pm.sum(), c=0.5, return_weights=True,
                                                       hard_cutoff=True).numpy()
    shp_b = (32, 32, 32)
    cls2 = ["organ", "pancreatic_lesion"]
    bb = synth.make_batch(["report", "mask"], cls2, shp_b, seed=21)
    lgb = synth.synthetic_logits(2, 2, shp_b, seed=6, scale=2.0)
    prob = torch.sigmoid(lgb[0, 1]) * lf.dilate_volume(bb["mask"][0, 1].float(), 31)
    dia, volm = bb["diameters"][0, 0].max().item(), bb["volumes"][0, 0].item()
    m, ms, mb = lf.isolate_tumor(prob, dia, True, 1.5, volm, diameter_margin=0.2, volume_margin=0.2)
    out["isolate_mask"], out["isolate_small"], out["isolate_big"] = pack(m), pack(ms), pack(mb)
    out["isolate_args"] = np.array([dia, volm], dtype=np.float64)
    lgb_g = lgb.clone().requires_grad_(True)
    bl = lf.ball_loss(out=lgb_g, labels=bb["label"].float(), unk_voxels=bb["unk_channels"].float(),
                      chosen_segment_mask=bb["mask"].float(), tumor_volumes=bb["volumes"], tumor_diameters=bb["diameters"],
                      classes=cls2, apply_dice_loss=True, diameter_margin=0.2, volume_margin=0.2)
    (bl["ball_loss_bce"] + bl["ball_loss_dice"]).backward()
    out["ball_loss_bce"], out["ball_loss_dice"] = np.float32(bl["ball_loss_bce"].item()), np.float32(bl["ball_loss_dice"].item())
    out["ball_grad_sum"] = np.float64(lgb_g.grad.double().abs().sum().item())

    # ---------------- calculate_loss dictionaries ----------------
    for tag, lossname, deep in (("ball_dice_last_deep", "ball_dice_last", True), ("dice", "dice", False),
                                ("ball", "ball", False), ("both", "ball_dice_both", False)):
        a = LR.default_args(loss=lossname)
        lgc = lgb.clone().requires_grad_(True)
        mo = {"segmentation": [lgc, lgc * 0.5 + 0.1]} if deep else {"segmentation": lgc}
        res = lf.calculate_loss(model_output=mo, label=bb["label"].long(), unk_voxels=bb["unk_channels"].float(), args=a,
                                matcher=None, chosen_segment_mask=bb["mask"].float(), tumor_volumes_report=bb["volumes"],
                                tumor_diameters=bb["diameters"], classes=cls2, input_tensor=bb["image"])
        res["overall"].backward()
        out[f"calc::{tag}::keys"] = np.array(sorted(res.keys()))
        for k, v in res.items():
            out[f"calc::{tag}::{k}"] = np.float32(v.item())
        out[f"calc::{tag}::grad_sum"] = np.float64(lgc.grad.double().abs().sum().item())

    # ---------------- calculate_loss on the class lists of BASELINE.json configs[3..4] (PanTS 3-class, 7-tumour 8-class) ----------------
    for tag, classes_c, shape_c, kinds_c in CONFIG_LOSS_CASES:
        bc = synth.make_batch(list(kinds_c), classes_c, shape_c, seed=17)
        lgc = synth.synthetic_logits(len(kinds_c), len(classes_c), shape_c, seed=13, scale=2.0).requires_grad_(True)
        res = lf.calculate_loss(model_output={"segmentation": lgc}, label=bc["label"].long(), unk_voxels=bc["unk_channels"].float(),
                                args=LR.default_args(), matcher=None, chosen_segment_mask=bc["mask"].float(),
                                tumor_volumes_report=bc["volumes"], tumor_diameters=bc["diameters"], classes=classes_c,
                                input_tensor=bc["image"])
        res["overall"].backward()
        out[f"cfgloss::{tag}::keys"] = np.array(sorted(res.keys()))
        for k, v in res.items():
            out[f"cfgloss::{tag}::{k}"] = np.float32(v.item())
        out[f"cfgloss::{tag}::grad_sum"] = np.float64(lgc.grad.double().abs().sum().item())

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "reference_outputs.npz"))
    print(f"wrote reference_outputs.npz ({sz / 1024:.1f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
