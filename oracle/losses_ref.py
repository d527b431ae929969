"""ORACLE (test infrastructure, never shipped, never imported by the product path).

Restatement in plain torch (CPU or CUDA, fp32) of the loss side of the R-Super train step,
rsuper_train/training/losses_foundation.py.  Each function cites the lines it follows.  Debug
dumps (NIfTI/yaml writers, prints) and baseline-only branches (classification, CLIP, Model Genesis,
Hungarian matching) are out of scope and are not restated.

Pinned by tests/golden/make_golden.py against the real reference module.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SMOOTH = 1e-5


# --------------------------------------------------------------------------------------------
# ball structuring elements and dilation
# --------------------------------------------------------------------------------------------
def _odd_ceil(v: float) -> int:
    k = math.ceil(v)
    return k + 1 if k % 2 == 0 else k


def create_ball_kernel(diameter, gaussian: bool = False, gaussian_std: float = 1.5) -> torch.Tensor:
    """losses_foundation.py:1161-1232.  Ball of odd(ceil(d)) inside a grid of odd(ceil(1.2*that));
    optionally a truncated Gaussian (sigma = gaussian_std * radius) normalised to sum 1."""
    d_odd = _odd_ceil(diameter)
    size = _odd_ceil(1.2 * d_odd)
    radius = d_odd / 2.0
    c = torch.arange(size, dtype=torch.float32) - (size - 1) / 2.0
    d2 = c[:, None, None] ** 2 + c[None, :, None] ** 2 + c[None, None, :] ** 2
    inside = (d2 <= radius ** 2).float()
    if not gaussian:
        return inside
    sigma = gaussian_std * radius
    k = torch.exp(-d2 / (2.0 * sigma ** 2)) * inside
    return k / k.sum()


def dilate_volume_conv(volume: torch.Tensor, kernel_size: int) -> torch.Tensor:
    """losses_foundation.py:50-99: depthwise conv with the 0/1 ball, then `> 0`."""
    shape = volume.shape
    v = volume.reshape((1,) * (5 - volume.dim()) + tuple(shape)) if volume.dim() < 5 else volume
    if kernel_size % 2 == 0:
        kernel_size += 1
    ball = create_ball_kernel(kernel_size).to(v)
    ch = v.shape[1]
    w = ball[None, None].repeat(ch, 1, 1, 1, 1)
    out = (F.conv3d(v, w, padding=ball.shape[-1] // 2, groups=ch) > 0).float()
    return out.reshape(shape)


def dilate_volume(volume: torch.Tensor, kernel_size: int, full_pass_radius: int = 3) -> torch.Tensor:
    """losses_foundation.py:22-46: one pass when k <= 7, else floor(r/3) passes of k=7 + remainder."""
    if kernel_size % 2 == 0:
        kernel_size += 1
    if kernel_size <= 2 * full_pass_radius + 1:
        return dilate_volume_conv(volume, kernel_size)
    radius = (kernel_size - 1) // 2
    for _ in range(radius // full_pass_radius):
        volume = dilate_volume_conv(volume, 2 * full_pass_radius + 1)
    rem = radius % full_pass_radius
    if rem > 0:
        volume = dilate_volume_conv(volume, 2 * rem + 1)
    return volume


def get_known_voxels(unk_voxels: torch.Tensor, dilation: int = 5) -> torch.Tensor:
    """losses_foundation.py:150-199 (sanity dumps omitted): 1 - dilate(unk, 5)."""
    unk = unk_voxels.float()
    if not torch.equal(unk.bool().float(), unk):
        raise AssertionError("unk_voxels must be binary")
    if dilation > 0:
        unk = dilate_volume(unk, dilation)
    return 1.0 - unk


# --------------------------------------------------------------------------------------------
# lesion channel selection
# --------------------------------------------------------------------------------------------
def lesion_groups(classes: Sequence[str]) -> Tuple[List[str], List[List[int]]]:
    """losses_foundation.py:209-218: channels whose name contains lesion|cyst|pdac|pnet, grouped by
    the organ prefix ('pancreatic' -> 'pancreas').  Mirrors the reference loop, including the fact
    that a name matching several suffixes is appended once per matching suffix."""
    names: List[str] = []
    groups: Dict[str, List[int]] = {}
    for i, cl in enumerate(classes):
        for suffix in ("lesion", "cyst", "pdac", "pnet"):
            if suffix in cl:
                key = cl[:cl.index("_" + suffix) + len("_" + suffix)].replace("pancreatic", "pancreas")
                if key not in groups:
                    groups[key] = []
                    names.append(key)
                groups[key].append(i)
    return names, [groups[k] for k in names]


def get_lesion_channels(t: torch.Tensor, classes: Sequence[str]) -> torch.Tensor:
    """losses_foundation.py:204-248: [B,C,...] -> [B,L,...], max-merged per organ."""
    assert t.shape[1] == len(classes)
    _, groups = lesion_groups(classes)
    outs = [torch.stack([t[:, i] for i in g], dim=0).max(dim=0).values for g in groups]
    return torch.stack(outs, dim=1).type_as(t)


# --------------------------------------------------------------------------------------------
# segmentation loss
# --------------------------------------------------------------------------------------------
def dice_loss_multiclass(preds, targets, known_voxels, sigmoid: bool = True, class_weights=None,
                         reduce: bool = True) -> torch.Tensor:
    """DiceLossMultiClass, losses_foundation.py:541-607 (adaptive Tversky; alpha carries gradient)."""
    while preds.dim() < 5:
        preds, targets, known_voxels = preds.unsqueeze(0), targets.unsqueeze(0), known_voxels.unsqueeze(0)
    assert preds.shape == targets.shape == known_voxels.shape
    n, c = preds.shape[:2]
    p = torch.sigmoid(preds) if sigmoid else preds
    p = p * known_voxels
    t = targets * known_voxels
    tp = (p * t).flatten(2).sum(-1)
    fp = (p * (1 - t)).flatten(2).sum(-1)
    fn = ((1 - p) * t).flatten(2).sum(-1)
    fp_c, fn_c = fp.sum(0), fn.sum(0)  # summed over the (local) batch, per class  (:581)
    alpha = (fp_c / (fp_c + fn_c + SMOOTH)).clamp(0.2, 0.8).unsqueeze(0).expand(n, c)
    beta = 1 - alpha
    dice = tp / (tp + alpha * fp + beta * fn + SMOOTH)
    loss = 1 - dice
    if class_weights is not None:
        cw = class_weights.mean(dim=(-1, -2, -3))
        while cw.dim() < loss.dim():
            cw = cw.unsqueeze(0)
        assert cw.shape == loss.shape
        loss = loss * cw
    return loss.mean() if reduce else loss


def seg_loss(logits, label, known_voxels, class_weights=None) -> torch.Tensor:
    """BCE term + Dice term of calculate_loss, losses_foundation.py:945-956 / 1030-1035."""
    bce = F.binary_cross_entropy_with_logits(logits, label.float(), reduction="none", weight=class_weights)
    bce = (bce * known_voxels).mean()
    return bce + dice_loss_multiclass(logits, label, known_voxels, sigmoid=True, class_weights=class_weights)


# --------------------------------------------------------------------------------------------
# Volume Loss
# --------------------------------------------------------------------------------------------
def dice_based_volume_loss(x, y, tolerance: float = 0.1, E: float = 500.0) -> torch.Tensor:
    """losses_foundation.py:352-395 (cross_entropy=False branch)."""
    if x.dim() == 5:
        x = x.sum((-1, -2, -3))
    loss = (x - y).abs() / (x + y + E)
    v = torch.max((1 - tolerance) * y, y.clamp(max=100))
    loss = loss - (v - y).abs() / (v + y + E)
    return loss.clamp(0, 1)


def volume_loss_basic(out, chosen_segment_mask, tumor_volumes, labels, unk_voxels, classes,
                      dilation_segment: int = 31, dilation_unk: int = 7, tolerance: float = 0.1,
                      sigmoid: bool = True, class_weights=None) -> Dict[str, torch.Tensor]:
    """losses_foundation.py:250-349."""
    assert tumor_volumes.dim() == 2 and out.dim() == 5
    assert chosen_segment_mask.shape == out.shape == unk_voxels.shape == labels.shape
    if class_weights is not None:
        class_weights = class_weights.repeat(1, 1, *out.shape[2:])
    out = get_lesion_channels(out, classes)
    csm = get_lesion_channels(chosen_segment_mask, classes)
    labels = get_lesion_channels(labels, classes)
    if sigmoid:
        out = torch.sigmoid(out)
    csm = dilate_volume(csm, dilation_segment)
    # (the reference also dilates unk_voxels by 7 here, :310, but the result never reaches the loss)
    per_voxel_positives = (labels.sum((-1, -2, -3), keepdim=True) > 0).float()
    out = out * (1 - per_voxel_positives)
    if class_weights is not None:
        class_weights = get_lesion_channels(class_weights, classes).mean(dim=(-1, -2, -3))
    in_segment = out * csm
    report_volume = tumor_volumes.sum(-1).unsqueeze(-1).repeat(1, csm.shape[1])
    gate = (csm.sum(dim=(-1, -2, -3)) > 0).float()
    loss = dice_based_volume_loss(in_segment, report_volume * gate, tolerance=tolerance, E=500)
    if class_weights is not None:
        loss = loss * class_weights
    return {"dice_volume_loss": loss.mean()}


# --------------------------------------------------------------------------------------------
# Ball Loss
# --------------------------------------------------------------------------------------------
def gwrp_weights(x: torch.Tensor, N, c: float = 0.5, hard_cutoff: bool = True) -> torch.Tensor:
    """GlobalWeightedRankPooling(return_weights=True), losses_foundation.py:442-537: rank-decay
    weights d^rank (d = (1-c)^(1/N)), normalised, zero beyond rank N, scattered back to voxel order."""
    shape = x.shape
    flat = x.reshape(1, 1, -1)
    L = flat.shape[-1]
    _, order = torch.sort(flat, dim=-1, descending=True)
    n_t = N.to(x.device).float() if torch.is_tensor(N) else torch.tensor(float(N), device=x.device)
    n_t = n_t.reshape(1, 1).clamp(min=1)
    d = ((1 - c) ** (1.0 / n_t)).unsqueeze(-1)
    idx = torch.arange(L, dtype=torch.float32, device=x.device).view(1, 1, L)
    raw = d ** idx
    w = raw / raw.sum(-1, keepdim=True)
    if hard_cutoff:
        w = w * (idx < n_t.unsqueeze(-1)).float()
        w = w / w.sum(-1, keepdim=True)
    inv = order.argsort(dim=-1)
    return w.gather(-1, inv).reshape(shape)


def insert_ball(shape, center, diameter, margin, like: torch.Tensor) -> torch.Tensor:
    """losses_foundation.py:1336-1385: binary ball of diameter*(1+margin) clipped to the volume."""
    ball = create_ball_kernel(diameter * (1 + margin), gaussian=False)
    vol = torch.zeros(shape, dtype=like.dtype, device=like.device)
    half = ball.shape[-1] // 2
    sl_v, sl_b = [], []
    for c, size in zip(center, shape):
        lo, hi = max(0, c - half), min(size, c + half + 1)
        b0 = 0 if c - half >= 0 else -(c - half)
        sl_v.append(slice(lo, hi))
        sl_b.append(slice(b0, b0 + (hi - lo)))
    vol[tuple(sl_v)] = ball[tuple(sl_b)].to(vol)
    return vol


def isolate_tumor(x: torch.Tensor, diameter, gaussian, gaussian_std, tumor_volume, diameter_margin=0.5,
                  volume_margin=0.5):
    """losses_foundation.py:1387-1532 for a 3-D input x >= 0: best ball centre by Gaussian-ball
    correlation + argmax, ball mask (grown if clipped), top-{t, t_small, t_big} voxels inside it,
    up to 6 fallback dilations.  Returns (mask, mask_small, mask_big)."""
    assert x.dim() == 3
    diameter = int(np.round(diameter).astype(int))
    tumor_volume = int(np.round(tumor_volume).astype(int))
    if diameter % 2 == 0:
        diameter += 1
    kernel = create_ball_kernel(diameter, gaussian, gaussian_std).to(x)
    support = int((kernel > 0).sum().item())
    if tumor_volume > 100000:
        assert tumor_volume <= support * 1.2
    if support > tumor_volume:
        tumor_volume = support - 1
    score = F.conv3d(x[None, None], kernel[None, None], padding=kernel.shape[-1] // 2)[0, 0]
    center = np.unravel_index(int(torch.argmax(score).item()), score.shape)
    center = tuple(int(c) for c in center)
    ball = insert_ball(x.shape, center, diameter, diameter_margin, x)
    new_dim = diameter
    while ball.sum() < tumor_volume:
        old = new_dim
        new_dim = int(np.round(new_dim * 1.1))
        if old == new_dim:
            new_dim += 1
        if new_dim % 2 == 0:
            new_dim += 1
        if new_dim >= max(x.shape):
            break
        ball = insert_ball(x.shape, center, new_dim, diameter_margin, x)
    if tumor_volume < 50 ** 3:
        assert ball.sum() > tumor_volume * 0.5
    if tumor_volume > 6 ** 3:
        assert ball.sum() < tumor_volume * ((1 + diameter_margin) ** 3) * 2
    assert (x >= 0).all()
    flat = (x * ball).reshape(-1)
    t = min(flat.shape[-1] - 1, tumor_volume)
    t_small = max(int(t * (1 - min(0.5, volume_margin))), min(100, tumor_volume))
    t_big = min(flat.shape[-1] - 1, int(tumor_volume * (1 + volume_margin)))
    masks = []
    for k in (t, t_small, t_big):
        m = torch.zeros_like(flat)
        m[torch.topk(flat, k).indices] = 1
        masks.append(m.view_as(x) * ball)
    mask, small, big = masks
    iters = 0
    while tumor_volume < 50 ** 3 and mask.sum() < tumor_volume * 0.7:
        if iters > 5:
            return mask, small, big
        mask = dilate_volume(mask, 7) * ball
        small = dilate_volume(small, 7) * ball
        big = dilate_volume(big, 7) * ball
        iters += 1
    if tumor_volume < 50 ** 3:
        assert mask.sum() > tumor_volume * 0.5
    if tumor_volume > 5 ** 3:
        assert mask.sum() < tumor_volume * ((1 + volume_margin) ** 3) * 3
    return mask, small, big


def ball_loss(out, labels, unk_voxels, chosen_segment_mask, tumor_volumes, tumor_diameters, classes,
              apply_dice_loss: bool, diameter_margin=0.2, volume_margin=0.2, gaussian=True, gaussian_std=1.5,
              gwrp=True, gwrp_concentration=0.5, dilation_for_background=7, subseg_dilation=31, unk_dilation=1,
              standard_ce=False, class_weights=None, debug: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """losses_foundation.py:1537-1864 with sigmoid=True, single_class=False, use_small_pseudo_mask=True
    (the only configuration calculate_loss uses, :926-932).  `debug`, if given, receives the discrete
    intermediates (pseudo masks, centres) for bit-exact comparison."""
    assert tumor_volumes.dim() == 2 and out.dim() == 5
    assert chosen_segment_mask.shape == out.shape == unk_voxels.shape == labels.shape
    if class_weights is not None:
        class_weights = class_weights.repeat(out.shape[0], 1, *out.shape[2:])
        class_weights = get_lesion_channels(class_weights, classes)
    out = get_lesion_channels(out, classes)
    csm = dilate_volume(get_lesion_channels(chosen_segment_mask, classes), subseg_dilation)
    unk = dilate_volume(get_lesion_channels(unk_voxels, classes), unk_dilation)
    labels = get_lesion_channels(labels, classes)
    to_penalize = ((torch.ones_like(out) * (1 - unk) * (1 - labels) + csm) > 0).float()
    losses, losses_dice = [], []
    for b in range(out.shape[0]):
        assert torch.equal(tumor_diameters[b].sum(-1) > 0, tumor_volumes[b] > 0)
        x = out[b]
        seg = csm[b]
        assert (seg.sum((-1, -2, -3)) > 0).float().sum() <= 1
        if seg.sum() == 0 or tumor_volumes[b].sum() == 0:  # no report for this sample (:1625-1661)
            zero = torch.zeros_like(x)
            l = F.binary_cross_entropy_with_logits(x, zero, reduction="none") * to_penalize[b]
            if class_weights is not None:
                l = l * class_weights[b]
            losses.append(l.mean())
            if apply_dice_loss:
                w = class_weights[b] if class_weights is not None else None
                losses_dice.append(dice_loss_multiclass(x, zero, to_penalize[b], sigmoid=True, class_weights=w).mean())
            continue
        c_sel = next(c for c in range(x.shape[0]) if seg[c].sum() > 0)
        xc = x[c_sel]
        penalize = to_penalize[b][c_sel]
        cw = class_weights[b][c_sel] if class_weights is not None else None
        seg1 = seg.sum(0)
        vols, dias = tumor_volumes[b], tumor_diameters[b]
        order = torch.argsort(vols, descending=True)
        order = order[vols[order] > 0]
        x_iter = torch.sigmoid(xc) * seg1
        smalls, bigs, centers = [], [], []
        for ti in order:
            vol = vols[ti].item()
            dmax = torch.max(dias[ti]).item()
            assert dmax > 0 and vol > 0
            if dmax <= 1:
                dmax = 3
            if vol <= 1:
                vol = 9
            m, ms, mb = isolate_tumor(x_iter, dmax, gaussian, gaussian_std, vol, diameter_margin, volume_margin)
            smalls.append(ms)
            bigs.append(mb)
            x_iter = x_iter * (1 - m)
        pseudo = (torch.stack(smalls).sum(0) > 0).float()
        dilated = (torch.stack(bigs).sum(0) > 0).float()
        if dilation_for_background > 0:
            dilated = dilate_volume(dilated, dilation_for_background)
        border = ((dilated - pseudo) > 0).float()
        penalize = penalize * (1 - border)
        if debug is not None:
            debug.setdefault("pseudo", []).append(pseudo.clone())
            debug.setdefault("dilated", []).append(dilated.clone())
            debug.setdefault("penalize", []).append(penalize.clone())
        bce = F.binary_cross_entropy_with_logits(xc, pseudo, reduction="none") * penalize
        if apply_dice_loss:
            dice = dice_loss_multiclass(xc, pseudo, penalize, sigmoid=True, class_weights=cw)
        if not standard_ce:
            if gwrp:
                assert pseudo.sum() > 0
                fw = gwrp_weights(torch.sigmoid(xc) * pseudo + pseudo, N=pseudo.sum(), c=gwrp_concentration,
                                  hard_cutoff=True)
                assert 0.95 < fw.sum() < 1.05
                fw = fw * pseudo.sum() * pseudo
                fg = bce * fw
            else:
                fg = bce * pseudo
            bg = bce * (1 - dilated)
            if cw is not None:
                fg, bg = fg * cw, bg * cw
            losses.append(fg.mean() + bg.mean())
        else:
            losses.append((bce * cw).mean() if cw is not None else bce.mean())
        if apply_dice_loss:
            losses_dice.append(dice.mean())
    bce_mean = torch.stack(losses).mean()
    return {"ball_loss_bce": bce_mean,
            "ball_loss_dice": torch.stack(losses_dice).mean() if apply_dice_loss else torch.zeros_like(bce_mean)}


# --------------------------------------------------------------------------------------------
# calculate_loss
# --------------------------------------------------------------------------------------------
def default_args(**kw) -> SimpleNamespace:
    """The args fields calculate_loss reads (:726-731) with train_ddp.py's defaults (:429-465) and
    aux_weight of config/abdomenatlas_ufo/medformer_3d.yaml:42."""
    a = dict(loss="ball_dice_last", aux_weight=[0.5, 0.5], seg_loss=1.0, report_volume_loss_basic=0.1,
             volume_loss_tolerance=0.2, ball_bce_weight=1.0, ball_dice_weight=1.0, ball_volume_margin=0.2,
             multi_ch_tumor=False, stardard_ce_ball=False, classification_branch=False)
    a.update(kw)
    return SimpleNamespace(**a)


def calculate_loss(model_output, label, unk_voxels, args, matcher, chosen_segment_mask, tumor_volumes_report,
                   tumor_diameters, classes, input_tensor=None, class_weights=None, **_ignored) -> Dict[str, torch.Tensor]:
    """losses_foundation.py:685-1076 restricted to the hot path: segmentation loss
    per head, Volume / Ball loss selection by args.loss, aux_weight accumulation, 'overall' sum."""
    result = model_output["segmentation"]
    if args.multi_ch_tumor or args.classification_branch:
        raise NotImplementedError("baseline branches are out of scope")
    if chosen_segment_mask is not None and chosen_segment_mask.sum() > 0:
        for b in range(chosen_segment_mask.shape[0]):
            if chosen_segment_mask[b].sum() > 0 and (unk_voxels[b].sum() == 0 or tumor_volumes_report[b].sum() == 0):
                raise ValueError("report sample without unk_voxels / tumor volumes")
    heads = list(result) if isinstance(result, (tuple, list)) else [result]
    deep = isinstance(result, (tuple, list))
    assert len(classes) == label.shape[1] == heads[0].shape[1]
    if class_weights is not None and torch.equal(class_weights, torch.ones_like(class_weights)):
        class_weights = None
    if class_weights is not None:
        class_weights = class_weights.to(label.device)[:, :, None, None, None]
    if unk_voxels is not None:
        known = get_known_voxels(unk_voxels)
        assert torch.equal((known * label).float().sum(), label.float().sum())
    else:
        known = torch.ones(label.shape).type_as(label)
    loss_seg_total = 0
    loss_report = 0
    for j, r in enumerate(heads):
        aw = args.aux_weight[j] if deep else 1.0
        assert not torch.isnan(r).any()
        if args.report_volume_loss_basic > 0:
            use_ball = any(k in args.loss for k in ("ball", "dynamic", "dll")) and not (deep and j != 0 and "last" in args.loss)
            if use_ball:
                lr = ball_loss(out=r, labels=label, unk_voxels=unk_voxels, chosen_segment_mask=chosen_segment_mask,
                               tumor_volumes=tumor_volumes_report, tumor_diameters=tumor_diameters, classes=classes,
                               apply_dice_loss=("dice" in args.loss), standard_ce=args.stardard_ce_ball,
                               class_weights=class_weights, diameter_margin=args.ball_volume_margin,
                               volume_margin=args.ball_volume_margin)
                if "both" in args.loss:
                    lr.update(volume_loss_basic(r, chosen_segment_mask, tumor_volumes_report, label, unk_voxels, classes,
                                                class_weights=class_weights, tolerance=args.volume_loss_tolerance))
            else:
                lr = volume_loss_basic(r, chosen_segment_mask, tumor_volumes_report, label, unk_voxels, classes,
                                       class_weights=class_weights, tolerance=args.volume_loss_tolerance)
        else:
            lr = torch.tensor(0).type_as(r)
        loss_seg_total = loss_seg_total + aw * args.seg_loss * seg_loss(r, label, known, class_weights)
        if not isinstance(lr, dict):
            loss_report = loss_report + aw * args.report_volume_loss_basic * lr
        else:
            if isinstance(loss_report, int):
                loss_report = {}
            for k, v in lr.items():
                wgt = args.ball_bce_weight if k == "ball_loss_bce" else (args.ball_dice_weight if k == "ball_loss_dice" else 1)
                term = aw * args.report_volume_loss_basic * wgt * v
                loss_report[k] = loss_report[k] + term if k in loss_report else term
    loss = {"segmentation": loss_seg_total}
    if isinstance(loss_report, dict):
        loss.update(loss_report)
    else:
        loss["report"] = loss_report
    overall = 0
    for v in loss.values():
        overall = overall + v
    loss["overall"] = overall
    if torch.isnan(overall).any():
        raise ValueError("loss is nan")
    assert overall.requires_grad
    return loss
