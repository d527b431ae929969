"""Report-supervised losses on the librsuper_b200.so kernels — host-side mirror of
rsuper_train/training/losses_foundation.py: volume_loss_basic (:250-349), dice_based_volume_loss (:352-395),
ball_loss (:1537-1864), isolate_tumor (:1387-1532), insert_ball (:1336-1385), GlobalWeightedRankPooling (:442-537)
and the report branch of calculate_loss (:899-1076).

Every per-voxel step runs in a CUDA kernel (csrc/report_loss.cu, seg_loss.cu, morph.cu); what stays here is the
reference's own scalar control flow (the <= 10-tumour loop, ball geometry, formulas on [B, L] values).  Like the
reference (`.item()` at :1404, 1419, 1500 ...) the tumour loop reads back a few scalars per tumour.

Deterministic selection rule: voxels are ranked by (value descending, flat index ascending) and only
voxels with a positive score are eligible for the top-k masks (torch.topk's choice among exact ties is unspecified).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops

_SUFFIXES = ("lesion", "cyst", "pdac", "pnet")
# isolate_tumor's ball correlation through the rows -> discs -> planes decomposition (csrc/report_loss.cu); False = the direct
# tap-list kernel (kept: it is the definition the decomposition is tested against)
SEPARABLE_CORRELATION = True


def lesion_channels(classes: Sequence[str]) -> List[List[int]]:
    """Channel indices of every lesion group (get_lesion_channels, :204-248): channels whose name contains
    lesion | cyst | pdac | pnet, grouped by the name up to and including the suffix ('pancreatic' -> 'pancreas'); a group
    with several members (e.g. 'liver_lesion_1', 'liver_lesion_2') is max-merged.  Mirrors the reference loop, including
    that a name matching several suffixes joins one group per matching suffix."""
    names: List[str] = []
    groups: Dict[str, List[int]] = {}
    for i, cl in enumerate(classes):
        for suffix in _SUFFIXES:
            if suffix in cl:
                key = cl[:cl.index("_" + suffix) + len("_" + suffix)].replace("pancreatic", "pancreas")
                if key not in groups:
                    groups[key] = []
                    names.append(key)
                groups[key].append(i)
    if not names:
        raise ValueError("no lesion channel in `classes`")
    return [groups[k] for k in names]


def _u8(t: torch.Tensor) -> torch.Tensor:
    return (t if t.dtype == torch.uint8 else t.to(torch.uint8)).contiguous()


def _odd_ceil(v: float) -> int:
    k = math.ceil(v)
    return k + 1 if k % 2 == 0 else k


def _ball_geometry(diameter: float):
    """create_ball_kernel (:1161-1232): (grid half-width, radius^2) of the ball of odd(ceil(d)) in odd(ceil(1.2*that))."""
    d_odd = _odd_ceil(diameter)
    size = _odd_ceil(1.2 * d_odd)
    radius = d_odd / 2.0
    return size // 2, np.float32(radius * radius)


_TAPS: Dict[tuple, tuple] = {}


def _gauss_ball_taps(diameter: int, gaussian: bool, std: float, device):
    """Taps {dz, dy, dx, weight bits} of create_ball_kernel(d, gaussian, std), its support and grid half-width."""
    key = (diameter, gaussian, float(std), str(device))
    hit = _TAPS.get(key)
    if hit is not None:
        return hit
    half, r2 = _ball_geometry(diameter)
    c = np.arange(-half, half + 1, dtype=np.float32)
    d2 = c[:, None, None] ** 2 + c[None, :, None] ** 2 + c[None, None, :] ** 2
    inside = d2 <= r2
    if gaussian:
        radius = np.float32(_odd_ceil(diameter) / 2.0)
        sigma = np.float32(std) * radius
        k = np.exp(-d2 / (np.float32(2.0) * sigma ** 2)).astype(np.float32) * inside
        k = (k / k.sum(dtype=np.float32)).astype(np.float32)
    else:
        k = inside.astype(np.float32)
    zz, yy, xx = np.nonzero(k > 0)
    taps = np.empty((len(zz), 4), dtype=np.int32)
    taps[:, 0], taps[:, 1], taps[:, 2] = zz - half, yy - half, xx - half
    taps[:, 3] = k[zz, yy, xx].view(np.int32)
    hit = (torch.from_numpy(taps).to(device), int(len(zz)), half)
    _TAPS[key] = hit
    return hit


_SEP: Dict[tuple, tuple] = {}


def _gauss_ball_sep(diameter: int, std: float, device):
    """The Gaussian ball of create_ball_kernel(d, gaussian=True, std) in separable form: g(d) for d = 0..R (fp32), the table
    w(a, b) = largest dx with a^2 + b^2 + dx^2 <= r^2 (or -1), and R = floor(radius)."""
    key = (diameter, float(std), str(device))
    hit = _SEP.get(key)
    if hit is not None:
        return hit
    _, r2 = _ball_geometry(diameter)
    radius = np.float32(_odd_ceil(diameter) / 2.0)
    reach = int(math.floor(float(radius)))
    c = np.arange(0, reach + 1, dtype=np.float32)
    sigma = np.float32(std) * radius
    g = np.exp(-(c * c) / (np.float32(2.0) * sigma ** 2)).astype(np.float32)
    d2 = c[:, None, None] ** 2 + c[None, :, None] ** 2 + c[None, None, :] ** 2       # same fp32 test as the tap list: d2 <= r2
    inside = d2 <= r2
    wtab = np.where(inside.any(-1), inside.sum(-1) - 1, -1).astype(np.int32)         # runs are contiguous from dx = 0
    wflat = np.ascontiguousarray(wtab.reshape(-1))
    hit = (torch.from_numpy(g).to(device), torch.from_numpy(wflat.copy()).to(device), reach, np.ascontiguousarray(g), wflat)
    _SEP[key] = hit
    return hit


def _clipped_ball_count(shape, center, half: int, r2) -> int:
    """insert_ball(...).sum() without touching the GPU: ball voxels that fall inside the volume."""
    axes = []
    for c, size in zip(center, shape):
        o = np.arange(max(0, c - half), min(size, c + half + 1), dtype=np.float32) - np.float32(c)
        axes.append(o * o)
    d2 = axes[0][:, None, None] + axes[1][None, :, None] + axes[2][None, None, :]
    return int((d2 <= r2).sum())


# ------------------------------------------------------------------------------------------------
# Volume loss
# ------------------------------------------------------------------------------------------------
class _MaskedSigmoidSum(torch.autograd.Function):
    """sums[b, l] = scale[b, l] * sum_v sigmoid(logits[b, ch_l, v]) * mask[b, l, v]  (in_segment.sum, :330-340)."""

    @staticmethod
    def forward(ctx, logits, row_map, mask_rows, scale, n_b, n_l):
        lg = logits.detach().contiguous().float()
        v = lg[0, 0].numel()
        x_rows = ops.rows_gather(lg, row_map, n_b * n_l, v)
        sums = ops.masked_sigmoid_sum(x_rows, mask_rows, scale, v)
        ctx.keep = (x_rows, row_map, mask_rows, scale, lg, v)
        return sums.view(n_b, n_l)

    @staticmethod
    def backward(ctx, g):
        x_rows, row_map, mask_rows, scale, lg, v = ctx.keep
        shape = lg.shape
        dx_rows = torch.empty_like(x_rows)
        ops.masked_sigmoid_grad(x_rows, mask_rows, scale, g.detach().reshape(-1).float().contiguous(), dx_rows, v)
        dl = torch.zeros(shape, dtype=torch.float32, device=x_rows.device)
        ops.rows_scatter_add(dx_rows, row_map, dl, v, x=lg)
        ctx.keep = None
        return dl, None, None, None, None, None


def dice_based_volume_loss(x, y, tolerance: float = 0.1, E: float = 500.0):
    """(:352-395, cross_entropy=False) on [B, L] values."""
    loss = (x - y).abs() / (x + y + E)
    v = torch.max((1 - tolerance) * y, y.clamp(max=100))
    loss = loss - (v - y).abs() / (v + y + E)
    return loss.clamp(0, 1)


_ROW_MAPS: Dict[tuple, torch.Tensor] = {}


def _row_map(n_b: int, n_c: int, groups: Sequence[Sequence[int]], device) -> torch.Tensor:
    """Source rows of the [B * L] lesion rows in a [B * C, V] view: int32 [B * L] when every group has one channel, else
    int32 [B * L, G] padded with -1 (ops.rows_gather max-merges).  Cached: the table is a host -> device copy."""
    key = (n_b, n_c, tuple(tuple(g) for g in groups), str(device))
    hit = _ROW_MAPS.get(key)
    if hit is None:
        width = max(len(g) for g in groups)
        rows = [[b * n_c + c for c in g] + [-1] * (width - len(g)) for b in range(n_b) for g in groups]
        t = torch.tensor(rows, dtype=torch.int32)
        hit = _ROW_MAPS[key] = (t[:, 0].contiguous() if width == 1 else t.contiguous()).to(device)
    return hit


def _group_weights(class_weights, n_c: int, groups, n_b: int, n_l: int):
    """class weights of the merged lesion channels ([B, L]): get_lesion_channels(class_weights) = max over the group."""
    cw = class_weights.reshape(class_weights.shape[0], n_c).float()
    return torch.stack([cw[:, g].max(dim=1).values for g in groups], dim=1).expand(n_b, n_l).contiguous()


def volume_loss_basic(out, chosen_segment_mask, tumor_volumes, labels, unk_voxels, classes, dilation_segment: int = 31,
                      tolerance: float = 0.1, class_weights=None) -> Dict[str, torch.Tensor]:
    assert tumor_volumes.dim() == 2 and out.dim() == 5
    assert chosen_segment_mask.shape == out.shape == unk_voxels.shape == labels.shape
    n_b, n_c = out.shape[:2]
    d, h, w_ = out.shape[2:]
    v = d * h * w_
    chans = lesion_channels(classes)
    n_l = len(chans)
    rm = _row_map(n_b, n_c, chans, out.device)
    csm = ops.rows_gather(_u8(chosen_segment_mask), rm, n_b * n_l, v)
    dcsm = ops.dilate_ball(csm.view(n_b * n_l, d, h, w_), dilation_segment).view(n_b * n_l, v)
    lab_cnt = ops.u8_row_count(ops.rows_gather(_u8(labels), rm, n_b * n_l, v), v)
    scale = (lab_cnt == 0).float()                       # 1 - per_voxel_positives (:322-324)
    gate = (ops.u8_row_count(dcsm, v) > 0).float().view(n_b, n_l)
    sums = _MaskedSigmoidSum.apply(out, rm, dcsm, scale, n_b, n_l)
    report_volume = tumor_volumes.float().sum(-1, keepdim=True).expand(n_b, n_l)
    loss = dice_based_volume_loss(sums, report_volume * gate, tolerance=tolerance, E=500.0)
    if class_weights is not None:
        loss = loss * _group_weights(class_weights, n_c, chans, n_b, n_l)
    return {"dice_volume_loss": loss.mean()}


# ------------------------------------------------------------------------------------------------
# Ball loss
# ------------------------------------------------------------------------------------------------
def isolate_tumor(x_iter: torch.Tensor, diameter, gaussian: bool, gaussian_std: float, tumor_volume, diameter_margin=0.5,
                  volume_margin=0.5):
    """isolate_tumor (:1387-1532) for a 3-D fp32 device volume x_iter >= 0.  Returns (mask, small, big) uint8."""
    assert x_iter.dim() == 3 and x_iter.dtype == torch.float32 and x_iter.is_contiguous()
    shape = tuple(x_iter.shape)
    v = x_iter.numel()
    diameter = int(np.round(diameter).astype(int))
    tumor_volume = int(np.round(tumor_volume).astype(int))
    if diameter % 2 == 0:
        diameter += 1
    taps, support, khalf = _gauss_ball_taps(diameter, gaussian, gaussian_std, x_iter.device)
    if tumor_volume > 100000:
        assert tumor_volume <= support * 1.2
    if support > tumor_volume:
        tumor_volume = support - 1
    if gaussian and SEPARABLE_CORRELATION:
        g1d, wtab, reach, g_host, w_host = _gauss_ball_sep(diameter, gaussian_std, x_iter.device)
        key = int(ops.ball_correlate_argmax_sep(x_iter, g1d, wtab, reach, g_host, w_host).item())
    else:
        key = int(ops.ball_correlate_argmax(x_iter, taps, khalf).item())
    flat_idx = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    center = tuple(int(c) for c in np.unravel_index(flat_idx, shape))
    half, r2 = _ball_geometry(diameter * (1 + diameter_margin))
    ball_sum = _clipped_ball_count(shape, center, half, r2)
    new_dim = diameter
    while ball_sum < tumor_volume:
        old = new_dim
        new_dim = int(np.round(new_dim * 1.1))
        if old == new_dim:
            new_dim += 1
        if new_dim % 2 == 0:
            new_dim += 1
        if new_dim >= max(shape):
            break
        half, r2 = _ball_geometry(new_dim * (1 + diameter_margin))
        ball_sum = _clipped_ball_count(shape, center, half, r2)
    if tumor_volume < 50 ** 3:
        assert ball_sum > tumor_volume * 0.5
    if tumor_volume > 6 ** 3:
        assert ball_sum < tumor_volume * ((1 + diameter_margin) ** 3) * 2
    t = min(v - 1, tumor_volume)
    t_small = max(int(t * (1 - min(0.5, volume_margin))), min(100, tumor_volume))
    t_big = min(v - 1, int(tumor_volume * (1 + volume_margin)))
    dev = x_iter.device
    max_cand = ball_sum
    cand = torch.empty((max_cand, 2), dtype=torch.int32, device=dev)
    n_cand = torch.empty(1, dtype=torch.int32, device=dev)
    ball = torch.empty(shape, dtype=torch.uint8, device=dev)
    masks = torch.zeros((3,) + shape, dtype=torch.uint8, device=dev)
    ops.ball_candidates(x_iter, None, 0, center, half, r2, cand, n_cand, ball)
    ops.ball_rank_select(cand, n_cand, max_cand, (t, t_small, t_big), (masks[0], masks[1], masks[2]))
    mask, small, big = masks[0], masks[1], masks[2]
    mask_sum = min(t, int(n_cand.item()))
    iters = 0
    while tumor_volume < 50 ** 3 and mask_sum < tumor_volume * 0.7:
        if iters > 5:
            return mask, small, big
        mask = ops.u8_binary(ops.dilate_ball(mask, 7), ball, ops.U8_AND)
        small = ops.u8_binary(ops.dilate_ball(small, 7), ball, ops.U8_AND)
        big = ops.u8_binary(ops.dilate_ball(big, 7), ball, ops.U8_AND)
        mask_sum = int(ops.u8_row_count(mask, v).item())
        iters += 1
    if tumor_volume < 50 ** 3:
        assert mask_sum > tumor_volume * 0.5
    if tumor_volume > 5 ** 3:
        assert mask_sum < tumor_volume * ((1 + volume_margin) ** 3) * 3
    return mask, small, big


class _BallLoss(torch.autograd.Function):
    """ball_loss (:1537-1864) with sigmoid=True, single_class=False, use_small_pseudo_mask=True — the configuration
    calculate_loss uses (:926-932).  Returns (ball_loss_bce, ball_loss_dice); `debug` receives the discrete intermediates."""

    @staticmethod
    def forward(ctx, logits, labels, unk_voxels, chosen_segment_mask, volumes_h, diameters_h, classes, apply_dice, diameter_margin,
                volume_margin, gaussian_std, gwrp_concentration, dilation_for_background, subseg_dilation, unk_dilation,
                standard_ce, class_weights, debug):
        lg = logits.detach().contiguous().float()
        n_b, n_c = lg.shape[:2]
        d, h, w_ = lg.shape[2:]
        v = d * h * w_
        dev = lg.device
        chans = lesion_channels(classes)
        n_l = len(chans)
        rm = _row_map(n_b, n_c, chans, dev)
        x_rows = ops.rows_gather(lg, rm, n_b * n_l, v)
        csm = ops.rows_gather(_u8(chosen_segment_mask), rm, n_b * n_l, v)
        csm = ops.dilate_ball(csm.view(n_b * n_l, d, h, w_), subseg_dilation).view(n_b * n_l, v)
        unk = ops.rows_gather(_u8(unk_voxels), rm, n_b * n_l, v)
        unk = ops.dilate_ball(unk.view(n_b * n_l, d, h, w_), unk_dilation).view(n_b * n_l, v)
        lab = ops.rows_gather(_u8(labels), rm, n_b * n_l, v)
        # to_penalize = ((1-unk)*(1-labels) + csm) > 0   (:1605)
        to_pen = ops.u8_binary(ops.u8_binary(unk, lab, ops.U8_NOR), csm, ops.U8_OR)
        seg_cnt = ops.u8_row_count(csm, v).view(n_b, n_l).cpu().numpy()  # `seg.sum(...) > 0` tests of the reference
        cw = None
        if class_weights is not None:
            cw = _group_weights(class_weights, n_c, chans, n_b, n_l)
        states, bce_terms, dice_terms = [], [], []
        for b in range(n_b):
            vols, dias = volumes_h[b], diameters_h[b]
            assert np.array_equal(dias.sum(-1) > 0, vols > 0)
            assert (seg_cnt[b] > 0).sum() <= 1
            rows = slice(b * n_l, (b + 1) * n_l)
            if seg_cnt[b].sum() == 0 or vols.sum() == 0:   # no report for this sample (:1625-1661)
                zeros = torch.zeros((1, n_l, d, h, w_), dtype=torch.uint8, device=dev)
                st = ops.seg_loss_forward(x_rows[rows].view(1, n_l, d, h, w_), zeros, to_pen[rows].view(1, n_l, d, h, w_),
                                          None if cw is None else cw[b].reshape(1, n_l).contiguous())
                states.append((st, rows))
                bce_terms.append(st.loss_out[1])
                dice_terms.append(st.loss_out[2])
                continue
            c_sel = int(np.nonzero(seg_cnt[b] > 0)[0][0])
            r = b * n_l + c_sel
            xc = x_rows[r].view(d, h, w_)
            seg1 = csm[r]
            x_iter = ops.ball_prepare(xc, seg1.view(d, h, w_))
            order = [int(i) for i in np.argsort(-vols, kind="stable") if vols[i] > 0]
            pseudo = torch.zeros(v, dtype=torch.uint8, device=dev)
            big_u = torch.zeros(v, dtype=torch.uint8, device=dev)
            for ti in order:
                vol, dmax = float(vols[ti]), float(dias[ti].max())
                assert dmax > 0 and vol > 0
                if dmax <= 1:
                    dmax = 3
                if vol <= 1:
                    vol = 9
                m, ms, mb = isolate_tumor(x_iter, dmax, True, gaussian_std, vol, diameter_margin, volume_margin)
                ops.u8_binary(pseudo, ms.reshape(-1), ops.U8_OR, out=pseudo)
                ops.u8_binary(big_u, mb.reshape(-1), ops.U8_OR, out=big_u)
                ops.ball_remove(x_iter, m.reshape(-1) if m.is_contiguous() else m.contiguous().reshape(-1))
            dilated = big_u
            if dilation_for_background > 0:
                dilated = ops.dilate_ball(big_u.view(1, d, h, w_), dilation_for_background).view(v)
            border = ops.u8_binary(dilated, pseudo, ops.U8_ANDNOT)
            penalize = ops.u8_binary(to_pen[r], border, ops.U8_ANDNOT)
            if debug is not None:
                debug.setdefault("pseudo", []).append(pseudo.view(d, h, w_).clone())
                debug.setdefault("dilated", []).append(dilated.view(d, h, w_).clone())
                debug.setdefault("penalize", []).append(penalize.view(d, h, w_).clone())
            wmap = None
            if not standard_ce:
                n_pseudo = int(ops.u8_row_count(pseudo, v).item())
                assert n_pseudo > 0
                wmap = torch.zeros(v, dtype=torch.float32, device=dev)
                cand = torch.empty((n_pseudo, 2), dtype=torch.int32, device=dev)
                n_cand = torch.empty(1, dtype=torch.int32, device=dev)
                ops.ball_candidates(xc, pseudo, 1, (0, 0, 0), 0, 0.0, cand, n_cand, None)
                ops.ball_rank_gwrp(cand, n_cand, n_pseudo, gwrp_concentration, wmap)
                ops.ball_weight_map(wmap, pseudo, dilated)
                wmap = wmap.view(1, 1, d, h, w_)
            st = ops.seg_loss_forward(xc.view(1, 1, d, h, w_), pseudo.view(1, 1, d, h, w_), penalize.view(1, 1, d, h, w_),
                                      None if cw is None else cw[b, c_sel].reshape(1, 1).contiguous(), wmap)
            states.append((st, slice(r, r + 1)))
            bce_terms.append(st.loss_out[1])
            dice_terms.append(st.loss_out[2])
        ctx.states, ctx.meta = states, (x_rows.shape, rm, lg, v, n_b, apply_dice)
        bce = torch.stack(bce_terms).mean()
        dice = torch.stack(dice_terms).mean() if apply_dice else torch.zeros_like(bce)
        return bce, dice

    @staticmethod
    def backward(ctx, g_bce, g_dice):
        rows_shape, rm, lg, v, n_b, apply_dice = ctx.meta
        shape = lg.shape
        dev = g_bce.device
        dx_rows = torch.zeros(rows_shape, dtype=torch.float32, device=dev)
        scales = torch.stack([g_bce.detach().float().reshape(()), (g_dice.detach().float().reshape(()) if apply_dice
                                                                  else torch.zeros((), device=dev))]) / n_b
        scales = scales.contiguous()
        for st, rows in ctx.states:
            ops.seg_loss_backward(st, scales, dx_rows[rows])
        dl = torch.zeros(shape, dtype=torch.float32, device=dev)
        ops.rows_scatter_add(dx_rows, rm, dl, v, x=lg)
        ctx.states = None
        return (dl,) + (None,) * 17


def ball_loss(out, labels, unk_voxels, chosen_segment_mask, tumor_volumes, tumor_diameters, classes, apply_dice_loss: bool,
              diameter_margin=0.2, volume_margin=0.2, gaussian=True, gaussian_std=1.5, gwrp=True, gwrp_concentration=0.5,
              dilation_for_background=7, subseg_dilation=31, unk_dilation=1, standard_ce=False, class_weights=None,
              debug: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    assert tumor_volumes.dim() == 2 and out.dim() == 5
    assert chosen_segment_mask.shape == out.shape == unk_voxels.shape == labels.shape
    if not gaussian or not gwrp:
        raise NotImplementedError("rsuper_b200.ball_loss implements the reference defaults gaussian=True, gwrp=True")
    vol_h = tumor_volumes.detach().float().cpu().numpy()
    dia_h = tumor_diameters.detach().float().cpu().numpy()
    bce, dice = _BallLoss.apply(out, labels, unk_voxels, chosen_segment_mask, vol_h, dia_h, tuple(classes), bool(apply_dice_loss),
                                float(diameter_margin), float(volume_margin), float(gaussian_std), float(gwrp_concentration),
                                int(dilation_for_background), int(subseg_dilation), int(unk_dilation), bool(standard_ce),
                                class_weights, debug)
    return {"ball_loss_bce": bce, "ball_loss_dice": dice}


# ------------------------------------------------------------------------------------------------
# calculate_loss, report branch (:899-1076)
# ------------------------------------------------------------------------------------------------
def calculate_loss_with_reports(heads, deep, label, unk_voxels, args, chosen_segment_mask, tumor_volumes_report, tumor_diameters,
                                classes, class_weights) -> Dict[str, torch.Tensor]:
    from .losses import get_known_voxels, seg_loss
    if chosen_segment_mask is not None and unk_voxels is not None:
        csm_any = ops.u8_row_count(_u8(chosen_segment_mask).view(label.shape[0], -1), label[0].numel()).cpu().numpy()
        unk_any = ops.u8_row_count(_u8(unk_voxels).view(label.shape[0], -1), label[0].numel()).cpu().numpy()
        vol_any = tumor_volumes_report.detach().float().sum(-1).cpu().numpy()
        for b in range(label.shape[0]):   # sanity checks of :864-869
            if csm_any[b] > 0 and (unk_any[b] == 0 or vol_any[b] == 0):
                raise ValueError("report sample without unk_voxels / tumor volumes")
    from .losses import _drop_unit_class_weights
    class_weights = _drop_unit_class_weights(class_weights)
    cw_seg = None if class_weights is None else class_weights.to(label.device)
    label_u8 = _u8(label)
    known = get_known_voxels(unk_voxels) if unk_voxels is not None else None
    loss_seg_total = 0
    loss_report: Dict[str, torch.Tensor] = {}
    for j, r in enumerate(heads):
        aw = args.aux_weight[j] if deep else 1.0
        use_ball = any(k in args.loss for k in ("ball", "dynamic", "dll")) and not (deep and j != 0 and "last" in args.loss)
        if use_ball:
            lr = ball_loss(r, label_u8, unk_voxels, chosen_segment_mask, tumor_volumes_report, tumor_diameters, classes,
                           apply_dice_loss=("dice" in args.loss), standard_ce=args.stardard_ce_ball, class_weights=cw_seg,
                           diameter_margin=args.ball_volume_margin, volume_margin=args.ball_volume_margin)
            if "both" in args.loss:
                lr.update(volume_loss_basic(r, chosen_segment_mask, tumor_volumes_report, label_u8, unk_voxels, classes,
                                            class_weights=cw_seg, tolerance=args.volume_loss_tolerance))
        else:
            lr = volume_loss_basic(r, chosen_segment_mask, tumor_volumes_report, label_u8, unk_voxels, classes,
                                   class_weights=cw_seg, tolerance=args.volume_loss_tolerance)
        loss_seg_total = loss_seg_total + aw * args.seg_loss * seg_loss(r, label_u8, known, cw_seg)
        for k, val in lr.items():
            wgt = args.ball_bce_weight if k == "ball_loss_bce" else (args.ball_dice_weight if k == "ball_loss_dice" else 1)
            term = aw * args.report_volume_loss_basic * wgt * val
            loss_report[k] = loss_report[k] + term if k in loss_report else term
    loss = {"segmentation": loss_seg_total}
    loss.update(loss_report)
    overall = 0
    for val in loss.values():
        overall = overall + val
    loss["overall"] = overall
    if getattr(args, "nan_check", True) and torch.isnan(overall).any():   # :1070-1071
        raise ValueError("loss is nan, propagating this can destroy the network weights, STOP!")
    assert overall.requires_grad, "Loss overall should require grad"
    return loss
