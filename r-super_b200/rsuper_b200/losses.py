"""calculate_loss — host-side mirror of rsuper_train/training/losses_foundation.py:685-1076 for the
hot path, computing on the librsuper_b200.so kernels (no torch math on [B,C,V] tensors; the segmentation
path has no host sync once `args.nan_check = False` moves the reference's NaN check to the caller).

Same call signature, same returned dict keys ('segmentation', 'report' | 'ball_loss_bce' /
'ball_loss_dice' / 'dice_volume_loss', 'overall'), same error behaviour (ValueError / AssertionError
on inconsistent report batches).  Baseline-only branches (classification head, CLIP, Model Genesis,
Hungarian matching) are out of scope and raise NotImplementedError.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import ops


def _as_u8(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Masks arrive as uint8 / int64 / float 0-1 tensors (train_ddp.py:246-262); kernels take uint8."""
    if t is None:
        return None
    if t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    return t.contiguous()


class _SegLoss(torch.autograd.Function):
    """mean(BCEWithLogits * known) + DiceLossMultiClass — losses_foundation.py:945-956, 541-607."""

    @staticmethod
    def forward(ctx, logits, label_u8, known_u8, class_weights):
        lg = logits.detach().contiguous().float()
        st = ops.seg_loss_forward(lg, label_u8, known_u8, class_weights)
        ctx.state = st
        return st.loss_out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.state
        dl = torch.empty_like(st.keep[0])
        ops.seg_loss_backward(st, grad_out.detach().reshape(1).float().repeat(2).contiguous(), dl)
        ctx.state = None
        return dl, None, None, None


def seg_loss(logits, label, known=None, class_weights=None) -> torch.Tensor:
    cw = None
    if class_weights is not None:
        cw = class_weights.reshape(logits.shape[0], logits.shape[1]).float().contiguous()
    return _SegLoss.apply(logits, _as_u8(label), _as_u8(known), cw)


class _DiceLoss(torch.autograd.Function):
    """The Dice (adaptive Tversky) term alone: loss_out[2] of the fused kernel, backward with the BCE scale set to 0."""

    @staticmethod
    def forward(ctx, logits, label_u8, known_u8, class_weights):
        st = ops.seg_loss_forward(logits.detach().contiguous().float(), label_u8, known_u8, class_weights)
        ctx.state = st
        return st.loss_out[2].clone()

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.state
        dl = torch.empty_like(st.keep[0])
        scale = torch.cat([torch.zeros(1, dtype=torch.float32, device=dl.device), grad_out.detach().reshape(1).float()]).contiguous()
        ops.seg_loss_backward(st, scale, dl)
        ctx.state = None
        return dl, None, None, None


def DiceLossMultiClass(preds, targets, known_voxels, alpha=0.5, beta=0.5, size_average=True, reduce=True, sigmoid=True,
                       class_weights=None) -> torch.Tensor:
    """Public copy-out function of the reference (README.md:119-128; losses_foundation.py:541-607), same signature.
    `alpha` / `beta` are ignored exactly like in the reference (it overwrites them with the batch-adaptive values, :581-586).
    targets / known_voxels are 0/1 masks; logits fp32.  The variants the reference's own call sites never use
    (sigmoid=False, reduce=False, size_average=False) raise."""
    if not (sigmoid and reduce and size_average):
        raise NotImplementedError("rsuper_b200.DiceLossMultiClass implements sigmoid=True, reduce=True, size_average=True")
    while preds.dim() < 5:                                   # :543-553: [D,H,W] -> [1,1,D,H,W], [C,D,H,W] -> [1,C,D,H,W]
        lead = 2 if preds.dim() == 3 else 1
        for _ in range(lead):
            preds, targets, known_voxels = preds.unsqueeze(0), targets.unsqueeze(0), known_voxels.unsqueeze(0)
    assert preds.dim() == 5
    assert preds.shape == targets.shape and targets.shape == known_voxels.shape, \
        f"Shapes do not match, pred, target and unk are: {preds.shape}, {targets.shape}, {known_voxels.shape}"
    cw = None
    if class_weights is not None:
        cw = class_weights.float().mean(dim=(-1, -2, -3))    # :593
        while cw.dim() < 2:
            cw = cw.unsqueeze(0)
        assert tuple(cw.shape) == tuple(preds.shape[:2]), \
            f"Class weights shape {tuple(cw.shape)} does not match the shape of dice loss {tuple(preds.shape[:2])}"
        cw = cw.contiguous()
    return _DiceLoss.apply(preds, _as_u8(targets), _as_u8(known_voxels), cw)


def get_known_voxels(unk_voxels: torch.Tensor, dilation: int = 5) -> torch.Tensor:
    """1 - dilate(unk, 5) as a uint8 mask (losses_foundation.py:150-199)."""
    unk = _as_u8(unk_voxels)
    if dilation > 0:
        unk = ops.dilate_ball(unk, dilation)
    return unk.logical_not().to(torch.uint8)


_ALL_ONES: Dict[tuple, bool] = {}


def _drop_unit_class_weights(class_weights: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """`if torch.equal(class_weights, ones): class_weights = None` (losses_foundation.py:871-873) without a host sync per
    step: the comparison runs once per (storage, version) of the weight tensor and is cached."""
    if class_weights is None:
        return None
    key = (class_weights.data_ptr(), class_weights._version, tuple(class_weights.shape), str(class_weights.device))
    hit = _ALL_ONES.get(key)
    if hit is None:
        if class_weights.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("rsuper_b200.calculate_loss: class_weights must have been seen by one eager step before CUDA-graph "
                               "capture (their all-ones test is a host sync)")
        if len(_ALL_ONES) > 64:
            _ALL_ONES.clear()
        hit = _ALL_ONES[key] = bool(torch.equal(class_weights, torch.ones_like(class_weights)))
    return None if hit else class_weights


def capturable(args, report_batches: bool) -> bool:
    """True when calculate_loss can sit inside a CUDA-graph capture for this configuration: no host-side control flow on
    device values.  The mask-only path qualifies once the NaN check (`args.nan_check`, a host sync mirroring
    losses_foundation.py:1070) is moved to the caller; the report path qualifies when report_losses runs device-side."""
    if getattr(args, "nan_check", True):
        return False
    if report_batches and float(args.report_volume_loss_basic) > 0:
        from . import report_losses
        return bool(getattr(report_losses, "DEVICE_SIDE", False))
    return True


def calculate_loss(model_output, label, unk_voxels, args, matcher, chosen_segment_mask, tumor_volumes_report,
                   tumor_diameters, classes, input_tensor=None, class_weights=None, model_genesis=False,
                   clip_only=False, report_embeddings=None, dist=None) -> Dict[str, torch.Tensor]:
    """Host syncs: none on the mask-only path when `args.nan_check` is False (the reference's NaN check, default on, reads
    the loss back) — that is the configuration `B200TrainStep(schedule='graph')` captures; `class_weights` cost one
    comparison the first time a tensor is seen."""
    dev = model_output["segmentation"][0].device if isinstance(model_output["segmentation"], (tuple, list)) \
        else model_output["segmentation"].device
    if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):          # kernels launch on the current device's stream: follow the tensors
            return calculate_loss(model_output, label, unk_voxels, args, matcher, chosen_segment_mask, tumor_volumes_report,
                                  tumor_diameters, classes, input_tensor, class_weights, model_genesis, clip_only,
                                  report_embeddings, dist)
    if model_genesis or clip_only or getattr(args, "classification_branch", False) or getattr(args, "multi_ch_tumor", False):
        raise NotImplementedError("rsuper_b200.calculate_loss implements the R-Super segmentation/report path only")
    result = model_output["segmentation"]
    heads = list(result) if isinstance(result, (tuple, list)) else [result]
    deep = isinstance(result, (tuple, list))
    assert len(classes) == label.shape[1], \
        f"Number of classes in classes: {len(classes)} does not match the number of channels in label: {label.shape[1]}"
    assert len(classes) == heads[0].shape[1]
    report_w = float(args.report_volume_loss_basic)
    if report_w > 0 and chosen_segment_mask is not None:
        from . import report_losses  # Volume / Ball loss kernels
        return report_losses.calculate_loss_with_reports(heads, deep, label, unk_voxels, args, chosen_segment_mask,
                                                         tumor_volumes_report, tumor_diameters, classes, class_weights)
    class_weights = _drop_unit_class_weights(class_weights)
    label_u8 = _as_u8(label)
    known = get_known_voxels(unk_voxels) if unk_voxels is not None else None
    loss_seg = 0
    for j, r in enumerate(heads):
        aw = args.aux_weight[j] if deep else 1.0
        loss_seg = loss_seg + aw * args.seg_loss * seg_loss(r, label_u8, known, class_weights)
    zero = torch.zeros((), dtype=torch.float32, device=heads[0].device)
    loss = {"segmentation": loss_seg, "report": zero}
    loss["overall"] = loss["segmentation"] + loss["report"]
    if getattr(args, "nan_check", True) and torch.isnan(loss["overall"]).any():  # losses_foundation.py:1070-1071
        raise ValueError("loss is nan, propagating this can destroy the network weights, STOP!")
    assert loss["overall"].requires_grad, "Loss overall should require grad"
    return loss
