"""Sliding-window inference and post-processing on the B200 kernels.

Host-side mirror of
  * `inference_sliding_window(net, img, args, pancreas=None)`   rsuper_train/inference/inference3d.py:28-107
  * `split_idx`                                                 rsuper_train/inference/utils.py:27-44
  * `postprocess_npz` (organ gating of lesion channels)         rsuper_train/predict_abdomenatlas.py:636-684
  * `keep_largest_component`                                    rsuper_train/predict_abdomenatlas.py:686-710
with the same argument meaning.  The reference blends on the CPU (`pred.cpu()` per window, two host-side `+=` over the
whole window); here sigmoid + blend + count are one kernel per window into device accumulators and the final
`/ counter` (+ optional `> 0.5`) is one more pass, so nothing leaves the GPU until the blended volume is complete.
Connected components run on the device too (`rsb_cc_label`, lock-free union-find).  No CPU / PyTorch fallback.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops


def split_idx(half_win: int, size: int, i: int) -> Tuple[int, int]:
    """inference/utils.py:27-44 — stride half a window; the last window is pulled back inside the volume."""
    start = half_win * i
    end = start + 2 * half_win
    if end > size:
        start, end = size - 2 * half_win, size
    return start, end


def _unwrap(pred):
    if isinstance(pred, dict):
        pred = pred["segmentation"]          # inference3d.py:84-85
    for _ in range(2):                       # :86-89 (deep-supervision list, then tuple)
        if isinstance(pred, (tuple, list)):
            pred = pred[0]
    return pred


def inference_sliding_window(net, img: torch.Tensor, args, pancreas: Optional[torch.Tensor] = None, keep_on_device: bool = False,
                             threshold: Optional[float] = None):
    """img [B, C, D, H, W] (CUDA) -> mean-blended sigmoid probabilities [B, args.classes, D, H, W].

    Returns a CPU tensor like the reference unless keep_on_device=True.  With `threshold` set, returns
    (prob, mask uint8 = prob > threshold) from the same finalize pass."""
    net.eval()
    if not ops._on_device(img):
        raise RuntimeError("rsuper_b200.inference has no CPU path: img must live on a CUDA (sm_100a) device")
    if pancreas is not None:
        while pancreas.dim() < img.dim():
            pancreas = pancreas.unsqueeze(0)
        assert pancreas.shape == img.shape, f"Pancreas mask shape must match image shape, got {pancreas.shape} and {img.shape}"
    B, C, D, H, W = img.shape
    win_d, win_h, win_w = args.window_size
    padded = D < win_d or H < win_h or W < win_w
    if padded:
        origin = (D, H, W)
        pad = (0, max(0, win_w - W), 0, max(0, win_h - H), 0, max(0, win_d - D))
        img = F.pad(img, pad)
        B, C, D, H, W = img.shape
    half = (win_d // 2, win_h // 2, win_w // 2)
    out = torch.zeros((B, args.classes, D, H, W), dtype=torch.float32, device=img.device)
    count = torch.zeros((B, 1, D, H, W), dtype=torch.float32, device=img.device)
    with torch.no_grad():
        for i in range(D // half[0]):
            d0, d1 = split_idx(half[0], D, i)
            for j in range(H // half[1]):
                h0, h1 = split_idx(half[1], H, j)
                for k in range(W // half[2]):
                    w0, w1 = split_idx(half[2], W, k)
                    # NOTE the reference slices the UNPADDED pancreas mask with padded-volume indices (inference3d.py:82);
                    # slicing clips at the mask's extent, which is what we reproduce here
                    if pancreas is None or bool(pancreas[:, :, d0:d1, h0:h1, w0:w1].sum() > 0):
                        pred = _unwrap(net(img[:, :, d0:d1, h0:h1, w0:w1].contiguous()))
                        pred = pred.float().contiguous()
                    else:
                        pred = None    # zeros are added, the window still counts (:92-100)
                    ops.sigmoid_window_accumulate(pred, out, count, (d0, h0, w0), (d1 - d0, h1 - h0, w1 - w0))
    prob, mask = ops.blend_finalize(out, count, threshold=threshold)
    if padded:
        prob = prob[:, :, :origin[0], :origin[1], :origin[2]]
        mask = mask[:, :, :origin[0], :origin[1], :origin[2]] if mask is not None else None
    if not keep_on_device:
        prob = prob.cpu()
        mask = mask.cpu() if mask is not None else None
    return prob if threshold is None else (prob, mask)


# organ of a lesion class (predict_abdomenatlas.py:657-671)
_PAIRED = {"kidney": ("kidney_right", "kidney_left"), "adrenal": ("adrenal_gland_right", "adrenal_gland_left"),
           "lung": ("lung_right", "lung_left")}
_RENAMED = {"uterus": "prostate", "gallbladder": "gall_bladder"}


def postprocess_npz(pred: torch.Tensor, classes: Sequence[str], args) -> Dict[str, torch.Tensor]:
    """pred [1, C, D, H, W] probabilities (CUDA) -> {class: [D, H, W]} with every lesion channel multiplied by its
    organ's (> 0.5, 3x3x3-dilated) mask when args.organ_mask_on_lesion (predict_abdomenatlas.py:636-684)."""
    if not ops._on_device(pred):
        raise RuntimeError("rsuper_b200.inference has no CPU path")
    pred = pred.squeeze(0).float().contiguous()
    out: Dict[str, torch.Tensor] = {}
    for i, name in enumerate(classes):
        if "lesion" not in name:
            out[name] = pred[i]
    for i, name in enumerate(classes):
        if "lesion" not in name:
            continue
        p = pred[i].clone()
        if getattr(args, "organ_mask_on_lesion", False):
            organ_name = name.split("_")[0].replace("pancreatic", "pancreas")
            if organ_name in _PAIRED:
                a, b = _PAIRED[organ_name]
                organ = out[a] + out[b]
            elif organ_name in ("bone", "breast"):
                organ = torch.ones_like(p)
            else:
                organ = out[_RENAMED.get(organ_name, organ_name)]     # KeyError for an unknown organ, like the reference
            mask = (organ > 0.5).to(torch.uint8).contiguous()
            ops.gate_by_mask(p, ops.dilate_box3(mask))
        out[name] = p
    return out


def connected_components(mask: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Face-connected labelling of one [D, H, W] volume (`mask > 0`).  Returns (labels int32, -1 = background, otherwise the
    smallest linear index of the component; count)."""
    m = (mask > 0).to(torch.uint8).contiguous()
    labels, n, _ = ops.cc_label(m)
    return labels, int(n.item())


def keep_largest_component(label_map: torch.Tensor) -> torch.Tensor:
    """predict_abdomenatlas.py:686-710 — uint8 mask of the first (raster order) component of maximal size."""
    m = (label_map > 0).to(torch.uint8).contiguous()
    _, _, largest = ops.cc_label(m, keep_largest=True)
    return largest
