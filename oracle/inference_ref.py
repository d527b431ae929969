"""ORACLE (test infrastructure, never shipped) — groundwork for SURVEY §8(f) N3: sliding-window inference, thresholding,
organ gating and connected components.

  inference_sliding_window           rsuper_train/inference/inference3d.py:28-107 (half-overlap windows, sigmoid, mean blend)
  split_idx                          rsuper_train/inference/utils.py:27-44
  organ gating of lesion channels    rsuper_train/predict_abdomenatlas.py:637-684 (threshold 0.5, 3x3x3 binary dilation)
  keep_largest_component             predict_abdomenatlas.py:686-710 (sitk.ConnectedComponentImageFilter: face connectivity)

The window logic is pinned against the real reference function by tests/golden/make_golden.py (`sliding_*` keys).  The
connected-component part restates SimpleITK's default (face / 6-connectivity) with scipy.ndimage.label; SimpleITK is not
installed in this image, so that part is "parity unpinned" (checked only against hand-built cases in the tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def split_idx(half_win: int, size: int, i: int) -> Tuple[int, int]:
    """Window i along one axis: stride = half a window, the last window is pulled back inside the volume (utils.py:27-44)."""
    start = half_win * i
    end = start + 2 * half_win
    if end > size:
        start, end = size - 2 * half_win, size
    return start, end


def window_starts(size: int, win: int):
    """All (start, end) pairs the reference visits along one axis: range(size // (win // 2)) windows (inference3d.py:72-78)."""
    half = win // 2
    return [split_idx(half, size, i) for i in range(size // half)]


def inference_sliding_window(net: Callable[[torch.Tensor], torch.Tensor], img: torch.Tensor, window_size: Sequence[int],
                             num_classes: int, gate: Optional[torch.Tensor] = None) -> torch.Tensor:
    """img [B,C,D,H,W] -> mean-blended sigmoid probabilities [B,classes,D,H,W] on the CPU (inference3d.py:28-107).

    Volumes smaller than the window are zero-padded at the far end and cropped back.  gate (the reference's `pancreas`
    argument): windows whose gate sum is zero are skipped and contribute zeros (they still count in the blend)."""
    B, C, D, H, W = img.shape
    wd, wh, ww = window_size
    padded = D < wd or H < wh or W < ww
    if padded:
        oD, oH, oW = D, H, W
        img = F.pad(img, (0, max(0, ww - W), 0, max(0, wh - H), 0, max(0, wd - D)))
        if gate is not None:
            gate = F.pad(gate, (0, max(0, ww - W), 0, max(0, wh - H), 0, max(0, wd - D)))
        B, C, D, H, W = img.shape
    out = torch.zeros((B, num_classes, D, H, W))
    count = torch.zeros((B, 1, D, H, W))
    with torch.no_grad():
        for d0, d1 in window_starts(D, wd):
            for h0, h1 in window_starts(H, wh):
                for w0, w1 in window_starts(W, ww):
                    if gate is None or gate[:, :, d0:d1, h0:h1, w0:w1].sum() > 0:
                        pred = net(img[:, :, d0:d1, h0:h1, w0:w1])
                        if isinstance(pred, dict):
                            pred = pred["segmentation"]
                        while isinstance(pred, (tuple, list)):
                            pred = pred[0]
                        pred = torch.sigmoid(pred).cpu()
                    else:
                        pred = torch.zeros((B, num_classes, wd, wh, ww))
                    out[:, :, d0:d1, h0:h1, w0:w1] += pred
                    count[:, :, d0:d1, h0:h1, w0:w1] += 1.0
    out /= count
    if padded:
        out = out[:, :, :oD, :oH, :oW]
    return out


def gate_lesion_by_organ(lesion_prob: np.ndarray, organ_prob: np.ndarray) -> np.ndarray:
    """predict_abdomenatlas.py:672-682: organ > 0.5, dilated with a full 3x3x3 structuring element, multiplies the lesion map."""
    from scipy import ndimage as ndi
    organ = ndi.binary_dilation((organ_prob > 0.5).astype(np.uint8), structure=np.ones((3, 3, 3)))
    return organ.astype(lesion_prob.dtype) * lesion_prob


def connected_components(mask: np.ndarray) -> Tuple[np.ndarray, int]:
    """Face-connected (6-connectivity) labelling — sitk.ConnectedComponentImageFilter's default, FullyConnected off
    (predict_abdomenatlas.py:690-695).  Returns (labels, count).  Parity unpinned: SimpleITK is absent from this image."""
    from scipy import ndimage as ndi
    labels, n = ndi.label(mask > 0)   # default structure = face connectivity
    return labels, int(n)


def keep_largest_component(mask: np.ndarray) -> np.ndarray:
    """predict_abdomenatlas.py:686-710: the first component of maximal size wins (strict '>' over labels 1..n)."""
    labels, n = connected_components(mask)
    if n == 0:
        return np.zeros_like(mask, dtype=bool)
    sizes = np.bincount(labels.reshape(-1), minlength=n + 1)[1:]
    return labels == (int(np.argmax(sizes)) + 1)
