"""GPU batch assembly from the reference's on-disk crop format.

The reference loader (`training/dataset/dim3/dataset_abdomenatlas_UFO.py:995-1117`) reads, per sample,
  * `<id>.npy`                            float image [D, H, W]
  * `<id>_gt.npy`                         label, np.packbits(bool[C, D, H, W], axis=0) -> uint8 [ceil(C/8), D, H, W] (:952-975)
  * `<id>_gt_unk.npy`                     unknown-voxel channels, same packing (:1079-1085)
  * `<id>_gt_chosen_tumor_segment.npy`    chosen-segment mask, same packing (:1095-1100)
  * `<id>.json`                           tumor volumes / diameters
unpacks the masks on the host with np.unpackbits and `train_ddp.py:246-275` uploads them (`label.long()`: 8 bytes per
voxel and class).  Here the packed bytes are uploaded AS STORED (1 bit per voxel and class: 64x fewer bytes than the int64
label, 8x fewer than uint8) from pinned memory and `rsb_unpack_masks` expands them on the device into the uint8
[B, C, D, H, W] tensors the loss kernels read.  Keys and shapes of the returned dict follow the reference's batch
(`dataset_abdomenatlas_UFO.py:1106-1111`); masks stay uint8 (calculate_loss takes them as they are).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import ops


def _stack_pinned(arrays: Sequence[np.ndarray], dtype, device) -> torch.Tensor:
    first = np.asarray(arrays[0])
    host = torch.empty((len(arrays),) + tuple(first.shape), dtype=dtype)
    if torch.device(device).type == "cuda":
        host = host.pin_memory()      # asynchronous H2D needs page-locked staging
    for i, a in enumerate(arrays):
        a = np.asarray(a)
        if a.shape != first.shape:
            raise ValueError(f"assemble_batch: sample {i} has shape {a.shape}, sample 0 has {first.shape}")
        host[i].copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return host


def upload_packed_masks(packed: Sequence[np.ndarray], num_classes: int, device) -> torch.Tensor:
    """List of per-sample packed masks uint8 [ceil(C/8), D, H, W] -> device uint8 [B, C, D, H, W] (0/1)."""
    cp = (num_classes + 7) // 8
    for i, a in enumerate(packed):
        a = np.asarray(a)
        if a.dtype != np.uint8 or a.ndim != 4:
            raise ValueError(f"assemble_batch: packed mask {i} must be uint8 [ceil(C/8), D, H, W], got {a.dtype} {a.shape}")
        # the reference's own sanity asserts on the unpacked channel count (:1010-1011)
        assert a.shape[0] * 8 < num_classes + 10 and a.shape[0] * 8 >= num_classes, \
            f"packed mask {i} has {a.shape[0]} byte planes for {num_classes} classes"
        assert a.shape[0] == cp
    host = _stack_pinned(packed, torch.uint8, device)
    return ops.unpack_masks(host.to(device, non_blocking=True), num_classes)


def assemble_batch(images: Sequence[np.ndarray], labels_packed: Sequence[np.ndarray], num_classes: int, device,
                   unk_packed: Optional[Sequence[Optional[np.ndarray]]] = None,
                   chosen_packed: Optional[Sequence[Optional[np.ndarray]]] = None,
                   volumes: Optional[Sequence[Sequence[float]]] = None,
                   diameters: Optional[Sequence[np.ndarray]] = None) -> Dict[str, torch.Tensor]:
    """One training batch on `device` from per-sample crop files.  A sample without `_gt_unk` / `_chosen_tumor_segment`
    files (fully annotated CT: None entries) gets all-zero masks, volumes and diameters (:1072-1077)."""
    device = torch.device(device)
    if not ops._on_device(device):
        raise RuntimeError("rsuper_b200.batch has no CPU path: device must be a CUDA (sm_100a) device")
    B = len(images)
    if B == 0 or len(labels_packed) != B:
        raise ValueError("assemble_batch: need one packed label per image")
    img = _stack_pinned([np.asarray(a, dtype=np.float32) for a in images], torch.float32, device)
    if img.dim() != 4:
        raise ValueError(f"assemble_batch: images must be [D, H, W], got {tuple(img.shape[1:])}")
    out = {"image": img.to(device, non_blocking=True).unsqueeze(1),
           "label": upload_packed_masks(labels_packed, num_classes, device)}
    cp = (num_classes + 7) // 8
    zero_planes = np.zeros((cp,) + tuple(img.shape[1:]), dtype=np.uint8)

    def masks(seq):
        if seq is None:
            return torch.zeros_like(out["label"])
        return upload_packed_masks([zero_planes if a is None else a for a in seq], num_classes, device)

    out["unk_channels"] = masks(unk_packed)
    out["mask"] = masks(chosen_packed)
    vol = torch.zeros((B, 10), dtype=torch.float32)
    dia = torch.zeros((B, 10, 3), dtype=torch.float32)
    for i in range(B):
        if volumes is not None and volumes[i] is not None:
            vol[i] = torch.as_tensor(np.asarray(volumes[i], dtype=np.float32))
        if diameters is not None and diameters[i] is not None:
            dia[i] = torch.as_tensor(np.asarray(diameters[i], dtype=np.float32))
    out["volumes"] = vol.to(device)
    out["diameters"] = dia.to(device)
    return out
