"""Synthetic batches with the reference data loader's batch contract — what `bench.py`, `smoke()`
and the tests feed the train step with (there is no network for the datasets): image, one-hot multi-label `label`,
`unk_channels`, `mask` (chosen_segment_mask), `volumes [B,10]`, `diameters [B,10,3]`.
  reference: training/dataset/dataset_abdomenatlas_UFO.py:551-557,1112-1117; train_ddp.py:246-275
Seeded and RNG-library independent: a tiny LCG drives the handful of geometric parameters, so the tensors are identical on
every torch version and device.  Pure tensor construction — no kernels, usable on the CPU.  `default_loss_args()` = the fields
`calculate_loss` reads with `train_ddp.py`'s argparse defaults.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch


class _LCG:
    def __init__(self, seed: int):
        self.s = (seed * 2654435761 + 12345) & 0xFFFFFFFF

    def next(self) -> float:
        self.s = (1664525 * self.s + 1013904223) & 0xFFFFFFFF
        return self.s / 4294967296.0

    def uniform(self, lo, hi):
        return lo + (hi - lo) * self.next()

    def randint(self, lo, hi):  # inclusive
        return lo + int(self.next() * (hi - lo + 1)) % (hi - lo + 1)


def _ellipsoid(shape, center, radii) -> torch.Tensor:
    d, h, w = shape
    zz = torch.arange(d, dtype=torch.float32)[:, None, None]
    yy = torch.arange(h, dtype=torch.float32)[None, :, None]
    xx = torch.arange(w, dtype=torch.float32)[None, None, :]
    v = ((zz - center[0]) / radii[0]) ** 2 + ((yy - center[1]) / radii[1]) ** 2 + ((xx - center[2]) / radii[2]) ** 2
    return (v <= 1.0).to(torch.uint8)


def lesion_channel_indices(classes: Sequence[str]) -> List[int]:
    return [i for i, c in enumerate(classes) if any(s in c for s in ("lesion", "cyst", "pdac", "pnet"))]


def make_sample(kind: str, classes: Sequence[str], shape, seed: int) -> Dict[str, torch.Tensor]:
    """One sample.  kind = 'mask' (per-voxel labels only) or 'report' (lesion known only from a report)."""
    d, h, w = shape
    C = len(classes)
    rng = _LCG(seed)
    label = torch.zeros((C, d, h, w), dtype=torch.uint8)
    unk = torch.zeros_like(label)
    mask = torch.zeros_like(label)
    volumes = torch.zeros(10, dtype=torch.float32)
    diameters = torch.zeros(10, 3, dtype=torch.float32)
    lesions = lesion_channel_indices(classes)
    organs = [i for i in range(C) if i not in lesions]
    m = min(shape)
    # organ ellipsoid shared by all organ channels' geometry seeds
    org_c = [rng.uniform(0.4, 0.6) * s for s in shape]
    org_r = [rng.uniform(0.28, 0.36) * s for s in shape]
    organ = _ellipsoid(shape, org_c, org_r)
    for i in organs:
        cc = [c + rng.uniform(-0.05, 0.05) * m for c in org_c]
        label[i] = _ellipsoid(shape, cc, org_r)
    if kind == "mask":
        for i in lesions:
            if rng.next() < 0.7:
                r = max(1.5, rng.uniform(0.05, 0.12) * m)
                cc = [c + rng.uniform(-0.4, 0.4) * rr for c, rr in zip(org_c, org_r)]
                label[i] = _ellipsoid(shape, cc, [r, r, r]) & organ
    elif kind == "report":
        if not lesions:
            raise ValueError("report samples need a lesion channel")
        lc = lesions[rng.randint(0, len(lesions) - 1)]
        unk[lc] = organ
        mask[lc] = organ
        n_t = rng.randint(1, 3)
        d_hi = max(4, min(30, int(0.28 * m)))
        d_lo = max(3, min(6, d_hi - 1))
        for t in range(n_t):
            dia = float(rng.randint(d_lo, d_hi))
            diameters[t] = dia
            volumes[t] = 4.0 / 3.0 * math.pi * (dia / 2.0) ** 3
    else:
        raise ValueError(kind)
    return dict(label=label, unk_channels=unk, mask=mask, volumes=volumes, diameters=diameters)


def make_batch(kinds: Sequence[str], classes: Sequence[str], shape, seed: int = 1234, device="cpu"):
    """Batch dict in the reference's key names; tensors are uint8 masks / fp32 report targets."""

    samples = [make_sample(k, classes, shape, seed * 131 + i) for i, k in enumerate(kinds)]
    out = {k: torch.stack([s[k] for s in samples]).to(device) for k in samples[0]}
    out["image"] = synthetic_image(len(kinds), *shape, seed=seed, device=device)
    return out


def synthetic_logits(B: int, C: int, shape, seed: int = 0, scale: float = 3.0, device="cpu") -> torch.Tensor:
    """Smooth deterministic pseudo-logits (used to test the loss kernels without a network)."""
    d, h, w = shape
    zz, yy, xx = torch.meshgrid(torch.arange(d, dtype=torch.float64), torch.arange(h, dtype=torch.float64),
                                torch.arange(w, dtype=torch.float64), indexing="ij")
    out = torch.empty((B, C, d, h, w), dtype=torch.float32)
    for b in range(B):
        for c in range(C):
            k = seed * 17 + b * 5 + c * 3 + 1
            cz, cy, cx = d * (0.35 + 0.3 * math.sin(k)), h * (0.5 + 0.2 * math.cos(2 * k)), w * (0.5 + 0.2 * math.sin(3 * k))
            r2 = (zz - cz) ** 2 + (yy - cy) ** 2 + (xx - cx) ** 2
            bump = torch.exp(-r2 / (2 * (0.18 * min(shape)) ** 2))
            v = scale * (2.2 * bump - 1.0) + 0.6 * torch.sin(0.9 * zz + 0.7 * yy * (1 + 0.1 * c) + 1.1 * xx + k) \
                + 0.3 * torch.sin(7.13 * zz + 3.71 * yy + 5.37 * xx + 0.5 * k)
            out[b, c] = v.to(torch.float32)
    return out.to(device)


# ------------------------------------------------------------------------------------------------
# On-disk crop format of the reference loader (SURVEY §8f N2 groundwork): masks are stored bit-packed along the
# CHANNEL axis — np.packbits(bool[C, D, H, W], axis=0) -> uint8[ceil(C / 8), D, H, W]
# (dataset_abdomenatlas_UFO.py:952-975) and unpacked + truncated to C on load (:1006-1015, 1071-1090).
# numpy packs big-endian: channel c lives in byte c // 8, bit 7 - (c % 8).  A GPU batch-assembly kernel can therefore
# read label[c][v] = (packed[c >> 3][v] >> (7 - (c & 7))) & 1 straight from the packed bytes (8x less H2D than uint8,
# 64x less than the int64 labels train_ddp.py uploads today).
# ------------------------------------------------------------------------------------------------
def pack_masks(mask: torch.Tensor):
    """[C, D, H, W] 0/1 -> numpy uint8 [ceil(C/8), D, H, W], the reference's storage format."""
    import numpy as np
    return np.packbits(mask.cpu().numpy().astype(np.bool_), axis=0)


def unpack_masks(packed, num_classes: int) -> torch.Tensor:
    """Inverse of pack_masks as the reference loads it: unpack along axis 0, keep the first num_classes channels."""
    import numpy as np
    full = np.unpackbits(packed, axis=0)
    assert num_classes <= full.shape[0] < num_classes + 10          # the reference's own sanity asserts (:1010-1011)
    return torch.from_numpy(full[:num_classes].copy())


def packed_bit(packed, c: int):
    """Channel c of a packed mask without unpacking (what a kernel would compute per voxel)."""
    return (packed[c >> 3] >> (7 - (c & 7))) & 1


def synthetic_image(n: int, d: int, h: int, w: int, seed: int = 0, device="cpu") -> torch.Tensor:
    """Deterministic CT-like patch in [-3, 3] (N(0,1)-scale, |x| <= 100)."""
    zz, yy, xx = torch.meshgrid(torch.arange(d, dtype=torch.float64), torch.arange(h, dtype=torch.float64),
                                torch.arange(w, dtype=torch.float64), indexing="ij")
    out = []
    for i in range(n):
        s = float(seed * 7 + i)
        v = (torch.sin(0.37 * zz + 0.11 * s) + torch.cos(0.23 * yy * (1 + 0.01 * s)) * torch.sin(0.31 * xx + s)
             + 0.5 * torch.sin(0.05 * (zz * yy + xx) + 0.7 * s) + 0.8 * torch.sin(12.9898 * zz + 78.233 * yy + 37.719 * xx + s))
        out.append(v)
    img = torch.stack(out).unsqueeze(1).to(torch.float32)
    return img.clamp_(-3, 3).to(device)


def default_loss_args(**kw):
    """The args fields calculate_loss reads (losses_foundation.py:726-731) with train_ddp.py's defaults (:429-465) and the
    aux_weight of config/abdomenatlas_ufo/medformer_3d.yaml:42."""
    from types import SimpleNamespace
    a = dict(loss="ball_dice_last", aux_weight=[0.5, 0.5], seg_loss=1.0, report_volume_loss_basic=0.1, volume_loss_tolerance=0.2,
             ball_bce_weight=1.0, ball_dice_weight=1.0, ball_volume_margin=0.2, multi_ch_tumor=False, stardard_ce_ball=False,
             classification_branch=False)
    a.update(kw)
    return SimpleNamespace(**a)
