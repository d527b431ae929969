"""Online intensity augmentations on the device (second half).

Host-side mirror of `training/augmentation.py` (`gaussian_noise` :17-19, `gaussian_blur` :48-66, `brightness_additive`
:69-83, `brightness_multiply` :86-103, `gamma` :106-138, `contrast` :140-168) and of the block of the loader that applies
them (`training/dataset/dim3/dataset_abdomenatlas_UFO.py:1048-1061`): same function names, same argument meaning, and the
random numbers are drawn from the SAME global generators, in the same order and shapes as the reference draws them
(`np.random.random()` for the 0.3 gates and the noise level, `torch.rand` / `torch.normal` / `torch.randn` on the CPU
generator for the parameters), so a seeded run reproduces the reference's augmentation stream.  The arithmetic on the
volume runs in csrc/augment.cu.  `device_noise=True` draws the Gaussian noise field on the GPU instead (same distribution,
different stream; avoids a volume-sized H2D copy).  Per-channel variants are not implemented (the loader never uses them;
R-Super images have one channel).  No CPU / PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


def _check(tensor_img: torch.Tensor) -> torch.Tensor:
    if not ops._on_device(tensor_img):
        raise RuntimeError("rsuper_b200.augment has no CPU path: the image must live on a CUDA (sm_100a) device")
    if tensor_img.dim() != 5:
        raise ValueError("Invalid input tensor dimension, should be 5d for volume image")   # message of augmentation.py:66
    if tensor_img.shape[1] != 1:
        raise NotImplementedError("rsuper_b200.augment: single-channel images only (per_channel variants are not implemented)")
    return tensor_img.float().contiguous()


def gaussian_kernel_1d(sigma: float) -> Sequence[float]:
    """1-D marginal of generate_3d_gaussian_kernel (augmentation.py:34-46): the normalised 3-D kernel is its outer product."""
    kernel_size = 2 * math.ceil(3 * sigma) + 1
    x = torch.arange(-kernel_size // 2 + 1, kernel_size // 2 + 1, dtype=torch.float32)
    k = torch.exp(-(x ** 2) / (2 * sigma ** 2)).double()
    return (k / k.sum()).tolist()


def brightness_multiply(tensor_img, multiply_range=(0.7, 1.3), factor: Optional[float] = None):
    x = _check(tensor_img)
    assert multiply_range[1] > multiply_range[0], "Invalid range"
    if factor is None:
        factor = (torch.rand(size=(1, 1, 1, 1, 1)) * (multiply_range[1] - multiply_range[0]) + multiply_range[0]).item()
    return ops.aug_affine(x, mul=factor)


def brightness_additive(tensor_img, std, mean=0, offset: Optional[float] = None):
    x = _check(tensor_img)
    if offset is None:
        offset = torch.normal(mean, std, size=(1, 1, 1, 1, 1)).item()
    return ops.aug_affine(x, add=offset)


def gamma(tensor_img, gamma_range=(0.5, 2), retain_stats=True, gamma: Optional[float] = None):
    x = _check(tensor_img)
    if gamma is None:
        gamma = (torch.rand(1, 1) * (gamma_range[1] - gamma_range[0]) + gamma_range[0]).item()
    return ops.aug_gamma(x, gamma, retain_stats=retain_stats)


def contrast(tensor_img, contrast_range=(0.65, 1.5), preserve_range=True, factor: Optional[float] = None):
    x = _check(tensor_img)
    if not preserve_range:
        raise NotImplementedError("rsuper_b200.augment.contrast: preserve_range=False is not implemented (the loader uses the default)")
    if factor is None:
        factor = (torch.rand(1, 1) * (contrast_range[1] - contrast_range[0]) + contrast_range[0]).item()
    return ops.aug_contrast(x, factor)


def gaussian_blur(tensor_img, sigma_range=(0.5, 1.0), sigma: Optional[float] = None):
    x = _check(tensor_img)
    if sigma is None:
        sigma = (torch.rand(1) * (sigma_range[1] - sigma_range[0]) + sigma_range[0]).item()
    return ops.aug_blur(x, gaussian_kernel_1d(sigma))


def gaussian_noise(tensor_img, std, mean=0, noise: Optional[torch.Tensor] = None, device_noise: bool = False):
    x = _check(tensor_img)
    if noise is None:
        noise = torch.randn(x.shape, device=x.device) if device_noise else torch.randn(x.shape).to(x.device)
    y = ops.aug_affine(x, noise=noise.float().contiguous(), noise_std=std)
    return y if mean == 0 else ops.aug_affine(y, add=mean)


def online_intensity_augmentation(tensor_img: torch.Tensor, device_noise: bool = False) -> torch.Tensor:
    """dataset_abdomenatlas_UFO.py:1048-1061 (mode == 'train'): six gates of probability 0.3, in the reference's order."""
    if np.random.random() < 0.3:
        tensor_img = brightness_multiply(tensor_img, multiply_range=[0.7, 1.3])
    if np.random.random() < 0.3:
        tensor_img = brightness_additive(tensor_img, std=0.1)
    if np.random.random() < 0.3:
        tensor_img = gamma(tensor_img, gamma_range=[0.7, 1.5])
    if np.random.random() < 0.3:
        tensor_img = contrast(tensor_img, contrast_range=[0.7, 1.3])
    if np.random.random() < 0.3:
        tensor_img = gaussian_blur(tensor_img, sigma_range=[0.5, 1.5])
    if np.random.random() < 0.3:
        std = np.random.random() * 0.2
        tensor_img = gaussian_noise(tensor_img, std=std, device_noise=device_noise)
    return tensor_img
