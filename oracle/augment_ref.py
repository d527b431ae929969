"""ORACLE (test infrastructure, never shipped) — SURVEY §8(f) N2: the online intensity augmentations of the loader.

  brightness_multiply / brightness_additive / gamma / contrast / gaussian_blur / gaussian_noise
                                  rsuper_train/training/augmentation.py:17-168
  the 0.3-probability gate block  rsuper_train/training/dataset/dim3/dataset_abdomenatlas_UFO.py:1048-1061

restated with the random draws as explicit arguments (plain torch, CPU or GPU tensors, [1, 1, D, H, W] fp32), plus
`draws_like_reference`, which makes the draws from the global generators in the order and shapes the reference uses.
Pinned against the REAL functions by tests/golden/make_golden_aug.py -> tests/golden/reference_augment.npz.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def brightness_multiply(x, factor):
    return x * factor


def brightness_additive(x, offset):
    return x + offset


def gamma(x, g, retain_stats=True):
    shape = x.shape
    t = x.reshape(1, -1)
    minm, maxm = t.min(dim=1)[0].unsqueeze(1), t.max(dim=1)[0].unsqueeze(1)
    rng = maxm - minm
    mean, std = t.mean(dim=1).unsqueeze(1), t.std(dim=1).unsqueeze(1)
    t = torch.pow((t - minm) / rng, g) * rng + minm
    if retain_stats:
        t = t - t.mean(dim=1).unsqueeze(1)
        t = t / t.std(dim=1).unsqueeze(1) * std + mean
    return t.reshape(shape)


def contrast(x, factor):
    shape = x.shape
    t = x.reshape(1, -1)
    minm, maxm = t.min(dim=1)[0].unsqueeze(1), t.max(dim=1)[0].unsqueeze(1)
    mean = t.mean(dim=1).unsqueeze(1)
    t = (t - mean) * factor + mean
    return torch.clamp(t, min=minm, max=maxm).reshape(shape)


def gaussian_kernel_3d(sigma: float) -> torch.Tensor:
    k = 2 * math.ceil(3 * sigma) + 1
    r = torch.arange(-k // 2 + 1, k // 2 + 1, dtype=torch.float32)
    x, y, z = torch.meshgrid(r, r, r, indexing="ij")
    ker = torch.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * sigma ** 2))
    ker = ker / (2 * math.pi * sigma ** 2) ** 1.5
    return (ker / ker.sum()).unsqueeze(0).unsqueeze(0)


def gaussian_blur(x, sigma):
    ker = gaussian_kernel_3d(sigma).to(x.device)
    return F.conv3d(x, ker, padding=ker.shape[-1] // 2)


def gaussian_noise(x, noise, std):
    return x + noise * std + 0


def draws_like_reference(shape, gates=None):
    """The random numbers of one pass through dataset_abdomenatlas_UFO.py:1048-1061, drawn from np.random / torch's CPU
    generator in the reference's order.  gates: optional 6 booleans overriding the np.random gates (tests)."""
    d = {}
    names = ("multiply", "additive", "gamma", "contrast", "blur", "noise")
    for i, name in enumerate(names):
        r = np.random.random()                       # the gate draw is consumed either way (keeps the streams aligned)
        on = (r < 0.3) if gates is None else bool(gates[i])
        if not on:
            continue
        if name == "multiply":
            d[name] = (torch.rand(size=(1, 1, 1, 1, 1)) * 0.6 + 0.7).item()
        elif name == "additive":
            d[name] = torch.normal(0, 0.1, size=(1, 1, 1, 1, 1)).item()
        elif name == "gamma":
            d[name] = (torch.rand(1, 1) * (1.5 - 0.7) + 0.7).item()
        elif name == "contrast":
            d[name] = (torch.rand(1, 1) * (1.3 - 0.7) + 0.7).item()
        elif name == "blur":
            d[name] = (torch.rand(1) * (1.5 - 0.5) + 0.5).item()
        else:
            std = np.random.random() * 0.2
            d[name] = (std, torch.randn(shape))
    return d


def apply(x, draws):
    if "multiply" in draws:
        x = brightness_multiply(x, draws["multiply"])
    if "additive" in draws:
        x = brightness_additive(x, draws["additive"])
    if "gamma" in draws:
        x = gamma(x, draws["gamma"])
    if "contrast" in draws:
        x = contrast(x, draws["contrast"])
    if "blur" in draws:
        x = gaussian_blur(x, draws["blur"])
    if "noise" in draws:
        x = gaussian_noise(x, draws["noise"][1].to(x.device), draws["noise"][0])
    return x
