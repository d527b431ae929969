"""Golden vectors for the online intensity augmentations from the REAL reference functions
(rsuper_train/training/augmentation.py), each run after torch.manual_seed(s) so that the oracle / the GPU mirror can
re-make the same draws, and for the whole gate block of the loader (dataset_abdomenatlas_UFO.py:1048-1061, copied call by
call below because the loader class itself needs the dataset on disk).

Run in the build container only:  python tests/golden/make_golden_aug.py  ->  tests/golden/reference_augment.npz
"""
import importlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402
from oracle.unet_ref import synthetic_image  # noqa: E402

SHAPE = (20, 24, 28)
SEEDS = {"multiply": 11, "additive": 12, "gamma": 13, "contrast": 14, "blur": 15, "noise": 16}


def sub(t):
    return t.detach().numpy()[0, 0, ::2, ::3, ::2].copy()


def main():
    os.chdir(tempfile.mkdtemp())
    import_reference()
    aug = importlib.import_module("training.augmentation")
    x = synthetic_image(1, *SHAPE, seed=21)                   # [1, 1, D, H, W]
    out = {}
    calls = {
        "multiply": lambda t: aug.brightness_multiply(t, multiply_range=[0.7, 1.3]),
        "additive": lambda t: aug.brightness_additive(t, std=0.1),
        "gamma": lambda t: aug.gamma(t, gamma_range=[0.7, 1.5]),
        "contrast": lambda t: aug.contrast(t, contrast_range=[0.7, 1.3]),
        "blur": lambda t: aug.gaussian_blur(t, sigma_range=[0.5, 1.5]),
        "noise": lambda t: aug.gaussian_noise(t, std=0.13),
    }
    for name, fn in calls.items():
        torch.manual_seed(SEEDS[name])
        y = fn(x.clone())
        out[f"{name}"] = sub(y)
        out[f"{name}_sum"] = np.float64(y.double().sum().item())
    # the whole gate block, all six gates open, then a seeded pass with the real 0.3 gates
    for tag, forced in (("all", True), ("gated", False)):
        np.random.seed(5)
        torch.manual_seed(6)
        t = x.clone()
        order = []
        for name in ("multiply", "additive", "gamma", "contrast", "blur"):
            if (np.random.random() < 0.3) or forced:
                t = calls[name](t)
                order.append(name)
        if (np.random.random() < 0.3) or forced:
            std = np.random.random() * 0.2
            t = aug.gaussian_noise(t, std=std)
            order.append("noise")
        out[f"seq_{tag}"] = sub(t)
        out[f"seq_{tag}_sum"] = np.float64(t.double().sum().item())
        out[f"seq_{tag}_applied"] = np.array([n in order for n in calls], dtype=np.bool_)
    np.savez_compressed(os.path.join(HERE, "reference_augment.npz"), **out)
    print(f"wrote {len(out)} arrays; gated pass applied {order}")


if __name__ == "__main__":
    main()
