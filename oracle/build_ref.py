"""ORACLE (test infrastructure, never shipped, never imported by the product path).

Recipe that makes the UNMODIFIED reference implementation of the hot path available as `oracle/_ref/` (git-ignored, NOT
gpurun-ignored: like a compiled reference it travels to the GPU box, where /root/reference does not exist), and the import
shim that loads it.  The reference is pure Python, so "building" it = staging the module files the hot path imports, byte
for byte, from where they lie under /root/reference (read-only) — nothing of them is committed; `MANIFEST.json` records
their sha256.

    python oracle/build_ref.py          # in the build container (also run by __graft_entry__.build())

Used by: bench.py --impl reference / cpu_baseline (kind "reference") and tests/golden/make_golden*.py.
Files staged:
    model/dim3/unet.py, unet_utils.py, conv_layers.py, utils.py, trans_layers.py     UNet and its blocks
    training/losses_foundation.py, info_nce.py                                        calculate_loss, Volume / Ball loss
    training/utils.py                                                                 get_optimizer, update_ema_variables
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/rsuper_train"
DST = os.path.join(HERE, "_ref", "rsuper_train")
FILES = [
    "model/dim3/unet.py", "model/dim3/unet_utils.py", "model/dim3/conv_layers.py", "model/dim3/utils.py",
    "model/dim3/trans_layers.py",
    "training/losses_foundation.py", "training/info_nce.py", "training/utils.py",
]


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def build(verbose: bool = True) -> bool:
    """Stage the files when /root/reference is present (build container); on the GPU box the staged copy is used as is."""
    if not os.path.isdir(SRC):
        return available()
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, "rb") as fh:
            data = fh.read()
        manifest[f] = hashlib.sha256(data).hexdigest()
        if not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"[oracle/_ref] staged {len(FILES)} reference files from {SRC}")
    return True


def import_reference(root: str | None = None):
    """Import the reference's own modules -> (model.dim3.unet, training.losses_foundation, training.utils).
    SURVEY §8c recipe: stub nibabel / matplotlib / SimpleITK (debug writers only), register bare `model` / `model.dim3`
    packages so that model/dim3/__init__.py (needs monai / timm / mmcv) is bypassed."""
    root = root or DST
    if not os.path.exists(os.path.join(root, FILES[0])):
        raise RuntimeError(f"{root} is not staged: run `python oracle/build_ref.py` in the build container")
    for name in ("nibabel", "matplotlib", "matplotlib.pyplot", "SimpleITK"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            if name == "nibabel":
                m.Nifti1Image = lambda *a, **k: None
                m.save = lambda *a, **k: None
            sys.modules[name] = m
    if root not in sys.path:
        sys.path.insert(0, root)
    for pkg, sub in (("model", "model"), ("model.dim3", "model/dim3"), ("training", "training")):
        if pkg not in sys.modules or getattr(sys.modules[pkg], "__path__", [None])[0] != os.path.join(root, sub):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(root, sub)]
            sys.modules[pkg] = m
    unet = importlib.import_module("model.dim3.unet")
    lf = importlib.import_module("training.losses_foundation")
    tu = importlib.import_module("training.utils")
    return unet, lf, tu


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref available:", ok)
