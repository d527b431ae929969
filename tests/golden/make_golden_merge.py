"""Golden vectors for lesion groups that MERGE several channels of one organ (get_lesion_channels,
training/losses_foundation.py:204-248: torch.stack(...).max(dim=0)), from the REAL reference:
the merged channels, volume_loss_basic (value + gradient sum) and ball_loss on a class list with a two-channel group
('liver_lesion_1' + 'liver_lesion_2') and a name matching two suffixes ('kidney_cyst_lesion').

Run in the build container only:  python tests/golden/make_golden_merge.py  ->  tests/golden/reference_merge.npz
"""
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
from oracle import synth  # noqa: E402

MERGE_CLASSES = ["liver", "liver_lesion_1", "liver_lesion_2", "pancreatic_cyst", "kidney_cyst_lesion"]
SHAPE = (16, 24, 32)


SEEDS = {"merged": 20, "single": 23}   # report lesion inside the two-channel group / in a single-channel group


def merge_inputs(tag):
    lg = synth.synthetic_logits(2, len(MERGE_CLASSES), SHAPE, seed=4)
    bt = synth.make_batch(["report", "mask"], MERGE_CLASSES, SHAPE, seed=SEEDS[tag])
    return lg, bt


def main():
    os.chdir(tempfile.mkdtemp())
    _, lf = import_reference()[:2]
    out = {}
    for tag in SEEDS:
        lg, bt = merge_inputs(tag)
        merged, names = lf.get_lesion_channels(lg, MERGE_CLASSES, return_class_names=True)
        out["names"] = np.array(names)
        out["merged_logits"] = merged.numpy()[:, :, 1::3, 1::3, 1::3].copy()
        out[f"{tag}_merged_mask"] = np.packbits(lf.get_lesion_channels(bt["mask"].float(), MERGE_CLASSES).numpy().astype(bool).reshape(-1))
        x = lg.clone().requires_grad_(True)
        vl = lf.volume_loss_basic(x, bt["mask"].float(), bt["volumes"], bt["label"].float(), bt["unk_channels"].float(), MERGE_CLASSES,
                                  tolerance=0.2)["dice_volume_loss"]
        vl.backward()
        out[f"{tag}_volume_loss"] = np.float32(vl.item())
        out[f"{tag}_volume_grad_per_channel"] = x.grad.double().abs().sum(dim=(0, 2, 3, 4)).numpy()
        x2 = lg.clone().requires_grad_(True)
        bl = lf.ball_loss(out=x2, labels=bt["label"].float(), unk_voxels=bt["unk_channels"].float(), chosen_segment_mask=bt["mask"].float(),
                          tumor_volumes=bt["volumes"], tumor_diameters=bt["diameters"], classes=MERGE_CLASSES, apply_dice_loss=True,
                          diameter_margin=0.2, volume_margin=0.2)
        (bl["ball_loss_bce"] + bl["ball_loss_dice"]).backward()
        out[f"{tag}_ball_loss_bce"], out[f"{tag}_ball_loss_dice"] = np.float32(bl["ball_loss_bce"].item()), np.float32(bl["ball_loss_dice"].item())
        out[f"{tag}_ball_grad_per_channel"] = x2.grad.double().abs().sum(dim=(0, 2, 3, 4)).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_merge.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.ndim else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
