"""Golden vectors for `UNet(block='Bottleneck')` (model/dim3/conv_layers.py:97-123, selectable by the yaml `block:` key —
model/dim3/utils.py:7-13) from the REAL reference module: logits (subsampled), loss and every parameter's gradient norm.

Run in the build container only:  python tests/golden/make_golden_bneck.py  ->  tests/golden/reference_bottleneck.npz
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
from oracle import losses_ref as LR  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.unet_ref import synthetic_image, synthetic_state_dict  # noqa: E402

BASE, C, S = 16, 2, 32


def main():
    os.chdir(tempfile.mkdtemp())
    unet_mod, lf = import_reference()[:2]
    net = unet_mod.UNet(1, BASE, num_classes=C, scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5, block="Bottleneck")
    sd = synthetic_state_dict(BASE, C, block="Bottleneck")
    assert [k for k, _ in net.named_parameters()] == list(sd.keys()), "state-dict contract: same names in the same order"
    net.load_state_dict(sd, strict=True)
    x = synthetic_image(1, S, S, S, seed=3)
    logits = net(x)
    classes = ["organ", "pancreatic_lesion"]
    batch = synth.make_batch(["mask"], classes, (S, S, S), seed=5)
    args = LR.default_args(report_volume_loss_basic=0.0)
    loss = lf.calculate_loss(model_output={"segmentation": logits}, label=batch["label"].long(), unk_voxels=None, args=args, matcher=None,
                             chosen_segment_mask=None, tumor_volumes_report=None, tumor_diameters=None, classes=classes)
    loss["overall"].backward()
    out = {"logits": logits.detach().numpy()[:, :, ::2, ::2, ::2].copy(), "loss": np.float32(loss["overall"].item()),
           "grad_norms": np.array([p.grad.norm().item() for _, p in net.named_parameters()], dtype=np.float64),
           "names": np.array([k for k, _ in net.named_parameters()])}
    np.savez_compressed(os.path.join(HERE, "reference_bottleneck.npz"), **out)
    print(out["logits"].shape, out["loss"], len(out["grad_norms"]))


if __name__ == "__main__":
    main()
