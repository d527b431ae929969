"""Golden vectors for MedFormer (model/dim3/medformer.py) from the REAL reference module on the WELL-CONDITIONED synthetic state
(oracle.medformer_ref.conditioned_state: the hash init with the semantic-projection weights scaled so that the semantic maps
are not degenerate): logits and deep-supervision head (subsampled), loss, every parameter's gradient norm, and the fp64
evaluation of the same module as the yardstick for the fp32 noise of the reference itself.

Run in the build container only:  python tests/golden/make_golden_medformer.py  ->  tests/golden/reference_medformer.npz
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

from make_golden import import_reference, reference_medformer  # noqa: E402
from oracle import losses_ref as LR  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.medformer_ref import SMALL_CFG, conditioned_state  # noqa: E402
from oracle.unet_ref import synthetic_image  # noqa: E402

C, S = 2, 32


def main():
    os.chdir(tempfile.mkdtemp())
    _, lf = import_reference()[:2]
    net = reference_medformer(SMALL_CFG, C)
    sd = conditioned_state([(k, tuple(v.shape)) for k, v in net.named_parameters()])
    net.load_state_dict(sd, strict=True)
    x = synthetic_image(1, S, S, S, seed=3)
    mo = net(x)
    classes = ["organ", "pancreatic_lesion"]
    batch = synth.make_batch(["mask"], classes, (S, S, S), seed=5)
    args = LR.default_args(report_volume_loss_basic=0.0)
    loss = lf.calculate_loss(model_output=mo, label=batch["label"].long(), unk_voxels=None, args=args, matcher=None,
                             chosen_segment_mask=None, tumor_volumes_report=None, tumor_diameters=None, classes=classes)
    loss["overall"].backward()
    out = {"names": np.array([k for k, _ in net.named_parameters()]),
           "shapes": np.array([",".join(str(d) for d in v.shape) for _, v in net.named_parameters()]),
           "logits": mo["segmentation"][0].detach().numpy()[:, :, ::2, ::2, ::2].copy(),
           "aux": mo["segmentation"][1].detach().numpy()[:, :, ::2, ::2, ::2].copy(),
           "loss": np.float32(loss["overall"].item()),
           "grad_norms": np.array([p.grad.norm().item() for _, p in net.named_parameters()], dtype=np.float64)}
    with torch.no_grad():
        mo64 = net.double()(x.double())
    out["logits_fp64_rel"] = np.float64(((mo64["segmentation"][0] - mo["segmentation"][0].double()).abs().max()
                                         / mo64["segmentation"][0].abs().max()).item())
    np.savez_compressed(os.path.join(HERE, "reference_medformer.npz"), **out)
    print(out["logits"].shape, out["loss"], len(out["grad_norms"]), "reference fp32 vs fp64:", out["logits_fp64_rel"])


if __name__ == "__main__":
    main()
