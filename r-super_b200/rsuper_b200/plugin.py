"""get_model — the plugin point of rsuper_train/model/utils.py:11-134 for the B200 path.

A maintainer adds ONE branch to the reference's get_model:

    elif args.model == 'b200_unet':
        from rsuper_b200.plugin import get_model as b200_get_model
        return b200_get_model(args, pretrain, classes)

selected with `--model b200_unet --dimension 3d` plus config/<dataset>/b200_unet_3d.yaml (a copy of
config/abdomenatlas/resunet_3d.yaml:9-14).  Argument meaning follows the reference's resunet branch
(model/utils.py:52-56).
"""
from __future__ import annotations

import torch

from .unet import B200UNet


def get_model(args, pretrain: bool = False, classes=None):
    if getattr(args, "dimension", "3d") != "3d":
        raise ValueError("b200_unet is a 3d model")
    if pretrain:
        raise ValueError("No pretrain model available")  # same message as model/utils.py:48,53
    num_classes = len(classes) if classes is not None else args.classes
    net = B200UNet(args.in_chan, args.base_chan, num_classes=num_classes, scale=args.down_scale,
                   norm=args.norm, kernel_size=args.kernel_size, block=args.block,
                   negative_slope=getattr(args, "negative_slope", 0.0),
                   precision=getattr(args, "precision", "bf16"), up_mode=getattr(args, "up_mode", "trilinear"))
    return net
