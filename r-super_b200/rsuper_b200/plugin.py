"""get_model — the plugin point of rsuper_train/model/utils.py:11-134 for the B200 path.

A maintainer adds ONE branch to the reference's get_model:

    elif args.model == 'b200_unet':
        from rsuper_b200.plugin import get_model as b200_get_model
        return b200_get_model(args, pretrain, classes)

selected with `--model b200_unet --dimension 3d` plus config/<dataset>/b200_unet_3d.yaml (a copy of
config/abdomenatlas/resunet_3d.yaml:9-14).  Argument meaning follows the reference's resunet branch
(model/utils.py:52-56).  `--model b200_medformer` builds B200MedFormer from the keys of config/abdomenatlas_ufo/medformer_3d.yaml
exactly like the reference's medformer branch (model/utils.py:97-133).
"""
from __future__ import annotations

import torch

from .unet import B200UNet


def _get_medformer(args, pretrain, classes):
    """model/utils.py:97-133 for B200MedFormer: same yaml keys, same pretrain handling (state dict loaded non-strictly)."""
    from .medformer import B200MedFormer
    if getattr(args, "classification_branch", False) or getattr(args, "clip_loss", False):
        raise NotImplementedError("b200_medformer: the classification / CLIP branches are outside the segmentation path")
    num_classes = len(classes) if classes is not None else args.classes
    net = B200MedFormer(args.in_chan, num_classes, args.base_chan, map_size=args.map_size, conv_block=args.conv_block,
                        conv_num=args.conv_num, trans_num=args.trans_num, num_heads=args.num_heads, fusion_depth=args.fusion_depth,
                        fusion_dim=args.fusion_dim, fusion_heads=args.fusion_heads, expansion=args.expansion, attn_drop=args.attn_drop,
                        proj_drop=args.proj_drop, proj_type=args.proj_type, norm=args.norm, act=args.act, kernel_size=args.kernel_size,
                        scale=args.down_scale, aux_loss=args.aux_loss, precision=getattr(args, "precision", "bf16"),
                        **({"chan_num": args.chan_num} if hasattr(args, "chan_num") else {}))
    if pretrain:
        checkpoint = torch.load(args.pretrained, weights_only=False)
        pretrained_model = checkpoint["model_state_dict"]
        state_dict = pretrained_model.state_dict() if hasattr(pretrained_model, "state_dict") else pretrained_model
        net.load_state_dict(state_dict, strict=False)
    return net


def get_model(args, pretrain: bool = False, classes=None):
    if getattr(args, "model", "b200_unet") == "b200_medformer":
        if getattr(args, "dimension", "3d") != "3d":
            raise ValueError("b200_medformer is a 3d model")
        return _get_medformer(args, pretrain, classes)
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
