"""CPU: the host-side contract of B200MedFormer (no kernels run): parameter names / shapes / order of the REAL reference module
(recorded by tests/golden/make_golden*.py), constructor checks, deepcopy / pickle, the plugin branch, and that the module
refuses CPU tensors instead of falling back."""
import copy
import pickle
import types

import numpy as np
import pytest
import torch


def _cfg():
    from oracle.medformer_ref import SMALL_CFG as c
    return dict(base_chan=c["base_chan"], map_size=c["map_size"], conv_block="BasicBlock", conv_num=c["conv_num"], trans_num=c["trans_num"],
                chan_num=c["chan_num"], num_heads=c["num_heads"], fusion_depth=c["fusion_depth"], fusion_dim=c["fusion_dim"],
                fusion_heads=c["fusion_heads"], expansion=c["expansion"], proj_type="depthwise", norm="in", act="relu",
                kernel_size=[[3, 3, 3]] * 5, scale=[[2, 2, 2]] * 4, aux_loss=True)


def test_medformer_parameter_tree_matches_the_reference_module(golden):
    from rsuper_b200.medformer import B200MedFormer
    net = B200MedFormer(1, 2, **_cfg())
    names = [str(k) for k in golden["medformer_param_names"]]
    shapes = [tuple(int(d) for d in str(s).split(",")) for s in golden["medformer_param_shapes"]]
    got = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    assert got == list(zip(names, shapes))                       # same names, shapes AND order as the real module (300 tensors)
    assert all(torch.isfinite(v).all() for _, v in net.named_parameters())
    # without the deep-supervision head the aux_out parameters disappear, nothing else changes
    cfg = _cfg()
    cfg["aux_loss"] = False
    plain = [k for k, _ in B200MedFormer(1, 2, **cfg).named_parameters()]
    assert plain == [k for k in names if not k.startswith("aux_out.")]


def test_medformer_full_configuration_parameter_count():
    """config/abdomenatlas_ufo/medformer_3d.yaml (base 32): the shapes follow the reference's channel arithmetic."""
    from rsuper_b200.medformer import B200MedFormer
    net = B200MedFormer(1, 2, base_chan=32, map_size=[3, 3, 3], conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 4, 6, 4, 2, 0, 0],
                        chan_num=[64, 128, 256, 320, 256, 128, 64, 32], num_heads=[1, 4, 8, 10, 8, 4, 1, 1], fusion_depth=2, fusion_dim=320,
                        fusion_heads=10, expansion=4, aux_loss=True)
    P = dict(net.named_parameters())
    assert tuple(P["down4.patch_merging.reduction.depthwise.weight"].shape) == (2048, 1, 3, 3, 3)
    assert tuple(P["down4.trans_blocks.blocks.5.feedforward.expand_proj.conv.weight"].shape) == (1280, 320, 1, 1, 1)
    assert tuple(P["up1.trans_blocks.blocks.0.attn.feat_qv.pointwise.weight"].shape) == (512, 576, 1, 1, 1)
    assert "up2.trans_blocks.blocks.1.attn.map_out.weight" not in P and "up2.trans_blocks.blocks.0.attn.map_out.weight" in P
    assert 30e6 < sum(p.numel() for p in P.values()) < 60e6


def test_medformer_module_contract_and_refusals():
    from rsuper_b200.medformer import B200MedFormer
    net = B200MedFormer(1, 2, **_cfg())
    twin = copy.deepcopy(net)
    assert [k for k, _ in twin.named_parameters()] == [k for k, _ in net.named_parameters()]
    assert all(a.data_ptr() != b.data_ptr() and torch.equal(a, b) for a, b in zip(net.parameters(), twin.parameters()))
    back = pickle.loads(pickle.dumps(net))
    back.load_state_dict(net.state_dict(), strict=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(torch.zeros(1, 1, 32, 32, 32))
    for bad in (dict(conv_block="Bottleneck"), dict(norm="bn"), dict(act="gelu"), dict(proj_type="linear"), dict(map_size=[4, 4, 4]),
                dict(attn_drop=0.1)):
        cfg = _cfg()
        cfg.update(bad)
        with pytest.raises(NotImplementedError):
            B200MedFormer(1, 2, **cfg)
    with pytest.raises(NotImplementedError):
        B200MedFormer(3, 2, **_cfg())


def test_medformer_plugin_branch_reads_the_yaml_keys():
    """rsuper_b200.plugin.get_model(args) with args.model == 'b200_medformer' (model/utils.py:97-133)."""
    from rsuper_b200.medformer import B200MedFormer
    from rsuper_b200.plugin import get_model
    c = _cfg()
    args = types.SimpleNamespace(model="b200_medformer", dimension="3d", in_chan=1, classes=5, down_scale=c.pop("scale"),
                                 attn_drop=0.0, proj_drop=0.0, classification_branch=False, clip_loss=False, **c)
    net = get_model(args, pretrain=False, classes=["organ", "pancreatic_lesion"])
    assert isinstance(net, B200MedFormer) and net.num_classes == 2 and net.aux_loss and net.precision == "bf16"
    assert get_model(args, classes=None).num_classes == 5
    args.classification_branch = True
    with pytest.raises(NotImplementedError):
        get_model(args, classes=None)
