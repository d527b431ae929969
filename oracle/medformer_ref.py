"""ORACLE (test infrastructure, never shipped, never imported by the product path) — groundwork for SURVEY §8(f) N1.

Functional restatement of the reference MedFormer forward (fp32, plain torch ops) from a state dict:

  MedFormer.forward / prepare_return        rsuper_train/model/dim3/medformer.py:175-222
  inconv, down_block, up_block               model/dim3/medformer_utils.py:263-378
  PatchMerging                               medformer_utils.py:148-183
  SemanticMapGeneration / SemanticMapFusion  medformer_utils.py:209-262
  BidirectionAttention(Block), BasicLayer    medformer_utils.py:14-146, 185-207
  ConvNormAct, BasicBlock, DepthwiseSeparableConv, SEBlock, MBConv   model/dim3/conv_layers.py:16-94, 125-230
  TransformerBlock (PreNorm, Attention, Mlp) model/dim3/trans_layers.py:17-118

Configuration = config/abdomenatlas_ufo/medformer_3d.yaml (norm 'in', act 'relu', proj_type 'depthwise', BasicBlock conv
blocks, 3x3x3 kernels, x2 scales, map_size [3,3,3]).  Two different InstanceNorm epsilons are in play, exactly like in the
reference: ConvNormAct builds `norm(ch, eps=1e-4)` (conv_layers.py:39-42) while PatchMerging / BidirectionAttentionBlock
build `norm(dim)` with torch's default 1e-5 (medformer_utils.py:122-123, 168).

Pinned against the real reference module by tests/golden/make_golden.py (key `medformer_*`) and
tests/test_oracle_golden.py::test_medformer_matches_reference.  No CUDA kernels consume this yet: it is the parity gate the
next round builds MedFormer (depthwise 3^3 convs, PatchMerging, 27-token bidirectional attention, MBConv) against.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

EPS_CNA = 1e-4   # ConvNormAct: norm(ch, eps=1e-4)
EPS_DEF = 1e-5   # nn.InstanceNorm3d default (PatchMerging.norm, BidirectionAttentionBlock.norm1/norm2)

DEFAULT_CFG = dict(base_chan=32, chan_num=[64, 128, 256, 320, 256, 128, 64, 32], conv_num=[2, 0, 0, 0, 0, 0, 2, 2],
                   trans_num=[0, 2, 4, 6, 4, 2, 0, 0], num_heads=[1, 4, 8, 10, 8, 4, 1, 1], map_size=[3, 3, 3], expansion=4,
                   fusion_depth=2, fusion_dim=320, fusion_heads=10, aux_loss=True)


def _cna(x, w, *, norm=True, act=True, groups=1):
    """ConvNormAct(preact=True): conv(act(norm(x))) (conv_layers.py:47-49); padding = k // 2, no bias."""
    h = x
    if norm:
        h = F.instance_norm(h, eps=EPS_CNA)
    if act:
        h = F.relu(h)
    return F.conv3d(h, w, padding=w.shape[-1] // 2, groups=groups)


def _basic_block(x, sd, pre):
    """BasicBlock (conv_layers.py:71-94)."""
    out = _cna(_cna(x, sd[pre + "conv1.conv.weight"]), sd[pre + "conv2.conv.weight"])
    key = pre + "shortcut.conv.weight"
    return out + (_cna(x, sd[key]) if key in sd else x)


def _dsconv(x, sd, pre):
    """DepthwiseSeparableConv (conv_layers.py:125-157): depthwise k^3 (groups = C) then pointwise 1^3, no bias."""
    wd = sd[pre + "depthwise.weight"]
    h = F.conv3d(x, wd, padding=wd.shape[-1] // 2, groups=wd.shape[0])
    return F.conv3d(h, sd[pre + "pointwise.weight"])


def _mbconv(x, sd, pre):
    """MBConv with SE (conv_layers.py:192-230), in_ch == out_ch, stride 1: identity shortcut, DropPath(p=0) = identity."""
    h = _cna(x, sd[pre + "expand_proj.conv.weight"])
    wd = sd[pre + "depthwise.conv.weight"]
    h = _cna(h, wd, groups=wd.shape[0])
    s = h.mean(dim=(2, 3, 4), keepdim=True)                                            # SEBlock (conv_layers.py:159-173)
    s = F.relu(F.conv3d(s, sd[pre + "se.excitation.0.weight"], sd[pre + "se.excitation.0.bias"]))
    s = torch.sigmoid(F.conv3d(s, sd[pre + "se.excitation.2.weight"], sd[pre + "se.excitation.2.bias"]))
    h = h * s
    h = _cna(h, sd[pre + "pointwise.conv.weight"], act=False)
    return h + x


def _split_heads(x, heads):
    """'b (dim_head heads) d h w -> b heads (d h w) dim_head' (medformer_utils.py:46-55)."""
    b, l = x.shape[:2]
    return x.reshape(b, l // heads, heads, -1).permute(0, 2, 3, 1)


def _merge_heads(x, d, h, w):
    """'b heads (d h w) dim_head -> b (dim_head heads) d h w' (medformer_utils.py:56-63)."""
    b, heads, _, dim_head = x.shape
    return x.permute(0, 3, 1, 2).reshape(b, heads * dim_head, d, h, w)


def _bidirection_attention(feat, smap, sd, pre, heads, map_size):
    """BidirectionAttention.forward (medformer_utils.py:67-103), proj_type='depthwise'."""
    D, H, W = feat.shape[2:]
    feat_q, feat_v = _dsconv(feat, sd, pre + "feat_qv.").chunk(2, dim=1)
    map_q, map_v = F.conv3d(smap, sd[pre + "map_qv.weight"]).chunk(2, dim=1)
    dim_head = feat_q.shape[1] // heads
    feat_q, feat_v, map_q, map_v = (_split_heads(t, heads) for t in (feat_q, feat_v, map_q, map_v))
    attn = torch.einsum("bhid,bhjd->bhij", feat_q, map_q) * dim_head ** (-0.5)
    feat_map_attn = F.softmax(attn, dim=-1)
    map_feat_attn = F.softmax(attn, dim=-2)
    feat_out = _merge_heads(torch.einsum("bhij,bhjd->bhid", feat_map_attn, map_v), D, H, W)
    map_out = _merge_heads(torch.einsum("bhji,bhjd->bhid", map_feat_attn, feat_v), *map_size)
    feat_out = _dsconv(feat_out, sd, pre + "feat_out.")
    key = pre + "map_out.weight"                      # absent when no_map_out (nn.Identity)
    if key in sd:
        map_out = F.conv3d(map_out, sd[key])
    return feat_out, map_out


def _attention_block(x, smap, sd, pre, heads, map_size):
    """BidirectionAttentionBlock.forward (medformer_utils.py:134-146)."""
    out, mapp = _bidirection_attention(F.instance_norm(x, eps=EPS_DEF), F.instance_norm(smap, eps=EPS_DEF), sd, pre + "attn.", heads,
                                       map_size)
    key = pre + "shortcut.conv.weight"
    out = out + (_cna(x, sd[key]) if key in sd else x)
    out = _mbconv(out, sd, pre + "feedforward.")
    return out, mapp + smap


def _basic_layer(x, smap, sd, pre, num_blocks, heads, map_size):
    for i in range(num_blocks):
        x, smap = _attention_block(x, smap, sd, f"{pre}blocks.{i}.", heads, map_size)
    return x, smap


def _patch_merging(x, sd, pre):
    """PatchMerging.forward (medformer_utils.py:170-183), down_scale [2,2,2]: 8 strided sub-grids concatenated on channels."""
    parts = [x[:, :, i::2, j::2, k::2] for i in range(2) for j in range(2) for k in range(2)]
    return _dsconv(F.instance_norm(torch.cat(parts, 1), eps=EPS_DEF), sd, pre + "reduction.")


def _map_generation(x, sd, pre, map_size):
    """SemanticMapGeneration.forward (medformer_utils.py:222-235)."""
    B = x.shape[0]
    feat = F.conv3d(x, sd[pre + "base_proj.weight"], padding=1)
    wmap = F.conv3d(x, sd[pre + "semantic_proj.weight"], padding=1)
    codes = wmap.shape[1]
    wmap = F.softmax(wmap.reshape(B, codes, -1), dim=2)
    smap = torch.einsum("bij,bkj->bik", feat.reshape(B, feat.shape[1], -1), wmap)
    return smap.reshape(B, feat.shape[1], *map_size)


def _transformer(x, sd, pre, depth, heads):
    """TransformerBlock (trans_layers.py:104-118): PreNorm(LayerNorm) attention + PreNorm MLP, residuals."""
    dim = x.shape[-1]
    for i in range(depth):
        p = f"{pre}layers.{i}.0."
        h = F.layer_norm(x, (dim,), sd[p + "norm.weight"], sd[p + "norm.bias"])
        q, k, v = F.linear(h, sd[p + "fn.to_qkv.weight"]).chunk(3, dim=-1)
        b, l, n = q.shape
        q, k, v = (t.reshape(b, l, heads, n // heads).permute(0, 2, 1, 3) for t in (q, k, v))
        attn = F.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * (n // heads) ** (-0.5), dim=-1)
        a = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, l, n)
        x = F.linear(a, sd[p + "fn.to_out.weight"], sd[p + "fn.to_out.bias"]) + x
        p = f"{pre}layers.{i}.1."
        h = F.layer_norm(x, (dim,), sd[p + "norm.weight"], sd[p + "norm.bias"])
        h = F.linear(F.gelu(F.linear(h, sd[p + "fn.fc1.weight"], sd[p + "fn.fc1.bias"])), sd[p + "fn.fc2.weight"], sd[p + "fn.fc2.bias"])
        x = h + x
    return x


def _map_fusion(maps: List[torch.Tensor], sd, depth, heads):
    """SemanticMapFusion.forward (medformer_utils.py:252-262)."""
    B, _, D, H, W = maps[0].shape
    proj = [F.conv3d(m, sd[f"map_fusion.in_proj.{i}.weight"]) for i, m in enumerate(maps)]
    dim = proj[0].shape[1]
    tokens = torch.cat([p.reshape(B, dim, -1).permute(0, 2, 1) for p in proj], dim=1)
    tokens = _transformer(tokens, sd, "map_fusion.fusion.", depth, heads)
    outs = tokens.chunk(len(maps), dim=1)
    return [F.conv3d(o.permute(0, 2, 1).reshape(B, dim, D, H, W), sd[f"map_fusion.out_proj.{i}.weight"]) for i, o in enumerate(outs)]


def _down(x, sd, pre, conv_num, trans_num, heads, map_size, map_generate):
    """down_block.forward (medformer_utils.py:311-325)."""
    x = _patch_merging(x, sd, pre + "patch_merging.")
    for i in range(conv_num):
        x = _basic_block(x, sd, f"{pre}conv_blocks.{i}.")
    smap = _map_generation(x, sd, pre + "map_gen.", map_size) if map_generate else None
    return _basic_layer(x, smap, sd, pre + "trans_blocks.", trans_num, heads, map_size)


def _up(x1, x2, map1, map2, sd, pre, conv_num, trans_num, heads, map_size):
    """up_block.forward (medformer_utils.py:358-378)."""
    x1 = F.interpolate(x1, size=x2.shape[-3:], mode="trilinear", align_corners=True)
    feat = torch.cat([x1, x2], dim=1)
    key = pre + "map_reduction.weight"
    smap = F.conv3d(torch.cat([map1, map2], dim=1), sd[key]) if (key in sd and map2 is not None) else map1
    out, smap = _basic_layer(feat, smap, sd, pre + "trans_blocks.", trans_num, heads, map_size)
    for i in range(conv_num):
        out = _basic_block(out, sd, f"{pre}conv_blocks.{i}.")
    return out, smap


def medformer_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: Optional[dict] = None, trace: Optional[dict] = None):
    """x [N,1,D,H,W] -> {'segmentation': [logits, aux]} (aux_loss) or {'segmentation': logits} (medformer.py:175-222)."""
    c = dict(DEFAULT_CFG)
    c.update(cfg or {})
    cn, tn, nh, ms = c["conv_num"], c["trans_num"], c["num_heads"], list(c["map_size"])
    tr = trace if trace is not None else {}
    x0 = _basic_block(F.conv3d(x, sd["inc.conv1.weight"], padding=1), sd, "inc.conv2.")
    x1, _ = _down(x0, sd, "down1.", cn[0], tn[0], nh[0], ms, False)
    x2, m2 = _down(x1, sd, "down2.", cn[1], tn[1], nh[1], ms, True)
    x3, m3 = _down(x2, sd, "down3.", cn[2], tn[2], nh[2], ms, True)
    x4, m4 = _down(x3, sd, "down4.", cn[3], tn[3], nh[3], ms, True)
    tr.update(x0=x0, x1=x1, x2=x2, x3=x3, x4=x4, map2=m2, map3=m3, map4=m4)
    maps = _map_fusion([m2, m3, m4], sd, c["fusion_depth"], c["fusion_heads"])
    out, smap = _up(x4, x3, maps[2], maps[1], sd, "up1.", cn[4], tn[4], nh[4], ms)
    out, smap = _up(out, x2, smap, maps[0], sd, "up2.", cn[5], tn[5], nh[5], ms)
    aux = None
    if c["aux_loss"]:
        aux = F.interpolate(F.conv3d(out, sd["aux_out.weight"], sd["aux_out.bias"]), size=x.shape[-3:], mode="trilinear", align_corners=True)
    out, smap = _up(out, x1, smap, None, sd, "up3.", cn[6], tn[6], nh[6], ms)
    out, smap = _up(out, x0, smap, None, sd, "up4.", cn[7], tn[7], nh[7], ms)
    logits = F.conv3d(out, sd["outc.weight"], sd["outc.bias"])
    return {"segmentation": [logits, aux] if c["aux_loss"] else logits}


SMALL_CFG = dict(base_chan=8, chan_num=[16, 32, 64, 80, 64, 32, 16, 8], conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 4, 6, 4, 2, 0, 0],
                 num_heads=[1, 4, 8, 10, 8, 4, 1, 1], map_size=[3, 3, 3], expansion=4, fusion_depth=2, fusion_dim=80, fusion_heads=10,
                 aux_loss=True)


def fill_like(named_shapes: Sequence[Tuple[str, Tuple[int, ...]]], gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Deterministic, version-independent values for a list of (name, shape): the hash init of unet_ref.synthetic_state_dict
    (conv / linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); LayerNorm weights 1 + small, biases small)."""
    sd = {}
    for k, (name, shp) in enumerate(named_shapes):
        numel = 1
        for s in shp:
            numel *= s
        idx = torch.arange(numel, dtype=torch.float64)
        u = torch.frac(torch.sin(idx * 12.9898 + k * 78.233) * 43758.5453) * 2 - 1
        if len(shp) > 1:
            v = u * gain / ((numel // shp[0]) ** 0.5)
        elif name.endswith("norm.weight"):
            v = 1.0 + 0.1 * u
        else:
            v = 0.1 * u
        sd[name] = v.to(torch.float32).reshape(shp)
    return sd


SEMANTIC_PROJ_SCALE = 0.02


def conditioned_state(named_shapes: Sequence[Tuple[str, Tuple[int, ...]]]) -> Dict[str, torch.Tensor]:
    """fill_like with the `map_gen.semantic_proj` weights scaled by SEMANTIC_PROJ_SCALE.  With the plain hash init the code
    logits of SemanticMapGeneration span +-1000, the softmax over the voxels is one-hot on ONE voxel for all 27 codes, every
    channel of the semantic map is constant and the InstanceNorm that follows (medformer_utils.py:134-136) returns rounding
    noise: the reference's own fp32 and fp64 evaluations then differ by 2.8e-2 on the logits.  Scaled, they agree to 3e-6 —
    a state on which an independent implementation CAN be compared (tests/golden/make_golden_medformer.py records the real
    reference on it)."""
    sd = fill_like(named_shapes)
    for k in sd:
        if k.endswith("semantic_proj.weight"):
            sd[k] = sd[k] * SEMANTIC_PROJ_SCALE
    return sd

