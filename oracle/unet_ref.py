"""ORACLE (test infrastructure, never shipped, never imported by the product path).

Functional restatement of the reference 3D UNet forward (fp32, plain torch ops, runs on CPU or
CUDA) from a state dict, following

  UNet.forward                         rsuper_train/model/dim3/unet.py:50-64
  inconv / down_block / up_block       rsuper_train/model/dim3/unet_utils.py:7-75
  BasicBlock, ConvNormAct(preact)      rsuper_train/model/dim3/conv_layers.py:16-94

Pinned against the real reference module by tests/golden/make_golden.py (run in the build
container, where /root/reference is importable); the committed vectors are checked by
tests/test_oracle_golden.py.  Gradients come from autograd over this forward.

`emulate` reproduces the rounding points of the bf16 tensor-core path so that kernels can be
checked tightly: operands of every conv (activated input, weights) are rounded to bf16;
with storage='bf16' every stored activation is rounded too (statistics are still taken from the
fp32 values before rounding, exactly like the kernels' epilogues do).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS_IN = 1e-4  # conv_layers.py:40-42: norm(ch, eps=1e-4)


def _r16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class _Cfg:
    def __init__(self, slope: float, emulate: bool, storage: str):
        self.slope, self.emulate, self.storage = slope, emulate, storage

    def store(self, t):  # rounding of a tensor written to HBM
        return _r16(t) if (self.emulate and self.storage == "bf16") else t

    def operand(self, t):  # rounding of a tensor-core operand
        return _r16(t) if self.emulate else t


def _norm_act(x_stored, x_full, cfg: _Cfg):
    """InstanceNorm3d(eps=1e-4, affine=False) + ReLU (conv_layers.py:39-43,47-49).

    x_full is the fp32 value the statistics are taken from, x_stored the (possibly rounded) tensor
    that is normalised; without emulation they are the same tensor and this is F.instance_norm."""
    if x_stored is x_full:
        h = F.instance_norm(x_stored, eps=EPS_IN)
    else:
        mean = x_full.mean(dim=(2, 3, 4), keepdim=True)
        var = x_full.var(dim=(2, 3, 4), keepdim=True, unbiased=False)
        h = (x_stored - mean) * torch.rsqrt(var + EPS_IN)
    return F.leaky_relu(h, cfg.slope) if cfg.slope != 0.0 else F.relu(h)


def _cna(xs, xf, w, cfg):
    """ConvNormAct with preact=True: conv(act(norm(x))) (conv_layers.py:47-49)."""
    a = cfg.operand(_norm_act(xs, xf, cfg))
    return F.conv3d(a, cfg.operand(w), padding=1)


def _basic_block(xs, xf, sd, prefix, cfg):
    """BasicBlock.forward (conv_layers.py:85-94); returns (stored, full) of the block output."""
    h_full = _cna(xs, xf, sd[prefix + "conv1.conv.weight"], cfg)
    h_st = cfg.store(h_full)
    out = _cna(h_st, h_full, sd[prefix + "conv2.conv.weight"], cfg)
    key = prefix + "shortcut.conv.weight"
    if key in sd:
        s_full = _cna(xs, xf, sd[key], cfg)
        out = out + cfg.store(s_full)
    else:
        out = out + xs
    return cfg.store(out), out


def _cna_k(xs, xf, w, cfg):
    """ConvNormAct(preact=True) with the kernel size read off the weight (1x1x1: padding 0; 3x3x3: padding 1)."""
    a = cfg.operand(_norm_act(xs, xf, cfg))
    return F.conv3d(a, cfg.operand(w), padding=w.shape[-1] // 2)


def _bottleneck(xs, xf, sd, prefix, cfg):
    """Bottleneck.forward (conv_layers.py:97-123): 1x1x1 (C -> C/2) -> 3x3x3 -> 1x1x1 (C/2 -> C), all pre-activation, plus
    the 3x3x3 ConvNormAct shortcut when the channel count changes; returns (stored, full) of the block output."""
    h1_full = _cna_k(xs, xf, sd[prefix + "conv1.conv.weight"], cfg)
    h1_st = cfg.store(h1_full)
    h2_full = _cna_k(h1_st, h1_full, sd[prefix + "conv2.conv.weight"], cfg)
    h2_st = cfg.store(h2_full)
    out = _cna_k(h2_st, h2_full, sd[prefix + "conv3.conv.weight"], cfg)
    key = prefix + "shortcut.conv.weight"
    if key in sd:
        out = out + cfg.store(_cna_k(xs, xf, sd[key], cfg))
    else:
        out = out + xs
    return cfg.store(out), out


def _single_conv(xs, sd, prefix, cfg):
    """SingleConv = ConvNormAct(preact=False): act(norm(conv(x))) (conv_layers.py:50-68); returns the stored output.
    With emulation the raw conv output is what gets stored / re-read (statistics from the fp32 accumulators)."""
    y_full = F.conv3d(cfg.operand(xs), cfg.operand(sd[prefix + "conv.conv.weight"]), padding=1)
    return cfg.store(_norm_act(cfg.store(y_full), y_full, cfg))


def _unet_forward_single(x, sd, cfg, tr):
    """UNet(block='SingleConv') (model/dim3/utils.py:7-13 block map; same inconv / down / up skeleton)."""
    xs = _single_conv(cfg.store(F.conv3d(x, sd["inc.conv1.weight"], padding=1)), sd, "inc.conv2.", cfg)
    tr["inc"] = xs
    skips = [xs]
    for l in range(1, 5):
        ys = _single_conv(F.max_pool3d(xs, 2), sd, f"down{l}.conv.1.", cfg)
        xs = _single_conv(ys, sd, f"down{l}.conv.2.", cfg)
        tr[f"down{l}.2"] = xs
        skips.append(xs)
    cur = skips[4]
    for j, l in enumerate((3, 2, 1, 0), start=1):
        up = cfg.store(F.interpolate(cur, size=skips[l].shape[2:], mode="trilinear", align_corners=True))
        ys = _single_conv(torch.cat([skips[l], up], dim=1), sd, f"up{j}.conv.0.", cfg)
        cur = _single_conv(ys, sd, f"up{j}.conv.1.", cfg)
        tr[f"up{j}.1"] = cur
    return F.conv3d(cur, sd["outc.weight"], sd["outc.bias"])


def unet_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], slope: float = 0.0, emulate: bool = False,
                 storage: str = "fp32", trace: Optional[dict] = None) -> torch.Tensor:
    """x [N,1,D,H,W] fp32 -> logits [N,C,D,H,W] (unet.py:50-64).  `trace` collects stage outputs.  The block type
    (BasicBlock | SingleConv) is read off the state-dict keys."""
    cfg = _Cfg(slope, emulate, storage)
    tr = trace if trace is not None else {}
    if "inc.conv2.conv.conv.weight" in sd:
        return _unet_forward_single(x, sd, cfg, tr)
    block = _bottleneck if "inc.conv2.conv3.conv.weight" in sd else _basic_block
    # inconv: raw conv then BasicBlock (unet_utils.py:17-21)
    t_full = F.conv3d(x, sd["inc.conv1.weight"], padding=1)
    t_st = cfg.store(t_full)
    xs, xf = block(t_st, t_full, sd, "inc.conv2.", cfg)
    tr["t0"], tr["inc"] = t_st, xs
    skips = [(xs, xf)]
    for l in range(1, 5):  # down_block: MaxPool3d then two blocks (unet_utils.py:33-41)
        ps = F.max_pool3d(xs, 2)
        ys, yf = block(ps, ps, sd, f"down{l}.conv.1.", cfg)
        xs, xf = block(ys, yf, sd, f"down{l}.conv.2.", cfg)
        tr[f"pool{l}"], tr[f"down{l}.1"], tr[f"down{l}.2"] = ps, ys, xs
        skips.append((xs, xf))
    cur = skips[4][0]
    for j, l in enumerate((3, 2, 1, 0), start=1):  # up_block (unet_utils.py:68-75)
        sk_s, sk_f = skips[l]
        if f"up{j}.up.weight" in sd:
            # up_mode='transposed': ConvTranspose3d(C, C, kernel_size=2, stride=2) in place of F.interpolate (vnet.py:108 semantics)
            up_full = F.conv_transpose3d(cfg.operand(cur), cfg.operand(sd[f"up{j}.up.weight"]), sd[f"up{j}.up.bias"], stride=2)
        else:
            up_full = F.interpolate(cur, size=sk_s.shape[2:], mode="trilinear", align_corners=True)
        up_st = cfg.store(up_full)
        cat_s = torch.cat([sk_s, up_st], dim=1)
        cat_f = torch.cat([sk_f, up_full], dim=1)
        ys, yf = block(cat_s, cat_f, sd, f"up{j}.conv.0.", cfg)
        cur, _ = block(ys, yf, sd, f"up{j}.conv.1.", cfg)
        tr[f"up{j}.cat"], tr[f"up{j}.0"], tr[f"up{j}.1"] = cat_s, ys, cur
    w = sd["outc.weight"]
    return F.conv3d(cur, w, sd["outc.bias"])  # unet.py:62


def synthetic_state_dict(base_ch: int, num_classes: int, in_ch: int = 1, device="cpu",
                         gain: float = 1.0, block: str = "BasicBlock", up_mode: str = "trilinear") -> Dict[str, torch.Tensor]:
    """Deterministic, version-independent weights with the reference UNet's names and shapes
    (SURVEY.md §8b state-dict contract).  Values mimic nn.Conv3d's default init
    (kaiming_uniform(a=sqrt(5)) => U(-1/sqrt(fan_in), 1/sqrt(fan_in))) through a hash
    frac(sin(12.9898 i + 78.233 k) * 43758.5453) evaluated in float64 on the host.  (A smooth
    sin(i) pattern makes the 5-level InstanceNorm stack so ill-conditioned that the reference's own
    fp32 gradients differ from its fp64 gradients by 2 %; with this init they agree to ~3e-5.)"""
    b = base_ch
    ch = [b, 2 * b, 4 * b, 8 * b, 10 * b]
    if block == "SingleConv":
        shapes = [("inc.conv1.weight", (b, in_ch, 3, 3, 3)), ("inc.conv2.conv.conv.weight", (b, b, 3, 3, 3))]
        for l in range(1, 5):
            shapes += [(f"down{l}.conv.1.conv.conv.weight", (ch[l], ch[l - 1], 3, 3, 3)),
                       (f"down{l}.conv.2.conv.conv.weight", (ch[l], ch[l], 3, 3, 3))]
        for j, l in enumerate((3, 2, 1, 0), start=1):
            shapes += [(f"up{j}.conv.0.conv.conv.weight", (ch[l], ch[l] + ch[l + 1], 3, 3, 3)),
                       (f"up{j}.conv.1.conv.conv.weight", (ch[l], ch[l], 3, 3, 3))]
        shapes += [("outc.weight", (num_classes, b, 1, 1, 1)), ("outc.bias", (num_classes,))]
        return _fill_state_dict(shapes, gain, device)
    if block == "Bottleneck":
        def bneck(p, ci, co):
            out = [(p + "conv1.conv.weight", (co // 2, ci, 1, 1, 1)), (p + "conv2.conv.weight", (co // 2, co // 2, 3, 3, 3)),
                   (p + "conv3.conv.weight", (co, co // 2, 1, 1, 1))]
            if ci != co:
                out.append((p + "shortcut.conv.weight", (co, ci, 3, 3, 3)))
            return out
        shapes = [("inc.conv1.weight", (b, in_ch, 3, 3, 3))] + bneck("inc.conv2.", b, b)
        for l in range(1, 5):
            shapes += bneck(f"down{l}.conv.1.", ch[l - 1], ch[l]) + bneck(f"down{l}.conv.2.", ch[l], ch[l])
        for j, l in enumerate((3, 2, 1, 0), start=1):
            if up_mode == "transposed":
                shapes += [(f"up{j}.up.weight", (ch[l + 1], ch[l + 1], 2, 2, 2)), (f"up{j}.up.bias", (ch[l + 1],))]
            shapes += bneck(f"up{j}.conv.0.", ch[l] + ch[l + 1], ch[l]) + bneck(f"up{j}.conv.1.", ch[l], ch[l])
        shapes += [("outc.weight", (num_classes, b, 1, 1, 1)), ("outc.bias", (num_classes,))]
        return _fill_state_dict(shapes, gain, device)
    assert block == "BasicBlock", block
    shapes = [("inc.conv1.weight", (b, in_ch, 3, 3, 3)),
              ("inc.conv2.conv1.conv.weight", (b, b, 3, 3, 3)),
              ("inc.conv2.conv2.conv.weight", (b, b, 3, 3, 3))]
    for l in range(1, 5):
        ci, co = ch[l - 1], ch[l]
        p = f"down{l}.conv."
        shapes += [(p + "1.conv1.conv.weight", (co, ci, 3, 3, 3)), (p + "1.conv2.conv.weight", (co, co, 3, 3, 3)),
                   (p + "1.shortcut.conv.weight", (co, ci, 3, 3, 3)),
                   (p + "2.conv1.conv.weight", (co, co, 3, 3, 3)), (p + "2.conv2.conv.weight", (co, co, 3, 3, 3))]
    for j, l in enumerate((3, 2, 1, 0), start=1):
        ci, co = ch[l] + ch[l + 1], ch[l]
        p = f"up{j}.conv."
        if up_mode == "transposed":
            shapes += [(f"up{j}.up.weight", (ch[l + 1], ch[l + 1], 2, 2, 2)), (f"up{j}.up.bias", (ch[l + 1],))]
        shapes += [(p + "0.conv1.conv.weight", (co, ci, 3, 3, 3)), (p + "0.conv2.conv.weight", (co, co, 3, 3, 3)),
                   (p + "0.shortcut.conv.weight", (co, ci, 3, 3, 3)),
                   (p + "1.conv1.conv.weight", (co, co, 3, 3, 3)), (p + "1.conv2.conv.weight", (co, co, 3, 3, 3))]
    shapes += [("outc.weight", (num_classes, b, 1, 1, 1)), ("outc.bias", (num_classes,))]
    return _fill_state_dict(shapes, gain, device)


def _fill_state_dict(shapes, gain, device):
    sd = {}
    for k, (name, shp) in enumerate(shapes):
        numel = 1
        for s in shp:
            numel *= s
        idx = torch.arange(numel, dtype=torch.float64)
        u = torch.frac(torch.sin(idx * 12.9898 + k * 78.233) * 43758.5453) * 2 - 1  # ~U(-1, 1)
        fan_in = numel // shp[0] if len(shp) > 1 else 1
        bound = gain / (fan_in ** 0.5) if len(shp) > 1 else 0.1
        sd[name] = (u * bound).to(torch.float32).reshape(shp).to(device)
    return sd


from oracle.synth import synthetic_image  # noqa: E402,F401  (data construction lives in rsuper_b200.synthetic)
