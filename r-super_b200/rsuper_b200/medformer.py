"""B200MedFormer — the reference MedFormer (rsuper_train/model/dim3/medformer.py:81-222) on the sm_100a kernels.

Same constructor arguments (the ones the yaml configuration uses), same parameter names and shapes — the state dict is
interchangeable with the reference's `MedFormer(..., conv_block='BasicBlock', norm='in', act='relu', proj_type='depthwise')`.

Where the work runs
  voxel side (every tensor with a D x H x W extent; NDHWC, bf16 or fp32 storage)      hand-written kernels:
    3x3x3 / 1x1x1 convolutions (ConvNormAct, BasicBlock, DepthwiseSeparableConv.pointwise, MBConv expand / project,
      SemanticMapGeneration projections)              rsb_conv3_forward / rsb_conv3_wgrad (tcgen05 implicit GEMM)
    InstanceNorm + ReLU pre-activations and their backward    rsb_norm_act / rsb_act_backward_stats / rsb_instnorm_backward_apply
    depthwise 3x3x3 convolutions                       rsb_dwconv3_forward / _wgrad
    PatchMerging's 2x2x2 gather                        rsb_space_to_depth2 / rsb_depth_to_space2
    SEBlock squeeze / scale                            rsb_channel_stats, rsb_scale_channels, rsb_channel_dot
    SemanticMapGeneration softmax-over-voxels pooling  rsb_softmax_pool_forward / _backward
    BidirectionAttention (both softmaxes, both einsums) rsb_biattention_forward / _backward
    trilinear up-sampling, stem conv, 1x1x1 head       the UNet kernels
  map side (the 3x3x3 = 27-token semantic maps, <= 81 tokens x fusion_dim in SemanticMapFusion, the SE excitation vectors,
    the deep-supervision head at 1/4 resolution): plain torch ops under autograd — a few thousand elements per sample.

Every voxel-side op is a torch.autograd.Function over the C-ABI kernels, so the parameter gradients arrive through autograd
like the reference's.  There is no CPU path: without the CUDA library every call raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

EPS_CNA = 1e-4   # ConvNormAct: norm(ch, eps=1e-4) (conv_layers.py:39-42)
EPS_DEF = 1e-5   # PatchMerging.norm / BidirectionAttentionBlock.norm1, norm2 (torch default)


# ------------------------------------------------------------------------------------------------
# weight gradients beside the backward chain (only under B200TrainStep, which owns the gradient buffers and joins the stream)
# ------------------------------------------------------------------------------------------------
class _Side:
    """Second stream for the weight gradients of the voxel-side convolutions.  The data-gradient chain (tensor-core dgrads,
    InstanceNorm-backward passes) is what the rest of backward waits for; the weight gradients (latency-bound depthwise
    reductions, 3x3x3 wgrad GEMMs, 1x1x1 GEMMs) are needed only by the optimizer.  With `enabled`, a Function's backward
    launches its weight gradient on this stream, adds it into the parameter's preallocated `.grad` there and returns None for
    it; B200TrainStep turns it on around its body (its `.grad`s are views of one flat buffer, zeroed on the main stream before
    the forward pass) and calls `join()` before the all-reduce / optimizer.  Operand tensors stay referenced until the join:
    the caching allocator must not hand their memory to main-stream work while the side stream still reads it."""
    enabled = False
    stream = None
    keep: list = []

    @classmethod
    def get(cls, device):
        if cls.stream is None or cls.stream.device != device:
            cls.stream = torch.cuda.Stream(device=device)
        return cls.stream

    @classmethod
    def usable(cls, p) -> bool:
        return cls.enabled and isinstance(p, torch.Tensor) and p.grad is not None and p.is_cuda

    @classmethod
    def run(cls, p, fn, *operands):
        """fn() -> gradient of p, launched on the side stream and accumulated into p.grad there."""
        side = cls.get(p.device)
        side.wait_stream(torch.cuda.current_stream(p.device))        # operands come from work already enqueued on the main stream
        with torch.cuda.stream(side):
            g = fn()
            p.grad.add_(g.reshape(p.grad.shape))
        cls.keep.append((operands, g))

    @classmethod
    def join(cls):
        if cls.stream is not None and cls.keep:
            torch.cuda.current_stream(cls.stream.device).wait_stream(cls.stream)
        cls.keep = []


def set_side_stream(enabled: bool) -> bool:
    prev, _Side.enabled = _Side.enabled, bool(enabled)
    return prev


def join_side_stream() -> None:
    _Side.join()


# ------------------------------------------------------------------------------------------------
# autograd Functions over the kernels (all activations: contiguous NDHWC)
# ------------------------------------------------------------------------------------------------
class _NormAct(torch.autograd.Function):
    """a = act(instance_norm(x)); slope = 0 -> ReLU, slope = 1 -> no activation."""

    @staticmethod
    def forward(ctx, x, st, eps, slope):
        if st is None:
            st = ops.channel_stats(x)
        a = torch.empty_like(x)
        ops.norm_act(x, st, slope=slope, eps=eps, full=a)
        ctx.save_for_backward(x, st)
        ctx.eps, ctx.slope = eps, slope
        return a

    @staticmethod
    def backward(ctx, da):
        x, st = ctx.saved_tensors
        da = da.contiguous()
        sums = torch.zeros_like(st)
        g = torch.empty_like(x)
        ops.act_backward_stats(da, x, st, sums, g, slope=ctx.slope, eps=ctx.eps)
        dx = torch.empty_like(x)
        ops.instnorm_backward_apply(g, x, st, sums, dx, eps=ctx.eps)
        return dx, None, None, None


def _gemm_tn_f32(dy2d: torch.Tensor, a2d: torch.Tensor) -> torch.Tensor:
    """dy2d [V, Cout]^T @ a2d [V, Cin] -> fp32 [Cout, Cin]: the weight gradient of a 1x1x1 convolution is a plain GEMM over the
    voxels with nothing to fuse — the one place this model calls cuBLAS (bf16 operands, fp32 accumulation and output)."""
    if dy2d.is_cuda:
        return torch.mm(dy2d.t(), a2d, out_dtype=torch.float32)
    return torch.mm(dy2d.t().float(), a2d.float())          # CPU emulation of the test suite


def _operand(t: torch.Tensor, st, slope, eps):
    """bf16 tensor-core operand pieces of a stored tensor: (hi,) for bf16 storage, (hi, lo) for fp32 storage."""
    split = t.dtype == torch.float32
    if st is None and not split:
        return (t,)
    r = ops.norm_act(t, st, slope=slope, eps=eps, split=split)
    return r if split else (r,)


class _ConvNA(torch.autograd.Function):
    """y = conv(act(instance_norm(x))) [+ res] — ConvNormAct(preact=True) (conv_layers.py:47-49) with a 3x3x3 or 1x1x1 kernel on
    the tensor cores; norm=False feeds x itself.  img / img_t: the packed forward / data-gradient weight images (PackPlan)."""

    @staticmethod
    def forward(ctx, x, x_st, w, res, img, img_t, norm, slope, eps, pointwise, cout, want_stats):
        st = (x_st if x_st is not None else ops.channel_stats(x)) if norm else None
        op = _operand(x, st, slope, eps)
        n, d, h, w_, _ = x.shape
        y = torch.empty((n, d, h, w_, cout), dtype=x.dtype, device=x.device)
        y_st = torch.zeros((n, cout, 2), dtype=torch.float32, device=x.device) if want_stats else None   # epilogue: (sum, sumsq) of y
        ops.conv3_forward(op[0], img, y, a_lo=op[1] if len(op) > 1 else None, slope=slope, res=res, out_stats=y_st, eps=eps,
                          pointwise=pointwise)
        ctx.save_for_backward(x, st, img_t, *op)
        ctx.cfg = (norm, slope, eps, pointwise, cout, tuple(w.shape), res is not None)
        ctx.param = w
        if y_st is None:
            y_st = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(y_st)
        return y, y_st

    @staticmethod
    def backward(ctx, dy, _unused):
        x, st, img_t, *op = ctx.saved_tensors
        norm, slope, eps, pointwise, cout, wshape, has_res = ctx.cfg
        dy = dy.contiguous()
        d_op = _operand(dy, None, 0.0, eps)
        split = len(op) > 1
        cin = x.shape[4]
        dw = dx = None
        def weight_gradient():
            if pointwise:
                a2, g2 = [t.reshape(-1, cin) for t in op], [t.reshape(-1, cout) for t in d_op]
                g = _gemm_tn_f32(g2[0], a2[0])
                if split:
                    g = g + _gemm_tn_f32(g2[0], a2[1]) + _gemm_tn_f32(g2[1], a2[0])
                return g[:wshape[0]].reshape(wshape)
            dw27 = torch.empty((cout, cin, 3, 3, 3), dtype=torch.float32, device=x.device)
            ops.conv3_wgrad(op[0], d_op[0], dw27)
            if split:
                ops.conv3_wgrad(op[1], d_op[0], dw27, accumulate=True)
                ops.conv3_wgrad(op[0], d_op[1], dw27, accumulate=True)
            return dw27[:wshape[0]]

        if ctx.needs_input_grad[2]:
            if _Side.usable(ctx.param):
                _Side.run(ctx.param, weight_gradient, op, d_op)       # lands in param.grad on the side stream; autograd gets None
            else:
                dw = weight_gradient()
        if ctx.needs_input_grad[0]:
            lo = d_op[1] if split else None
            dx = torch.empty_like(x)
            if norm:
                g = torch.empty_like(x)
                sums = torch.zeros_like(st)
                ops.conv3_forward(d_op[0], img_t, g, a_lo=lo, slope=slope, eps=eps, mask_x=x, mask_stats=st, bwd_sums=sums, pointwise=pointwise)
                ops.instnorm_backward_apply(g, x, st, sums, dx, eps=eps)
            else:
                ops.conv3_forward(d_op[0], img_t, dx, a_lo=lo, pointwise=pointwise)
        return dx, None, dw, (dy if has_res else None), None, None, None, None, None, None, None, None


class _DwConv(torch.autograd.Function):
    """nn.Conv3d(C, C, 3, padding=1, groups=C, bias=False)."""

    @staticmethod
    def forward(ctx, a, w):
        ctx.save_for_backward(a, w)
        ctx.param = w
        return ops.dwconv3(a, w)

    @staticmethod
    def backward(ctx, dy):
        a, w = ctx.saved_tensors
        dy = dy.contiguous()
        dw = None
        if ctx.needs_input_grad[1]:
            if _Side.usable(ctx.param):
                _Side.run(ctx.param, lambda: ops.dwconv3_wgrad(a, dy), a, dy)
            else:
                dw = ops.dwconv3_wgrad(a, dy).view_as(w)
        da = ops.dwconv3(dy, w, flip=True) if ctx.needs_input_grad[0] else None
        return da, dw


class _Stem(torch.autograd.Function):
    """inc.conv1: Conv3d(1, base, 3, padding=1, bias=False) on the fp32 NCDHW image -> NDHWC."""

    @staticmethod
    def forward(ctx, x, w, dtype):
        n, _, d, h, w_ = x.shape
        y = torch.empty((n, d, h, w_, w.shape[0]), dtype=dtype, device=x.device)
        ops.stem_conv_forward(x, w, y)
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dw = torch.empty_like(w)
        ops.stem_conv_wgrad(x, dy.contiguous(), dw)
        return None, dw, None


class _Head(torch.autograd.Function):
    """outc: Conv3d(C, classes, 1) with bias, NDHWC -> fp32 NCDHW logits."""

    @staticmethod
    def forward(ctx, x, w, b):
        n, d, h, w_, _ = x.shape
        logits = torch.empty((n, w.shape[0], d, h, w_), dtype=torch.float32, device=x.device)
        ops.head_forward(x, w, b, logits)
        ctx.save_for_backward(x, w)
        return logits

    @staticmethod
    def backward(ctx, dl):
        x, w = ctx.saved_tensors
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        db = torch.empty(w.shape[0], dtype=torch.float32, device=x.device)
        ops.head_backward(x, w, dl.contiguous(), dx, dw, db)
        return dx, dw, db


class _SpaceToDepth(torch.autograd.Function):
    """PatchMerging's torch.cat of the eight stride-2 sub-grids (medformer_utils.py:170-181): channel (4i + 2j + k) * C + c."""

    @staticmethod
    def forward(ctx, x):
        n, d, h, w_, c = x.shape
        q = torch.empty((n, d // 2, h // 2, w_ // 2, 8 * c), dtype=x.dtype, device=x.device)
        ops.space_to_depth2(x, q)
        return q

    @staticmethod
    def backward(ctx, dq):
        dq = dq.contiguous()
        n, d, h, w_, c8 = dq.shape
        dx = torch.empty((n, 2 * d, 2 * h, 2 * w_, c8 // 8), dtype=dq.dtype, device=dq.device)
        ops.depth_to_space2(dq, None, dx)
        return dx


class _Up2(torch.autograd.Function):
    """F.interpolate(size = 2x, mode='trilinear', align_corners=True) (medformer_utils.py:360)."""

    @staticmethod
    def forward(ctx, x, size):
        n, _, _, _, c = x.shape
        y = torch.empty((n,) + tuple(size) + (c,), dtype=x.dtype, device=x.device)
        y_st = torch.zeros((n, c, 2), dtype=torch.float32, device=x.device)      # (sum, sumsq) of y from the kernel's epilogue
        ops.upsample_forward(x, y, y_st)
        ctx.in_shape = x.shape
        ctx.mark_non_differentiable(y_st)
        return y, y_st

    @staticmethod
    def backward(ctx, dy, _unused):
        dy = dy.contiguous()
        dx = torch.empty(ctx.in_shape, dtype=dy.dtype, device=dy.device)
        ops.upsample_backward(dy, dx)
        return dx, None


class _ChannelMean(torch.autograd.Function):
    """SEBlock.squeeze: mean over the voxels -> fp32 [N, C]."""

    @staticmethod
    def forward(ctx, h):
        ctx.shape, ctx.dtype = h.shape, h.dtype
        v = h.shape[1] * h.shape[2] * h.shape[3]
        st = ops.channel_stats(h)                       # (sum, sumsq): the caller derives the statistics of h * gate from them
        ctx.mark_non_differentiable(st)
        return st[..., 0].contiguous() / v, st

    @staticmethod
    def backward(ctx, dm, _unused):
        n, d, h, w_, c = ctx.shape
        return (dm / (d * h * w_)).to(ctx.dtype).view(n, 1, 1, 1, c).expand(ctx.shape)


class _Scale(torch.autograd.Function):
    """SEBlock: x * excitation, per (sample, channel)."""

    @staticmethod
    def forward(ctx, h, s):
        s = s.contiguous()
        ctx.save_for_backward(h, s)
        return ops.scale_channels(h, s)

    @staticmethod
    def backward(ctx, dy):
        h, s = ctx.saved_tensors
        dy = dy.contiguous()
        return ops.scale_channels(dy, s), ops.channel_dot(dy, h)


class _SoftmaxPool(torch.autograd.Function):
    """SemanticMapGeneration (medformer_utils.py:229-234): softmax of the code logits over the voxels, then
    einsum('bij,bkj->bik', feat, weights) -> fp32 [N, C, codes]."""

    @staticmethod
    def forward(ctx, feat, logit, codes):
        smap, ms = ops.softmax_pool_forward(feat, logit, codes)
        ctx.save_for_backward(feat, logit, ms, smap)
        return smap

    @staticmethod
    def backward(ctx, ds):
        feat, logit, ms, smap = ctx.saved_tensors
        ds = ds.contiguous().float()
        tk = (ds * smap).sum(1).contiguous()
        dfeat, dlogit = ops.softmax_pool_backward(feat, logit, ms, ds, tk)
        return dfeat, dlogit, None


class _BiAttention(torch.autograd.Function):
    """BidirectionAttention's attention core (medformer_utils.py:84-97): qv = voxel-side [query | value] channels, mq / mv the
    map-side query / value as fp32 [N, heads, 27, dim_head] -> (voxel-side output, map-side output [N, heads, 27, dim_head])."""

    @staticmethod
    def forward(ctx, qv, mq, mv, heads):
        mq, mv = mq.contiguous(), mv.contiguous()
        fo, mo, ms = ops.biattention_forward(qv, mq, mv, heads)
        ctx.save_for_backward(qv, mq, mv, ms, mo)
        ctx.heads = heads
        return fo, mo

    @staticmethod
    def backward(ctx, dfo, dmo):
        qv, mq, mv, ms, mo = ctx.saved_tensors
        dfo = torch.zeros_like(qv[..., :qv.shape[4] // 2]).contiguous() if dfo is None else dfo.contiguous()
        dmo = torch.zeros_like(mo) if dmo is None else dmo.contiguous().float()
        tj = (dmo * mo).sum(-1).contiguous()
        dqv, dmq, dmv = ops.biattention_backward(qv, mq, mv, ms, dfo, dmo, tj, ctx.heads)
        return dqv, dmq, dmv, None


# ------------------------------------------------------------------------------------------------
# parameter table
# ------------------------------------------------------------------------------------------------
def medformer_param_shapes(in_chan, num_classes, base_chan, chan_num, conv_num, trans_num, num_heads, fusion_dim, fusion_depth,
                           expansion, aux_loss) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every parameter, named like the reference module tree (medformer.py:116-146, medformer_utils.py)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, co, ci, k):
        out.append((name, (co, ci, k, k, k)))

    def basic_block(pre, ci, co):
        conv(pre + "conv1.conv.weight", co, ci, 3)
        conv(pre + "conv2.conv.weight", co, co, 3)
        if ci != co:
            conv(pre + "shortcut.conv.weight", co, ci, 3)

    def dsconv(pre, ci, co):
        conv(pre + "depthwise.weight", ci, 1, 3)
        conv(pre + "pointwise.weight", co, ci, 1)

    def attn_block(pre, ci, cm, co, no_map_out):
        dsconv(pre + "attn.feat_qv.", ci, 2 * co)
        dsconv(pre + "attn.feat_out.", co, co)
        conv(pre + "attn.map_qv.weight", 2 * co, cm, 1)
        if not no_map_out:
            conv(pre + "attn.map_out.weight", cm, co, 1)
        if ci != co:
            conv(pre + "shortcut.conv.weight", co, ci, 1)
        ff = pre + "feedforward."
        e = co * expansion
        conv(ff + "expand_proj.conv.weight", e, co, 1)
        conv(ff + "depthwise.conv.weight", e, 1, 3)
        conv(ff + "se.excitation.0.weight", e // 4, e, 1)
        out.append((ff + "se.excitation.0.bias", (e // 4,)))
        conv(ff + "se.excitation.2.weight", e, e // 4, 1)
        out.append((ff + "se.excitation.2.bias", (e,)))
        conv(ff + "pointwise.conv.weight", co, e, 1)

    def layer(pre, ci, cm, co, nblk, no_map_out=False):
        for i in range(nblk):
            attn_block(f"{pre}blocks.{i}.", ci if i == 0 else co, cm, co, no_map_out and i == nblk - 1)

    conv("inc.conv1.weight", base_chan, in_chan, 3)
    basic_block("inc.conv2.", base_chan, base_chan)
    prev = base_chan
    for i in range(4):
        pre, co = f"down{i + 1}.", chan_num[i]
        if i >= 1:
            conv(pre + "map_gen.base_proj.weight", co, co, 3)
            conv(pre + "map_gen.semantic_proj.weight", 27, co, 3)
        dsconv(pre + "patch_merging.reduction.", 8 * prev, co)
        for j in range(conv_num[i]):
            basic_block(f"{pre}conv_blocks.{j}.", co, co)
        layer(pre + "trans_blocks.", co, co, co, trans_num[i])
        prev = co
    for i, c in enumerate(chan_num[1:4]):
        conv(f"map_fusion.in_proj.{i}.weight", fusion_dim, c, 1)
    for l in range(fusion_depth):
        p = f"map_fusion.fusion.layers.{l}."
        out.extend([(p + "0.norm.weight", (fusion_dim,)), (p + "0.norm.bias", (fusion_dim,)),
                    (p + "0.fn.to_qkv.weight", (3 * fusion_dim, fusion_dim)), (p + "0.fn.to_out.weight", (fusion_dim, fusion_dim)),
                    (p + "0.fn.to_out.bias", (fusion_dim,)), (p + "1.norm.weight", (fusion_dim,)), (p + "1.norm.bias", (fusion_dim,)),
                    (p + "1.fn.fc1.weight", (fusion_dim, fusion_dim)), (p + "1.fn.fc1.bias", (fusion_dim,)),
                    (p + "1.fn.fc2.weight", (fusion_dim, fusion_dim)), (p + "1.fn.fc2.bias", (fusion_dim,))])
    for i, c in enumerate(chan_num[1:4]):
        conv(f"map_fusion.out_proj.{i}.weight", c, fusion_dim, 1)
    skips = [chan_num[2], chan_num[1], chan_num[0], base_chan]
    for i in range(4):
        pre, ci, co = f"up{i + 1}.", chan_num[3 + i], chan_num[4 + i]
        cat = ci + skips[i]
        if trans_num[4 + i] > 0:
            conv(pre + "map_reduction.weight", co, ci + co, 1)
            layer(pre + "trans_blocks.", cat, co, co, trans_num[4 + i], no_map_out=(i == 1))
            for j in range(conv_num[4 + i]):
                basic_block(f"{pre}conv_blocks.{j}.", co, co)
        else:
            for j in range(conv_num[4 + i]):
                basic_block(f"{pre}conv_blocks.{j}.", cat if j == 0 else co, co)
    if aux_loss:
        conv("aux_out.weight", num_classes, chan_num[5], 1)
        out.append(("aux_out.bias", (num_classes,)))
    conv("outc.weight", num_classes, chan_num[7], 1)
    out.append(("outc.bias", (num_classes,)))
    return out


def _register(root: nn.Module, dotted: str, p: nn.Parameter):
    mod = root
    parts = dotted.split(".")
    for name in parts[:-1]:
        nxt = getattr(mod, name, None) if name in mod._modules else None
        if nxt is None:
            nxt = nn.Module()
            mod.add_module(name, nxt)
        mod = nxt
    mod.register_parameter(parts[-1], p)


class B200MedFormer(nn.Module):
    """Drop-in for `MedFormer(in_chan, num_classes, ...)` (medformer.py:83-113) in the yaml configuration (BasicBlock conv blocks,
    InstanceNorm, ReLU, depthwise projections, 3x3x3 kernels, x2 scales, map_size 3x3x3).  precision 'bf16' | 'fp32' as in B200UNet."""

    def __init__(self, in_chan, num_classes, base_chan=32, map_size=(3, 3, 3), conv_block="BasicBlock", conv_num=(2, 1, 0, 0, 0, 1, 2, 2),
                 trans_num=(0, 1, 2, 2, 2, 1, 0, 0), chan_num=(64, 128, 256, 320, 256, 128, 64, 32), num_heads=(1, 4, 8, 16, 8, 4, 1, 1),
                 fusion_depth=2, fusion_dim=320, fusion_heads=4, expansion=4, attn_drop=0.0, proj_drop=0.0, proj_type="depthwise", norm="in",
                 act="relu", kernel_size=None, scale=None, aux_loss=False, precision: str = "bf16"):
        super().__init__()
        if in_chan != 1:
            raise NotImplementedError("B200MedFormer: in_chan must be 1 (the stem kernel, like the reference's CT pipeline)")
        if conv_block != "BasicBlock" or norm != "in" or act != "relu" or proj_type != "depthwise":
            raise NotImplementedError("B200MedFormer implements the yaml configuration: BasicBlock / InstanceNorm / ReLU / depthwise projections")
        if attn_drop or proj_drop:
            raise NotImplementedError("B200MedFormer: dropout is 0 in the reference configuration")
        if tuple(map_size) != (3, 3, 3):
            raise NotImplementedError("B200MedFormer: map_size must be [3, 3, 3] (27 map tokens)")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        for i in range(8):
            if trans_num[i] and chan_num[i] // num_heads[i] > 32:
                raise NotImplementedError("B200MedFormer: dim_head must be <= 32 where attention blocks are used")
        if trans_num[0] or trans_num[6] or trans_num[7]:
            raise NotImplementedError("B200MedFormer: attention blocks at down1 / up3 / up4 have no semantic map in the reference either")
        self.num_classes, self.base_chan, self.precision, self.aux_loss = num_classes, base_chan, precision, bool(aux_loss)
        self.cfg = dict(chan_num=list(chan_num), conv_num=list(conv_num), trans_num=list(trans_num), num_heads=list(num_heads),
                        fusion_depth=fusion_depth, fusion_heads=fusion_heads, expansion=expansion)
        shapes = medformer_param_shapes(in_chan, num_classes, base_chan, list(chan_num), list(conv_num), list(trans_num), list(num_heads),
                                        fusion_dim, fusion_depth, expansion, aux_loss)
        for name, shp in shapes:
            t = torch.empty(shp)
            if name.endswith("norm.weight"):
                nn.init.ones_(t)
            elif name.endswith("norm.bias"):
                nn.init.zeros_(t)
            elif len(shp) > 1:
                nn.init.kaiming_uniform_(t, a=5 ** 0.5)                 # nn.Conv3d / nn.Linear default
            else:
                fan_in = dict(shapes)[name[:-4] + "weight"]
                bound = 1.0 / (int(torch.tensor(fan_in[1:]).prod()) ** 0.5)
                nn.init.uniform_(t, -bound, bound)
            _register(self, name, nn.Parameter(t))
        self._plan = None
        self._plan_key = None
        self._pad = {}

    def __getstate__(self):
        # copy.deepcopy (EMA copies, training/utils.py:154-161) / pickling: the pack plan holds raw pointers of THIS module's
        # parameters and is rebuilt on the first forward of the copy
        d = self.__dict__.copy()
        d["_plan"], d["_plan_key"], d["_pad"] = None, None, {}
        d.pop("_P", None)
        return d

    # ---- packed tensor-core weight images: one persistent plan, refreshed by one launch per forward --------------------------
    @staticmethod
    def _is_tensor_conv(name: str, p: torch.Tensor) -> bool:
        if p.dim() != 5 or name in ("inc.conv1.weight", "outc.weight", "aux_out.weight"):
            return False
        if ".depthwise." in name or "map_" in name.split(".")[-2] or name.startswith("map_fusion.") or ".se." in name:
            return False
        return True

    def _prepare(self, P: Dict[str, torch.Tensor]):
        names = [n for n, p in P.items() if self._is_tensor_conv(n, p)]
        key = tuple(P[n].data_ptr() for n in names)
        if self._plan is None or key != self._plan_key:
            jobs = []
            for n in names:
                w = P[n].detach()
                pad = None
                if w.shape[0] % 8:
                    pad = self._pad.get(n)
                    rows = (w.shape[0] + 7) // 8 * 8 - w.shape[0]
                    if pad is None or pad.device != w.device:
                        pad = self._pad[n] = torch.zeros((rows,) + tuple(w.shape[1:]), dtype=torch.float32, device=w.device)
                jobs.append((n, w, pad, False))
                jobs.append((n + "T", w, pad, True))
            self._plan = ops.PackPlan(jobs, split=self.precision == "fp32")
            self._plan_key = key
        self._plan.refresh()

    # ---- blocks (oracle/medformer_ref.py is the line-by-line statement of the same graph) ------------------------------------
    def _cna(self, x, key, *, norm=True, act=True, res=None, eps=EPS_CNA, stats=True):
        """ConvNormAct on the tensor cores.  The conv epilogue also leaves the (sum, sumsq) InstanceNorm statistics of its output
        (attribute `rsb_stats` of the returned tensor), which the next normalisation of that tensor picks up instead of
        running a statistics pass."""
        w = self._P[key]
        cout = (w.shape[0] + 7) // 8 * 8
        y, y_st = _ConvNA.apply(x, getattr(x, "rsb_stats", None) if norm else None, w, res, self._plan.images[key],
                                self._plan.images[key + "T"], norm, 0.0 if act else 1.0, eps, w.shape[-1] == 1, cout, stats)
        if stats:
            y.rsb_stats = y_st
        return y

    @staticmethod
    def _norm_act(x, eps, slope):
        return _NormAct.apply(x, getattr(x, "rsb_stats", None), eps, slope)

    def _basic_block(self, x, pre):
        key = pre + "shortcut.conv.weight"
        sc = self._cna(x, key) if key in self._P else x
        return self._cna(self._cna(x, pre + "conv1.conv.weight"), pre + "conv2.conv.weight", res=sc)

    def _dsconv(self, x, pre, res=None, stats=True):
        h = _DwConv.apply(x, self._P[pre + "depthwise.weight"])
        return self._cna(h, pre + "pointwise.weight", norm=False, act=False, res=res, stats=stats)

    def _mbconv(self, x, pre):
        P = self._P
        h = self._cna(x, pre + "expand_proj.conv.weight")
        h = _DwConv.apply(self._norm_act(h, EPS_CNA, 0.0), P[pre + "depthwise.conv.weight"])
        s, st_h = _ChannelMean.apply(h)                                                # SEBlock (conv_layers.py:159-173)
        w0, w2 = P[pre + "se.excitation.0.weight"], P[pre + "se.excitation.2.weight"]
        s = F.relu(F.linear(s, w0.flatten(1), P[pre + "se.excitation.0.bias"]))
        s = torch.sigmoid(F.linear(s, w2.flatten(1), P[pre + "se.excitation.2.bias"]))
        h = _Scale.apply(h, s)
        g = s.detach().contiguous()
        h.rsb_stats = torch.stack([st_h[..., 0] * g, st_h[..., 1] * g * g], dim=-1).contiguous()   # sums of h * gate: no statistics pass
        return self._cna(h, pre + "pointwise.conv.weight", act=False, res=x)

    def _attention_block(self, x, smap, pre, heads):
        P = self._P
        n, d, h, w_, _ = x.shape
        xn = self._norm_act(x, EPS_DEF, 1.0)
        mn = F.instance_norm(smap, eps=EPS_DEF)
        qv = self._dsconv(xn, pre + "attn.feat_qv.", stats=False)
        c = qv.shape[4] // 2
        dh = c // heads
        mqv = F.conv3d(mn, P[pre + "attn.map_qv.weight"])                             # [N, 2C, 3, 3, 3]
        mq, mv = (t.reshape(n, dh, heads, 27).permute(0, 2, 3, 1).contiguous() for t in mqv.chunk(2, dim=1))
        fo, mo = _BiAttention.apply(qv, mq, mv, heads)
        key = pre + "shortcut.conv.weight"
        sc = self._cna(x, key) if key in P else x
        out = self._dsconv(fo, pre + "attn.feat_out.", res=sc)
        mo = mo.permute(0, 3, 1, 2).reshape(n, dh * heads, 3, 3, 3)
        key = pre + "attn.map_out.weight"
        if key in P:
            mo = F.conv3d(mo, P[key])
        return self._mbconv(out, pre + "feedforward."), mo + smap

    def _layer(self, x, smap, pre, nblk, heads):
        for i in range(nblk):
            x, smap = self._attention_block(x, smap, f"{pre}blocks.{i}.", heads)
        return x, smap

    def _map_generation(self, x, pre):
        feat = self._cna(x, pre + "base_proj.weight", norm=False, act=False, stats=False)
        logit = self._cna(x, pre + "semantic_proj.weight", norm=False, act=False, stats=False)   # 27 codes in a 32-channel tensor
        smap = _SoftmaxPool.apply(feat, logit, 27)
        return smap.reshape(x.shape[0], feat.shape[4], 3, 3, 3)

    def _down(self, x, i, map_generate):
        pre = f"down{i}."
        c = self.cfg
        x = self._dsconv(self._norm_act(_SpaceToDepth.apply(x), EPS_DEF, 1.0), pre + "patch_merging.reduction.")
        for j in range(c["conv_num"][i - 1]):
            x = self._basic_block(x, f"{pre}conv_blocks.{j}.")
        smap = self._map_generation(x, pre + "map_gen.") if map_generate else None
        return self._layer(x, smap, pre + "trans_blocks.", c["trans_num"][i - 1], c["num_heads"][i - 1])

    def _up(self, x1, x2, map1, map2, i):
        pre = f"up{i}."
        c = self.cfg
        up, up_st = _Up2.apply(x1, tuple(x2.shape[1:4]))
        feat = torch.cat([up, x2], dim=4)
        if getattr(x2, "rsb_stats", None) is not None:
            feat.rsb_stats = torch.cat([up_st, x2.rsb_stats], dim=1).contiguous()      # both halves already carry their sums
        key = pre + "map_reduction.weight"
        smap = F.conv3d(torch.cat([map1, map2], dim=1), self._P[key]) if (key in self._P and map2 is not None) else map1
        out, smap = self._layer(feat, smap, pre + "trans_blocks.", c["trans_num"][3 + i], c["num_heads"][3 + i])
        for j in range(c["conv_num"][3 + i]):
            out = self._basic_block(out, f"{pre}conv_blocks.{j}.")
        return out, smap

    def _transformer(self, x, pre, depth, heads):
        """TransformerBlock on the <= 81 map tokens (trans_layers.py:104-118): plain torch."""
        P = self._P
        dim = x.shape[-1]
        for i in range(depth):
            p = f"{pre}layers.{i}.0."
            h = F.layer_norm(x, (dim,), P[p + "norm.weight"], P[p + "norm.bias"])
            q, k, v = F.linear(h, P[p + "fn.to_qkv.weight"]).chunk(3, dim=-1)
            b, l, n = q.shape
            q, k, v = (t.reshape(b, l, heads, n // heads).permute(0, 2, 1, 3) for t in (q, k, v))
            attn = F.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * (n // heads) ** (-0.5), dim=-1)
            a = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, l, n)
            x = F.linear(a, P[p + "fn.to_out.weight"], P[p + "fn.to_out.bias"]) + x
            p = f"{pre}layers.{i}.1."
            h = F.layer_norm(x, (dim,), P[p + "norm.weight"], P[p + "norm.bias"])
            x = F.linear(F.gelu(F.linear(h, P[p + "fn.fc1.weight"], P[p + "fn.fc1.bias"])), P[p + "fn.fc2.weight"], P[p + "fn.fc2.bias"]) + x
        return x

    def _map_fusion(self, maps):
        P = self._P
        b = maps[0].shape[0]
        proj = [F.conv3d(m, P[f"map_fusion.in_proj.{i}.weight"]) for i, m in enumerate(maps)]
        dim = proj[0].shape[1]
        tokens = torch.cat([p.reshape(b, dim, -1).permute(0, 2, 1) for p in proj], dim=1)
        tokens = self._transformer(tokens, "map_fusion.fusion.", self.cfg["fusion_depth"], self.cfg["fusion_heads"])
        outs = tokens.chunk(len(maps), dim=1)
        return [F.conv3d(o.permute(0, 2, 1).reshape(b, dim, 3, 3, 3), P[f"map_fusion.out_proj.{i}.weight"]) for i, o in enumerate(outs)]

    def forward(self, x):
        if not ops._on_device(x):
            raise RuntimeError("B200MedFormer has no CPU path: the input must live on a CUDA (sm_100a) device")
        if x.dim() != 5 or x.shape[1] != 1 or any(s % 16 for s in x.shape[2:]):
            raise ValueError("B200MedFormer: input must be [N, 1, D, H, W] with D, H, W multiples of 16")
        dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        self._P = P = dict(self.named_parameters())
        self._prepare(P)
        x = x.float().contiguous()
        x0 = self._basic_block(_Stem.apply(x, P["inc.conv1.weight"], dtype), "inc.conv2.")
        x1, _ = self._down(x0, 1, False)
        x2, m2 = self._down(x1, 2, True)
        x3, m3 = self._down(x2, 3, True)
        x4, m4 = self._down(x3, 4, True)
        maps = self._map_fusion([m2, m3, m4])
        out, smap = self._up(x4, x3, maps[2], maps[1], 1)
        out, smap = self._up(out, x2, smap, maps[0], 2)
        aux = None
        if self.aux_loss:
            # deep-supervision head at 1/4 resolution (medformer.py:205-208): a [N, classes] 1x1x1 conv + trilinear x4 in torch
            a = F.conv3d(out.permute(0, 4, 1, 2, 3).float(), P["aux_out.weight"], P["aux_out.bias"])
            aux = F.interpolate(a, size=x.shape[-3:], mode="trilinear", align_corners=True)
        out, smap = self._up(out, x1, smap, None, 3)
        out, smap = self._up(out, x0, smap, None, 4)
        logits = _Head.apply(out, P["outc.weight"], P["outc.bias"])
        return {"segmentation": [logits, aux] if self.aux_loss else logits}
