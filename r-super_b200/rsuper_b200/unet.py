"""B200UNet — drop-in for the reference 3D UNet (rsuper_train/model/dim3/unet.py:12-64).

Same constructor arguments, same parameter names / shapes / init (the state dict is
interchangeable with the reference's `UNet(..., block='BasicBlock', norm='in')`), but forward and
backward run entirely on the hand-written sm_100a kernels of librsuper_b200.so:

  reference op (file:line)                                   kernel
  ---------------------------------------------------------  ---------------------------------
  inconv.conv1  Conv3d(1,b,3)          unet_utils.py:15,18    rsb_stem_conv_forward / _wgrad
  ConvNormAct preact IN->ReLU          conv_layers.py:47-49   rsb_norm_act (bf16 conv operand, written once)
  ConvNormAct Conv3d(k3,p1,no bias)    conv_layers.py:29-38   rsb_conv3_forward (TMA-fed tcgen05 implicit GEMM)
  BasicBlock `out += shortcut(x)`      conv_layers.py:92      rsb_conv3_forward epilogue (res)
  BasicBlock conv1 || shortcut conv    conv_layers.py:85-92   ONE GEMM with N = 2*Cout (same
                                                              IN(x): affine-free norms coincide)
  nn.MaxPool3d                         unet_utils.py:36       rsb_maxpool2_forward / _backward
  F.interpolate trilinear              unet_utils.py:69       rsb_upsample_trilinear_*
  torch.cat([skip, up], 1)             unet_utils.py:71       free: producers write channel slices
  outc Conv3d(b,C,1)                   unet.py:47,62          rsb_head_forward / _backward
  autograd of all of the above                                dgrad (same tcgen05 kernel, flipped
                                                              weights), rsb_conv3_wgrad,
                                                              rsb_instnorm_backward_apply

`negative_slope` (0 => the reference's ReLU, 0.01 => LeakyReLU) is the north-star variant knob.
There is no fallback path: without the CUDA library every call raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import os

import torch
import torch.nn as nn

from . import ops


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's module tree (names only; forward is never called)
# ------------------------------------------------------------------------------------------------
class _CNA(nn.Module):
    """ConvNormAct(preact=True, norm=InstanceNorm3d) — only the conv carries parameters."""

    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel_size=k, padding=k // 2, bias=False)


class _Bottleneck(nn.Module):
    """Bottleneck (conv_layers.py:97-123): 1x1x1 (C -> C/2), 3x3x3, 1x1x1 (C/2 -> C), 3x3x3 shortcut when cin != cout."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _CNA(cin, cout // 2, 1)
        self.conv2 = _CNA(cout // 2, cout // 2)
        self.conv3 = _CNA(cout // 2, cout, 1)
        self.shortcut = nn.Sequential()
        if cin != cout:
            self.shortcut = _CNA(cin, cout)


class _BasicBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _CNA(cin, cout)
        self.conv2 = _CNA(cout, cout)
        self.shortcut = nn.Sequential()
        if cin != cout:
            self.shortcut = _CNA(cin, cout)

    def weights(self):
        sc = self.shortcut.conv.weight if isinstance(self.shortcut, _CNA) else None
        return self.conv1.conv.weight, self.conv2.conv.weight, sc


class _SingleConv(nn.Module):
    """SingleConv (conv_layers.py:56-68): .conv = ConvNormAct(preact=False) whose .conv is the only parameter."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _CNA(cin, cout)


class _InConv(nn.Module):
    def __init__(self, cin, cout, block=None):
        super().__init__()
        self.conv1 = nn.Conv3d(cin, cout, kernel_size=3, padding=1, bias=False)
        self.conv2 = (block or _BasicBlock)(cout, cout)


class _Stage(nn.Module):
    """down_block (index 0 is the parameter-free MaxPool3d) / up_block."""

    def __init__(self, cin, cout, down: bool, block=None, up_channels: int = 0):
        super().__init__()
        block = block or _BasicBlock
        if up_channels:
            # up_mode='transposed': ConvTranspose3d(kernel = stride = 2) replaces F.interpolate (semantics of vnet.py:108)
            self.up = nn.ConvTranspose3d(up_channels, up_channels, kernel_size=2, stride=2)
        blocks: List[nn.Module] = []
        if down:
            blocks.append(nn.Identity())  # placeholder for nn.MaxPool3d: keeps the index -> name map
        blocks.append(block(cin, cout))
        blocks.append(block(cout, cout))
        self.conv = nn.Sequential(*blocks)

    def blocks(self):
        return [m for m in self.conv if isinstance(m, _BasicBlock)]


# ------------------------------------------------------------------------------------------------
# activation buffers: NDHWC data + statistics sharing the channel pitch
# ------------------------------------------------------------------------------------------------
class Act:
    __slots__ = ("t", "st")

    def __init__(self, t: torch.Tensor, st: Optional[torch.Tensor]):
        self.t = t
        self.st = st

    @staticmethod
    def new(n, d, h, w, c, dtype, device, stats=True) -> "Act":
        t = torch.empty((n, d, h, w, c), dtype=dtype, device=device)
        st = torch.zeros((n, c, 2), dtype=torch.float32, device=device) if stats else None
        return Act(t, st)

    def view(self, c0, c1) -> "Act":
        return Act(self.t[..., c0:c1], None if self.st is None else self.st[:, c0:c1])

    @property
    def C(self):
        return self.t.shape[4]


def _sums_like(a: Act) -> torch.Tensor:
    """Zeroed (S1,S2) buffer indexed like a's statistics (same pitch, same channel offset)."""
    pitch = a.t.stride(3)
    n = a.t.shape[0]
    full = torch.zeros((n, pitch, 2), dtype=torch.float32, device=a.t.device)
    c0 = (a.t.storage_offset() % pitch) if pitch > 0 else 0
    return full[:, c0:c0 + a.C]


# Opt-in (RSB_SIDE_STREAM=1 / set_side_stream(True)).  Measured on B200: 21.8-22.3 ms/step against 22.7-23 ms serialised when
# the host enqueues far ahead of the GPU, but 27-30 ms on boxes whose host cores were slower or shared (the enqueue order
# of the two streams then lets a persistent wgrad grab the SMs ahead of the dgrad on the critical path): the serialised
# schedule is the robust default.
SIDE_STREAM = os.environ.get("RSB_SIDE_STREAM", "0") == "1"


def set_side_stream(enabled: bool) -> bool:
    """Run weight gradients / weight packing on a second stream, or serialise everything on the caller's stream
    (default; also what per-kernel timing with CUDA events needs).  Returns the previous setting."""
    global SIDE_STREAM
    prev, SIDE_STREAM = SIDE_STREAM, bool(enabled)
    return prev


class _Engine:
    """Kernel schedule of one forward / backward pass (pure host logic; all compute is in the .so).

    precision 'bf16': bf16 activation storage, bf16 tensor-core operands (one MMA pass), fp32 accumulation.
    precision 'fp32': fp32 activation storage; every tensor-core product is the 3-pass split
                      a_hi*w_hi + a_lo*w_hi + a_hi*w_lo of bf16 halves (~2^-17 per operand, fp32 accumulation) — the
                      parity mode measured against the fp32 reference.  (The kernels also take a third piece / six
                      products, but the tensor pipe's truncating fp32 accumulation already bounds a conv at ~1e-5 of
                      its output range, so the third piece buys nothing — tests/test_kernels_gpu.py.)"""

    def __init__(self, base_ch: int, slope: float, dtype: torch.dtype, block: str = "BasicBlock", up_mode: str = "trilinear"):
        self.block = block
        self.up_mode = up_mode
        self._upw = {}      # up_mode='transposed': persistent [8C, C, 1, 1, 1] images of the ConvTranspose3d weights
        self.b = base_ch
        self.slope = float(slope)
        self.dtype = dtype
        self.split = dtype == torch.float32
        self._zp, self._zp_off = None, 0
        # Optional second stream (see SIDE_STREAM): the weight gradients (tensor-bound, 52 registers x 192 threads per SM)
        # run beside the main backward chain, whose HBM-bound InstanceNorm-backward / pooling / upsampling passes fit on
        # the same SMs, and their persistent CTAs fill the tails of the dgrad launches; the weight packing overlaps the stem.
        self._side = None
        self._side_keep = []
        self._plan_key = None
        # Optional gradient sink (train_step.B200TrainStep): name -> preallocated fp32 tensor the weight gradient is written
        # INTO (views of one flat buffer); `stage_done(k)` is called when a bucket of them has been enqueued
        self.sink = None
        self.plan = None  # ops.PackPlan: persistent packed weight images, refreshed by ONE launch per forward

    # ---- zeroed statistics storage: one fill per pass instead of ~50 tiny ones --------------------------
    def _zeros(self, numel: int, device) -> torch.Tensor:
        zp = self._zp
        if zp is None or self._zp_off + numel > zp.numel():
            zp = self._zp = torch.zeros(max(1 << 16, numel), dtype=torch.float32, device=device)
            self._zp_off = 0
        out = zp[self._zp_off:self._zp_off + numel]
        self._zp_off += (numel + 3) // 4 * 4
        return out

    def _new_act(self, n, d, h, w, c, dtype, device, stats=True) -> Act:
        t = torch.empty((n, d, h, w, c), dtype=dtype, device=device)
        return Act(t, self._zeros(n * c * 2, device).view(n, c, 2) if stats else None)

    def _sums_like(self, a: Act) -> torch.Tensor:
        """Zeroed (S1,S2) buffer indexed like a's statistics (same pitch, same channel offset)."""
        pitch = a.t.stride(3)
        n = a.t.shape[0]
        full = self._zeros(n * pitch * 2, a.t.device).view(n, pitch, 2)
        c0 = (a.t.storage_offset() % pitch) if pitch > 0 else 0
        return full[:, c0:c0 + a.C]

    # ---- packed weights ----------------------------------------------------------------------------
    BLOCKS = ([("inc.conv2.", False)]
              + [(f"down{l}.conv.{i}.", i == 1) for l in range(1, 5) for i in (1, 2)]
              + [(f"up{j}.conv.{i}.", i == 0) for j in range(1, 5) for i in (0, 1)])

    SINGLE_CONVS = (["inc.conv2."] + [f"down{l}.conv.{i}." for l in range(1, 5) for i in (1, 2)]
                    + [f"up{j}.conv.{i}." for j in range(1, 5) for i in (0, 1)])

    def _up_weight_images(self, P: dict):
        """up_mode='transposed': W'[(4a + 2b + c) * C + co][ci] = w[ci][co][a][b][c] — the ConvTranspose3d(k = s = 2) weight as
        the [8C, C, 1, 1, 1] kernel of the equivalent 1x1x1 conv, kept in persistent buffers (the pack plan holds their
        pointers) and refreshed from the parameters every forward."""
        out = []
        if self.up_mode != "transposed":
            return out
        for j in range(1, 5):
            w = P[f"up{j}.up.weight"]
            c = w.shape[0]
            buf = self._upw.get(j)
            if buf is None or buf.device != w.device or buf.shape[1] != c:
                buf = self._upw[j] = torch.empty((8 * c, c, 1, 1, 1), dtype=torch.float32, device=w.device)
            buf.view(2, 2, 2, c, c).copy_(w.permute(2, 3, 4, 1, 0))
            out.append((f"up{j}.upw", buf))
        return out

    def prepare(self, P: dict):
        """(Re)build the pack plan when the parameter storage changed, then re-pack every conv (forward and dgrad
        images; conv1 || shortcut merged row-wise without a torch.cat) from the current values."""
        if self.block == "SingleConv":
            pairs = [(P[pre + "conv.conv.weight"], None) for pre in self.SINGLE_CONVS]
            key = ops.PackPlan.pointer_key(pairs)
            if self.plan is None or self._plan_key != key:
                jobs = []
                for pre, (wa, _) in zip(self.SINGLE_CONVS, pairs):
                    jobs.append((pre + "c", wa, None, False))
                    jobs.append((pre + "cT", wa, None, True))
                self.plan, self._plan_key = ops.PackPlan(jobs, split=self.split), key
            self.plan.refresh()
            return
        if self.block == "Bottleneck":
            names = [pre + t for pre, has_sc in self.BLOCKS for t in ("conv1", "conv2", "conv3") + (("shortcut",) if has_sc else ())]
            pairs = [(P[n + ".conv.weight"], None) for n in names]
            ups = self._up_weight_images(P)
            key = ops.PackPlan.pointer_key(pairs + [(t, None) for _, t in ups])
            if self.plan is None or self._plan_key != key:
                short = {"conv1": "c1", "conv2": "c2", "conv3": "c3", "shortcut": "sc"}
                jobs = []
                for n, (w, _) in zip(names, pairs):
                    pre, t = n.rsplit(".", 1)[0] + ".", n.rsplit(".", 1)[1]
                    jobs.append((pre + short[t], w, None, False))
                    jobs.append((pre + short[t] + "T", w, None, True))
                for k, t in ups:
                    jobs += [(k, t, None, False), (k + "T", t, None, True)]
                self.plan, self._plan_key = ops.PackPlan(jobs, split=self.split), key
            self.plan.refresh()
            return
        pairs = []
        for pre, has_sc in self.BLOCKS:
            pairs.append((P[pre + "conv1.conv.weight"], P[pre + "shortcut.conv.weight"] if has_sc else None))
            pairs.append((P[pre + "conv2.conv.weight"], None))
        # (the key is the parameter storage: one entry per conv — the plan itself holds two jobs per conv, fprop + dgrad image)
        ups = self._up_weight_images(P)
        key = ops.PackPlan.pointer_key(pairs + [(t, None) for _, t in ups])
        if self.plan is None or self._plan_key != key:
            jobs = []
            for (pre, _), k in zip(self.BLOCKS, range(0, len(pairs), 2)):
                for tag, (wa, wb) in (("c1", pairs[k]), ("c2", pairs[k + 1])):
                    jobs.append((pre + tag, wa, wb, False))
                    jobs.append((pre + tag + "T", wa, wb, True))
            for kk, t in ups:
                jobs += [(kk, t, None, False), (kk + "T", t, None, True)]
            self.plan, self._plan_key = ops.PackPlan(jobs, split=self.split), key
        self.plan.refresh()

    # ---- operand / conv / wgrad helpers ------------------------------------------------------------
    def _operand(self, t: torch.Tensor, stats=None):
        """bf16 conv operand pieces of a stored tensor: (hi,) in bf16 mode, (hi, lo) in fp32 mode;
        act(instnorm(t)) when stats are given, else t itself."""
        if stats is None and not self.split and t.dtype == torch.bfloat16:
            return (t,)
        r = ops.norm_act(t, stats, slope=self.slope, split=self.split)
        return r if self.split else (r,)

    def _conv(self, op, key: str, y, flip=False, **kw):
        wp = self.plan.images[key + "T" if flip else key]
        return ops.conv3_forward(op[0], wp, y, a_lo=op[1] if self.split else None, slope=self.slope, **kw)

    def _side_stream(self, device):
        if not SIDE_STREAM:
            return None
        if self._side is None or self._side.device != device:
            self._side = torch.cuda.Stream(device=device, priority=-1 if os.environ.get("RSB_SIDE_PRIORITY", "0") == "1" else 0)
        return self._side

    def _wgrad(self, a_op, dy_op, dw):
        side = self._side_stream(dw.device)
        if side is None:
            self._wgrad_launch(a_op, dy_op, dw)
            return dw
        side.wait_stream(torch.cuda.current_stream())   # operands were produced by work already enqueued on the main stream
        with torch.cuda.stream(side):
            self._wgrad_launch(a_op, dy_op, dw)
        # The operands must outlive the side-stream kernel: they are kept referenced until backward() has made the main
        # stream wait for the side stream (simpler for the caching allocator than tensor.record_stream, whose blocks
        # cannot be recycled while cross-stream uses are pending).
        self._side_keep.append((a_op, dy_op))
        return dw

    def _wgrad_launch(self, a_op, dy_op, dw):
        ops.conv3_wgrad(a_op[0], dy_op[0], dw)
        if self.split:
            ops.conv3_wgrad(a_op[1], dy_op[0], dw, accumulate=True)
            ops.conv3_wgrad(a_op[0], dy_op[1], dw, accumulate=True)

    # ---- up-sampling: trilinear (the reference, unet_utils.py:69) or ConvTranspose3d(k = s = 2) (up_mode='transposed') ----
    def _up_fwd(self, cur: Act, upv: Act, j: int, P: dict, up_ops: dict):
        if self.up_mode != "transposed":
            ops.upsample_forward(cur.t, upv.t, upv.st)
            return
        n, d, h, w_, c = cur.t.shape
        a_cur = self._operand(cur.t)          # the raw block output is the operand: no normalisation ahead of the transposed conv
        q = torch.empty((n, d, h, w_, 8 * c), dtype=self.dtype, device=cur.t.device)
        self._conv(a_cur, f"up{j}.upw", q, pointwise=True)
        ops.depth_to_space2(q, P[f"up{j}.up.bias"], upv.t, upv.st)
        up_ops[j] = a_cur

    def _up_bwd(self, d_up: torch.Tensor, j: int, S: dict, G: dict) -> torch.Tensor:
        """d_up = dL/d(up-sampled tensor) (a channel slice of d_cat) -> dL/d(low-resolution input); parameter gradients into G."""
        n, d2, h2, w2, c = d_up.shape
        d, h, w_ = d2 // 2, h2 // 2, w2 // 2
        d_cur = torch.empty((n, d, h, w_, c), dtype=self.dtype, device=d_up.device)
        if self.up_mode != "transposed":
            ops.upsample_backward(d_up, d_cur)
            return d_cur
        dq = torch.empty((n, d, h, w_, 8 * c), dtype=self.dtype, device=d_up.device)
        ops.space_to_depth2(d_up, dq)
        G[f"up{j}.up.bias"] = ops.channel_stats(d_up)[..., 0].sum(0)
        dq_op = self._operand(dq)
        self._conv(dq_op, f"up{j}.upw", d_cur, flip=True, pointwise=True)
        dwp = self._wgrad_pointwise(S["up_ops"][j], dq_op, 8 * c, c, dq)                  # [8C, C, 1, 1, 1]
        G[f"up{j}.up.weight"] = dwp.reshape(2, 2, 2, c, c).permute(4, 3, 0, 1, 2)        # -> [ci, co, a, b, c]
        return d_cur

    def _wgrad_pointwise(self, a_op, dy_op, cout: int, cin: int, like: torch.Tensor) -> torch.Tensor:
        """Weight gradient of a 1x1x1 conv: the centre tap of the 27-tap weight-gradient kernel's output (the other 26 taps are
        computed and dropped — a dedicated single-tap kernel is future work), returned as [cout, cin, 1, 1, 1]."""
        dw27 = self._wgrad(a_op, dy_op, self._dw(like, cout, cin))
        return dw27[:, :, 1:2, 1:2, 1:2]

    # ---- block='Bottleneck' (conv_layers.py:97-123) ------------------------------------------------
    def _bneck_fwd(self, x: Act, pre: str, has_sc: bool, cout: int, out: Act, saved: list):
        n, d, h, w_, _ = x.t.shape
        dev = x.t.device
        mid = cout // 2
        a_x = self._operand(x.t, x.st)
        h1 = self._new_act(n, d, h, w_, mid, self.dtype, dev)
        self._conv(a_x, pre + "c1", h1.t, out_stats=h1.st, pointwise=True)
        a_1 = self._operand(h1.t, h1.st)
        h2 = self._new_act(n, d, h, w_, mid, self.dtype, dev)
        self._conv(a_1, pre + "c2", h2.t, out_stats=h2.st)
        a_2 = self._operand(h2.t, h2.st)
        if has_sc:
            sc = torch.empty((n, d, h, w_, cout), dtype=self.dtype, device=dev)
            self._conv(a_x, pre + "sc", sc)
            res = sc
        else:
            res = x.t
        self._conv(a_2, pre + "c3", out.t, res=res, out_stats=out.st, pointwise=True)
        saved.append((x, h1, h2, a_x, a_1, a_2, has_sc))

    def _bneck_bwd(self, blk, pre: str, d_out: torch.Tensor, dx_dest: torch.Tensor, G: dict):
        """d_out = dL/d(block output) -> weight gradients into G, dL/d(block input) into dx_dest."""
        x, h1, h2, a_x, a_1, a_2, has_sc = blk
        cin, mid, cout = x.C, h1.C, d_out.shape[4]
        d_op = self._operand(d_out)
        # conv3 (1x1x1): data gradient with the act'(IN(h2)) mask and the InstanceNorm-backward sums in the epilogue
        g2 = self._new(d_out, mid)
        sums2 = self._sums_like(h2)
        self._conv(d_op, pre + "c3", g2, flip=True, mask_x=h2.t, mask_stats=h2.st, bwd_sums=sums2, pointwise=True)
        G[pre + "conv3.conv.weight"] = self._wgrad_pointwise(a_2, d_op, cout, mid, d_out)
        ops.instnorm_backward_apply(g2, h2.t, h2.st, sums2, g2)
        g2op = self._operand(g2)
        # conv2 (3x3x3)
        g1 = self._new(d_out, mid)
        sums1 = self._sums_like(h1)
        self._conv(g2op, pre + "c2", g1, flip=True, mask_x=h1.t, mask_stats=h1.st, bwd_sums=sums1)
        G[pre + "conv2.conv.weight"] = self._wgrad(a_1, g2op, self._dw(d_out, mid, mid))
        ops.instnorm_backward_apply(g1, h1.t, h1.st, sums1, g1)
        g1op = self._operand(g1)
        # conv1 (1x1x1) and the shortcut both consume a_x = act(IN(x))
        G[pre + "conv1.conv.weight"] = self._wgrad_pointwise(a_x, g1op, mid, cin, d_out)
        sums_x = self._sums_like(x)
        if has_sc:
            G[pre + "shortcut.conv.weight"] = self._wgrad(a_x, d_op, self._dw(d_out, cout, cin))
            t = self._new(d_out, cin)                 # d(a_x): the two data gradients summed BEFORE the activation mask
            self._conv(g1op, pre + "c1", t, flip=True, pointwise=True)
            self._conv(d_op, pre + "sc", t, flip=True, res=t)
            g = self._new(d_out, cin)
            ops.act_backward_stats(t, x.t, x.st, sums_x, g, slope=self.slope)
            ops.instnorm_backward_apply(g, x.t, x.st, sums_x, dx_dest)
        else:
            g = self._new(d_out, cin)
            self._conv(g1op, pre + "c1", g, flip=True, mask_x=x.t, mask_stats=x.st, bwd_sums=sums_x, pointwise=True)
            ops.instnorm_backward_apply(g, x.t, x.st, sums_x, dx_dest, add=d_out)

    def _backward_bottleneck(self, S: dict, P: dict, dlogits: torch.Tensor) -> dict:
        ch, up_in, saved = S["ch"], S["up_in"], S["saved"]
        self._zp = None
        b = self.b
        G = {}
        num_classes = dlogits.shape[1]
        final = S["final"]
        w_out = P["outc.weight"].reshape(num_classes, b).contiguous()
        d_cur = self._new(final.t, b)
        dw_out, db_out = torch.empty_like(w_out), torch.empty_like(P["outc.bias"])
        ops.head_backward(final.t, w_out, dlogits, d_cur, dw_out, db_out)
        G["outc.weight"], G["outc.bias"] = dw_out.reshape(P["outc.weight"].shape), db_out
        dskip = [None] * 4
        for j, l in zip((4, 3, 2, 1), (0, 1, 2, 3)):          # up4 .. up1 (block index map: see backward())
            pre = f"up{j}.conv."
            d_y = self._new(d_cur, ch[l])
            self._bneck_bwd(saved[8 + 2 * j], pre + "1.", d_cur, d_y, G)
            d_cat = self._new(d_cur, ch[l] + up_in[l])
            self._bneck_bwd(saved[7 + 2 * j], pre + "0.", d_y, d_cat, G)
            dskip[l] = d_cat[..., :ch[l]]
            d_cur = self._up_bwd(d_cat[..., ch[l]:], j, S, G)
        for l in (4, 3, 2, 1):
            pre = f"down{l}.conv."
            d_y = self._new(d_cur, ch[l])
            self._bneck_bwd(saved[2 * l], pre + "2.", d_cur, d_y, G)
            d_p = self._new(d_cur, ch[l - 1])
            self._bneck_bwd(saved[2 * l - 1], pre + "1.", d_y, d_p, G)
            x_prev = S["enc_out"][l - 1]
            d_cur = self._new(x_prev.t, ch[l - 1])
            ops.maxpool2_backward(x_prev.t, d_p, d_cur, dskip=dskip[l - 1])
        d_t0 = self._new(d_cur, b)
        self._bneck_bwd(saved[0], "inc.conv2.", d_cur, d_t0, G)
        dws = torch.empty_like(P["inc.conv1.weight"])
        ops.stem_conv_wgrad(S["x"], d_t0, dws)
        G["inc.conv1.weight"] = dws
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self._side_keep.clear()
        return G

    # ---- forward -------------------------------------------------------------------------------
    def _block_fwd(self, x: Act, pre: str, has_sc: bool, cout: int, out: Act, saved: list):
        n, d, h, w_, _ = x.t.shape
        dev = x.t.device
        a_x = self._operand(x.t, x.st)
        if has_sc:
            hs = self._new_act(n, d, h, w_, 2 * cout, self.dtype, dev)
            self._conv(a_x, pre + "c1", hs.t, out_stats=hs.st)  # conv1 || shortcut as one GEMM with N = 2*cout
            hh, ss = hs.view(0, cout), hs.view(cout, 2 * cout)
            a_h = self._operand(hh.t, hh.st)
            self._conv(a_h, pre + "c2", out.t, res=ss.t, out_stats=out.st)
        else:
            hh = self._new_act(n, d, h, w_, cout, self.dtype, dev)
            self._conv(a_x, pre + "c1", hh.t, out_stats=hh.st)
            a_h = self._operand(hh.t, hh.st)
            self._conv(a_h, pre + "c2", out.t, res=x.t, out_stats=out.st)
        saved.append((x, hh, a_x, a_h))

    # ---- block='SingleConv': post-activation conv -> IN -> act (conv_layers.py:50-68) ---------------------
    def _post_act(self, y: Act, out_t: torch.Tensor) -> torch.Tensor:
        """out_t = act(instnorm(y)) in the storage dtype (bf16 storage: that tensor is also the next conv's operand)."""
        if self.split:
            return ops.norm_act(y.t, y.st, slope=self.slope, full=out_t)
        return ops.norm_act(y.t, y.st, slope=self.slope, out=out_t)

    def _single_fwd(self, x_t: torch.Tensor, pre: str, cout: int, out_t, saved: list) -> torch.Tensor:
        n, d, h, w_, _ = x_t.shape
        op = self._operand(x_t)
        y = self._new_act(n, d, h, w_, cout, self.dtype, x_t.device)
        self._conv(op, pre + "c", y.t, out_stats=y.st)
        if out_t is None:
            out_t = torch.empty((n, d, h, w_, cout), dtype=self.dtype, device=x_t.device)
        self._post_act(y, out_t)
        saved.append((op, y))
        return out_t

    def _forward_single(self, x, P, num_classes, save):
        b, dt, dev = self.b, self.dtype, x.device
        n, _, D, H, W = x.shape
        ch = [b, 2 * b, 4 * b, 8 * b, 10 * b]
        dims = [(D >> l, H >> l, W >> l) for l in range(5)]
        up_in = [ch[1], ch[2], ch[3], ch[4]]
        saved: list = []
        self._zp = None
        side = self._side_stream(dev)
        if side is None:
            self.prepare(P)
        else:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.prepare(P)
        cat = [torch.empty((n, *dims[l], ch[l] + up_in[l]), dtype=dt, device=dev) for l in range(4)]
        t0 = torch.empty((n, *dims[0], b), dtype=dt, device=dev)
        ops.stem_conv_forward(x, P["inc.conv1.weight"], t0, None)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        enc = [self._single_fwd(t0, "inc.conv2.", b, cat[0][..., :b], saved)]
        for l in range(1, 5):
            p = torch.empty((n, *dims[l], ch[l - 1]), dtype=dt, device=dev)
            ops.maxpool2_forward(enc[l - 1], p, None)
            y = self._single_fwd(p, f"down{l}.conv.1.", ch[l], None, saved)
            enc.append(self._single_fwd(y, f"down{l}.conv.2.", ch[l], cat[l][..., :ch[l]] if l < 4 else None, saved))
        cur = enc[4]
        for j, l in enumerate((3, 2, 1, 0), start=1):
            ops.upsample_forward(cur, cat[l][..., ch[l]:], None)
            y = self._single_fwd(cat[l], f"up{j}.conv.0.", ch[l], None, saved)
            cur = self._single_fwd(y, f"up{j}.conv.1.", ch[l], None, saved)
        logits = torch.empty((n, num_classes, D, H, W), dtype=torch.float32, device=dev)
        ops.head_forward(cur, P["outc.weight"].reshape(num_classes, b).contiguous(), P["outc.bias"], logits)
        if not save:
            return logits, None
        return logits, dict(saved=saved, enc=enc, final=cur, x=x, ch=ch, up_in=up_in)

    def _single_bwd(self, op_in, y: Act, pre: str, d_a: torch.Tensor, cin: int, need_dx: bool = True):
        """d_a = dL/d(act(instnorm(y))) -> (dW, dL/d(input activation))."""
        g = self._new(y.t, y.C)
        sums = self._sums_like(y)
        ops.act_backward_stats(d_a, y.t, y.st, sums, g, slope=self.slope)
        ops.instnorm_backward_apply(g, y.t, y.st, sums, g)          # in place: g becomes d(y)
        dy_op = self._operand(g)
        dw = self._wgrad(op_in, dy_op, self._dw(g, y.C, cin))
        dx = None
        if need_dx:
            dx = self._new(y.t, cin)
            self._conv(dy_op, pre + "c", dx, flip=True)               # plain dgrad: the input is not normalised here
        return dw, dx

    def _backward_single(self, S: dict, P: dict, dlogits: torch.Tensor) -> dict:
        ch, up_in, saved, enc = S["ch"], S["up_in"], S["saved"], S["enc"]
        b = self.b
        self._zp = None
        G = {}
        num_classes = dlogits.shape[1]
        final = S["final"]
        w_out = P["outc.weight"].reshape(num_classes, b).contiguous()
        d_cur = self._new(final, b)
        dw_out, db_out = torch.empty_like(w_out), torch.empty_like(P["outc.bias"])
        ops.head_backward(final, w_out, dlogits, d_cur, dw_out, db_out)
        G["outc.weight"], G["outc.bias"] = dw_out.reshape(P["outc.weight"].shape), db_out
        # saved order: 0 inc | 1,2 down1 | 3,4 down2 | 5,6 down3 | 7,8 down4 | 9,10 up1 | 11,12 up2 | 13,14 up3 | 15,16 up4
        dskip = [None] * 4
        for j, l in zip((4, 3, 2, 1), (0, 1, 2, 3)):
            ia, ib = 7 + 2 * j, 8 + 2 * j
            op_b, y_b = saved[ib]
            G[f"up{j}.conv.1.conv.conv.weight"], d_y = self._single_bwd(op_b, y_b, f"up{j}.conv.1.", d_cur, ch[l])
            op_a, y_a = saved[ia]
            G[f"up{j}.conv.0.conv.conv.weight"], d_cat = self._single_bwd(op_a, y_a, f"up{j}.conv.0.", d_y, ch[l] + up_in[l])
            dskip[l] = d_cat[..., :ch[l]]
            n, d, h, w_, _ = d_cat.shape
            d_cur = torch.empty((n, d // 2, h // 2, w_ // 2, up_in[l]), dtype=self.dtype, device=d_cat.device)
            ops.upsample_backward(d_cat[..., ch[l]:], d_cur)
        for l in (4, 3, 2, 1):
            ia, ib = 2 * l - 1, 2 * l
            op_b, y_b = saved[ib]
            G[f"down{l}.conv.2.conv.conv.weight"], d_y = self._single_bwd(op_b, y_b, f"down{l}.conv.2.", d_cur, ch[l])
            op_a, y_a = saved[ia]
            G[f"down{l}.conv.1.conv.conv.weight"], d_p = self._single_bwd(op_a, y_a, f"down{l}.conv.1.", d_y, ch[l - 1])
            x_prev = enc[l - 1]
            d_cur = self._new(x_prev, ch[l - 1])
            ops.maxpool2_backward(x_prev, d_p, d_cur, dskip=dskip[l - 1])
        op0, y0 = saved[0]
        G["inc.conv2.conv.conv.weight"], d_t0 = self._single_bwd(op0, y0, "inc.conv2.", d_cur, b)
        dws = torch.empty_like(P["inc.conv1.weight"])
        ops.stem_conv_wgrad(S["x"], d_t0, dws)
        G["inc.conv1.weight"] = dws
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self._side_keep.clear()
        return G

    def forward(self, x: torch.Tensor, P: dict, num_classes: int, save: bool):
        """x fp32 [N,1,D,H,W]; P maps parameter names to fp32 tensors. Returns (logits, saved)."""
        if self.block == "SingleConv":
            if x.shape[1] != 1 or any(s % 16 for s in x.shape[2:]):
                raise ValueError(f"B200UNet needs in_ch == 1 and spatial dims that are multiples of 16, got {tuple(x.shape)}")
            return self._forward_single(x, P, num_classes, save)
        b, dt, dev = self.b, self.dtype, x.device
        n, cin0, D, H, W = x.shape
        if cin0 != 1:
            raise ValueError("B200UNet supports in_ch == 1 (CT), like every R-Super config")
        if D % 16 or H % 16 or W % 16:
            raise ValueError(f"spatial dims must be multiples of 16 (4 MaxPool3d(2) levels), got {(D, H, W)}")
        ch = [b, 2 * b, 4 * b, 8 * b, 10 * b]
        dims = [(D >> l, H >> l, W >> l) for l in range(5)]
        saved: list = []
        self._zp = None  # fresh zero pool: the statistics of this pass live in it until backward is done
        side = self._side_stream(dev)
        if side is None:
            self.prepare(P)
        else:
            side.wait_stream(torch.cuda.current_stream())   # the previous step's convs still read the packed images
            with torch.cuda.stream(side):
                self.prepare(P)

        # skip/concat buffers of decoder levels 0..3: [skip ch[l] | upsampled ch_up[l]]
        up_in = [ch[1], ch[2], ch[3], ch[4]]  # channels arriving from below at level l
        cat = [self._new_act(n, *dims[l], ch[l] + up_in[l], dt, dev) for l in range(4)]

        # inc: stem conv + BasicBlock(b, b)
        t0 = self._new_act(n, *dims[0], b, dt, dev)
        ops.stem_conv_forward(x, P["inc.conv1.weight"], t0.t, t0.st)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)   # packed weights ready before the first tensor-core conv
        blk = self._bneck_fwd if self.block == "Bottleneck" else self._block_fwd
        blk(t0, "inc.conv2.", False, b, cat[0].view(0, ch[0]), saved)
        enc_out = [cat[0].view(0, ch[0])]
        pooled = []
        for l in range(1, 5):
            p = self._new_act(n, *dims[l], ch[l - 1], dt, dev)
            ops.maxpool2_forward(enc_out[l - 1].t, p.t, p.st)
            pooled.append(p)
            y = self._new_act(n, *dims[l], ch[l], dt, dev)
            pre = f"down{l}.conv."
            blk(p, pre + "1.", True, ch[l], y, saved)
            out = cat[l].view(0, ch[l]) if l < 4 else self._new_act(n, *dims[4], ch[4], dt, dev)
            blk(y, pre + "2.", False, ch[l], out, saved)
            enc_out.append(out)
        cur = enc_out[4]
        up_ops: dict = {}
        for j, l in enumerate((3, 2, 1, 0), start=1):
            upv = cat[l].view(ch[l], ch[l] + up_in[l])
            self._up_fwd(cur, upv, j, P, up_ops)
            y = self._new_act(n, *dims[l], ch[l], dt, dev)
            pre = f"up{j}.conv."
            blk(cat[l], pre + "0.", True, ch[l], y, saved)
            out = self._new_act(n, *dims[l], ch[l], dt, dev, stats=(l != 0))
            blk(y, pre + "1.", False, ch[l], out, saved)
            cur = out
        logits = torch.empty((n, num_classes, D, H, W), dtype=torch.float32, device=dev)
        w_out = P["outc.weight"].reshape(num_classes, b).contiguous()
        ops.head_forward(cur.t, w_out, P["outc.bias"], logits)
        if not save:
            return logits, None
        return logits, dict(saved=saved, cat=cat, enc_out=enc_out, final=cur, x=x, dims=dims, ch=ch, up_in=up_in, up_ops=up_ops)

    # ---- backward ------------------------------------------------------------------------------
    def _new(self, like: torch.Tensor, c: int) -> torch.Tensor:
        n, d, h, w_, _ = like.shape
        return torch.empty((n, d, h, w_, c), dtype=self.dtype, device=like.device)

    def _dw(self, like: torch.Tensor, cout: int, cin: int, name=None) -> torch.Tensor:
        """Destination of a weight gradient: the sink's tensor for `name` (a parameter name, or a (conv1, shortcut) pair whose
        gradients are adjacent in the sink's flat buffer) when a sink is attached, else a fresh tensor."""
        if self.sink is not None and name is not None:
            t = self.sink.buffer(name)
            if t is not None:
                assert tuple(t.shape) == (cout, cin, 3, 3, 3) and t.dtype == torch.float32 and t.is_contiguous()
                return t
        return torch.empty((cout, cin, 3, 3, 3), dtype=torch.float32, device=like.device)

    def _stage_done(self, k: int):
        if self.sink is not None:
            self.sink.stage_done(k, self._side)

    def _block_bwd_identity(self, x: Act, hh: Act, a_x, a_h, pre: str, d_out: torch.Tensor, dx_dest: torch.Tensor):
        c = hh.C
        d_op = self._operand(d_out)
        g_h = self._new(d_out, c)
        sums_h = self._sums_like(hh)
        self._conv(d_op, pre + "c2", g_h, flip=True, mask_x=hh.t, mask_stats=hh.st, bwd_sums=sums_h)
        dw2 = self._wgrad(a_h, d_op, self._dw(d_out, c, c, pre + "conv2.conv.weight"))
        ops.instnorm_backward_apply(g_h, hh.t, hh.st, sums_h, g_h)  # in place: g_h becomes d(h)
        gh_op = self._operand(g_h)
        g_x = self._new(d_out, x.C)
        sums_x = self._sums_like(x)
        self._conv(gh_op, pre + "c1", g_x, flip=True, mask_x=x.t, mask_stats=x.st, bwd_sums=sums_x)
        dw1 = self._wgrad(a_x, gh_op, self._dw(d_out, c, x.C, pre + "conv1.conv.weight"))
        ops.instnorm_backward_apply(g_x, x.t, x.st, sums_x, dx_dest, add=d_out)
        return dw1, dw2

    def _block_bwd_shortcut(self, x: Act, hh: Act, a_x, a_h, pre: str, dcat2: torch.Tensor, dx_dest: torch.Tensor):
        """dcat2 [.., 2C]: channels [C:2C] hold d(out) on entry; [0:C] receives d(h)."""
        c = hh.C
        d_out = dcat2[..., c:]
        dh = dcat2[..., :c]
        d_op = self._operand(d_out)
        sums_h = self._sums_like(hh)
        self._conv(d_op, pre + "c2", dh, flip=True, mask_x=hh.t, mask_stats=hh.st, bwd_sums=sums_h)
        dw2 = self._wgrad(a_h, d_op, self._dw(dcat2, c, c, pre + "conv2.conv.weight"))
        ops.instnorm_backward_apply(dh, hh.t, hh.st, sums_h, dh)
        dcat_op = self._operand(dcat2)
        g_x = self._new(dcat2, x.C)
        sums_x = self._sums_like(x)
        self._conv(dcat_op, pre + "c1", g_x, flip=True, mask_x=x.t, mask_stats=x.st, bwd_sums=sums_x)
        dwcat = self._wgrad(a_x, dcat_op, self._dw(dcat2, 2 * c, x.C, (pre + "conv1.conv.weight", pre + "shortcut.conv.weight")))
        ops.instnorm_backward_apply(g_x, x.t, x.st, sums_x, dx_dest)
        return dwcat[:c], dw2, dwcat[c:]

    def backward(self, S: dict, P: dict, dlogits: torch.Tensor) -> dict:
        if self.block == "SingleConv":
            return self._backward_single(S, P, dlogits)
        if self.block == "Bottleneck":
            return self._backward_bottleneck(S, P, dlogits)
        ch, up_in, cat = S["ch"], S["up_in"], S["cat"]
        saved = S["saved"]
        self._zp = None
        b = self.b
        G = {}
        num_classes = dlogits.shape[1]
        final = S["final"]
        # block index map (forward order): 0 inc | 1,2 down1 | 3,4 down2 | 5,6 down3 | 7,8 down4 |
        #                                  9,10 up1 | 11,12 up2 | 13,14 up3 | 15,16 up4
        w_out = P["outc.weight"].reshape(num_classes, b).contiguous()
        d_cur = self._new(final.t, b)
        dw_out = db_out = None
        if self.sink is not None:
            dw_out, db_out = self.sink.buffer("outc.weight"), self.sink.buffer("outc.bias")
        dw_out = torch.empty_like(w_out) if dw_out is None else dw_out.view(num_classes, b)
        db_out = torch.empty_like(P["outc.bias"]) if db_out is None else db_out
        ops.head_backward(final.t, w_out, dlogits, d_cur, dw_out, db_out)
        G["outc.weight"] = dw_out.reshape(P["outc.weight"].shape)
        G["outc.bias"] = db_out

        dskip = [None] * 4
        for j, l in zip((4, 3, 2, 1), (0, 1, 2, 3)):  # up4 .. up1
            pre = f"up{j}.conv."
            ia, ib = 7 + 2 * j, 8 + 2 * j
            # block B (identity shortcut): d_out = d_cur, dx goes into the d(out) slot of block A
            xb, hb, axb, ahb = saved[ib]
            dcat2 = self._new(xb.t, 2 * ch[l])
            dw1, dw2 = self._block_bwd_identity(xb, hb, axb, ahb, pre + "1.", d_cur, dcat2[..., ch[l]:])
            G[pre + "1.conv1.conv.weight"], G[pre + "1.conv2.conv.weight"] = dw1, dw2
            # block A (conv shortcut) on the concat buffer
            xa, ha, axa, aha = saved[ia]
            d_cat = self._new(xa.t, xa.C)
            dw1, dw2, dwsc = self._block_bwd_shortcut(xa, ha, axa, aha, pre + "0.", dcat2, d_cat)
            G[pre + "0.conv1.conv.weight"], G[pre + "0.conv2.conv.weight"] = dw1, dw2
            G[pre + "0.shortcut.conv.weight"] = dwsc
            dskip[l] = d_cat[..., :ch[l]]
            d_cur = self._up_bwd(d_cat[..., ch[l]:], j, S, G)
            if j == 4:
                self._stage_done(0)           # head + up4 (the full-resolution layers: the longest weight gradients)
        self._stage_done(1)                   # up3 .. up1
        # encoder
        for l in (4, 3, 2, 1):
            pre = f"down{l}.conv."
            ia, ib = 2 * l - 1, 2 * l
            xb, hb, axb, ahb = saved[ib]
            dcat2 = self._new(xb.t, 2 * ch[l])
            dw1, dw2 = self._block_bwd_identity(xb, hb, axb, ahb, pre + "2.", d_cur, dcat2[..., ch[l]:])
            G[pre + "2.conv1.conv.weight"], G[pre + "2.conv2.conv.weight"] = dw1, dw2
            xa, ha, axa, aha = saved[ia]  # xa = pooled input
            d_p = self._new(xa.t, xa.C)
            dw1, dw2, dwsc = self._block_bwd_shortcut(xa, ha, axa, aha, pre + "1.", dcat2, d_p)
            G[pre + "1.conv1.conv.weight"], G[pre + "1.conv2.conv.weight"] = dw1, dw2
            G[pre + "1.shortcut.conv.weight"] = dwsc
            x_prev = S["enc_out"][l - 1]
            d_cur = self._new(x_prev.t, ch[l - 1])
            ops.maxpool2_backward(x_prev.t, d_p, d_cur, dskip=dskip[l - 1])
            if l == 3:
                self._stage_done(2)           # down4 + down3: half of the parameters, with down2 / down1 / inc still to run
        # inc block + stem
        x0, h0, ax0, ah0 = saved[0]
        d_t0 = self._new(x0.t, b)
        dw1, dw2 = self._block_bwd_identity(x0, h0, ax0, ah0, "inc.conv2.", d_cur, d_t0)
        G["inc.conv2.conv1.conv.weight"], G["inc.conv2.conv2.conv.weight"] = dw1, dw2
        dws = self.sink.buffer("inc.conv1.weight") if self.sink is not None else None
        if dws is None:
            dws = torch.empty_like(P["inc.conv1.weight"])
        ops.stem_conv_wgrad(S["x"], d_t0, dws)
        G["inc.conv1.weight"] = dws
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)   # all weight gradients done before autograd hands them on
        self._side_keep.clear()
        self._stage_done(3)                   # down2, down1, inc, stem (9 MB: the only all-reduce that is not hidden)
        return G


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine, names, num_classes, grad_mode, *params):
        P = dict(zip(names, [p.detach() for p in params]))
        # grad_mode = torch.is_grad_enabled() of the CALLER (it is always off inside Function.forward): under torch.no_grad()
        # (sliding-window inference, validation, the EMA net) nothing is kept for a backward pass
        need_grad = grad_mode and any(p.requires_grad for p in params)
        with torch.no_grad():
            logits, saved = engine.forward(x.detach().contiguous(), P, num_classes, save=need_grad)
        ctx.engine, ctx.names, ctx.saved_acts = engine, names, saved
        ctx.save_for_backward(*params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.saved_acts is None:
            raise RuntimeError("B200UNet: backward called but no parameter required grad in forward")
        params = ctx.saved_tensors
        P = dict(zip(ctx.names, [p.detach() for p in params]))
        with torch.no_grad():
            G = ctx.engine.backward(ctx.saved_acts, P, dlogits.contiguous().float())
        ctx.saved_acts = None
        sink = ctx.engine.sink
        # gradients the kernels wrote straight into the sink's buffers (= p.grad) are not handed to autograd again
        grads = tuple(G[nm] if (p.requires_grad and not (sink is not None and sink.owns(nm))) else None
                      for nm, p in zip(ctx.names, params))
        return (None, None, None, None, None) + grads


class B200UNet(nn.Module):
    """Constructor mirrors `UNet(in_ch, base_ch, scale, kernel_size, num_classes, block, pool, norm)`
    (rsuper_train/model/dim3/unet.py:13); unsupported variants raise instead of silently differing."""

    def __init__(self, in_ch, base_ch, scale=((2, 2, 2),) * 4, kernel_size=((3, 3, 3),) * 5, num_classes=1,
                 block="BasicBlock", pool=True, norm="in", negative_slope: float = 0.0,
                 precision: str = "bf16", return_dict: bool = True, up_mode: str = "trilinear"):
        super().__init__()
        if in_ch != 1:
            raise NotImplementedError("B200UNet: in_ch must be 1")
        if block not in ("BasicBlock", "SingleConv", "Bottleneck") or norm != "in" or not pool:
            raise NotImplementedError("B200UNet implements block='BasicBlock' | 'SingleConv' | 'Bottleneck', norm='in', pool=True "
                                      "(config/abdomenatlas/resunet_3d.yaml:9-14; model/dim3/utils.py:7-13)")
        if base_ch % 8 or (block == "Bottleneck" and base_ch % 16):
            raise ValueError("base_ch must be a multiple of 8 (16-byte channel groups; 16 for block='Bottleneck': its inner "
                             "convolutions run at half the channels)")
        sc = [list(s) if isinstance(s, (list, tuple)) else [s] * 3 for s in scale]
        ks = [list(k) if isinstance(k, (list, tuple)) else [k] * 3 for k in kernel_size]
        if any(s != [2, 2, 2] for s in sc) or len(sc) != 4 or any(k != [3, 3, 3] for k in ks):
            raise NotImplementedError("B200UNet implements scale=[[2,2,2]]*4 and kernel_size=[[3,3,3]]*5")
        if up_mode not in ("trilinear", "transposed"):
            raise ValueError("up_mode must be 'trilinear' (the reference up_block: F.interpolate, unet_utils.py:69) or 'transposed' "
                             "(ConvTranspose3d(kernel_size=2, stride=2) instead — the transposed-conv variant; vnet.py:108)")
        if up_mode == "transposed" and block == "SingleConv":
            raise NotImplementedError("up_mode='transposed' is implemented for block='BasicBlock' | 'Bottleneck'")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (bf16 storage / operands) or 'fp32' (fp32 storage, split 3xbf16 products)")
        b = base_ch
        self.base_ch, self.num_classes = b, num_classes
        self.negative_slope, self.precision, self.return_dict = negative_slope, precision, return_dict
        self.block, self.up_mode = block, up_mode
        blk = {"SingleConv": _SingleConv, "Bottleneck": _Bottleneck}.get(block, _BasicBlock)
        self.inc = _InConv(in_ch, b, blk)
        self.down1 = _Stage(b, 2 * b, True, blk)
        self.down2 = _Stage(2 * b, 4 * b, True, blk)
        self.down3 = _Stage(4 * b, 8 * b, True, blk)
        self.down4 = _Stage(8 * b, 10 * b, True, blk)
        tr = up_mode == "transposed"
        self.up1 = _Stage(10 * b + 8 * b, 8 * b, False, blk, up_channels=10 * b if tr else 0)
        self.up2 = _Stage(8 * b + 4 * b, 4 * b, False, blk, up_channels=8 * b if tr else 0)
        self.up3 = _Stage(4 * b + 2 * b, 2 * b, False, blk, up_channels=4 * b if tr else 0)
        self.up4 = _Stage(2 * b + b, b, False, blk, up_channels=2 * b if tr else 0)
        self.outc = nn.Conv3d(b, num_classes, kernel_size=1)

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop("_engine", None)  # raw device pointers: rebuilt on the next forward
        st.pop("_grad_sink", None)
        return st

    def forward(self, x):
        if not ops._on_device(x):
            raise RuntimeError("B200UNet has no CPU path: inputs must live on a CUDA (sm_100a) device")
        names, params = zip(*self.named_parameters())
        dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        engine = self.__dict__.get("_engine")
        if engine is None or engine.dtype != dtype or engine.slope != float(self.negative_slope):
            engine = _Engine(self.base_ch, self.negative_slope, dtype, self.block, getattr(self, "up_mode", "trilinear"))
            self.__dict__["_engine"] = engine  # persistent packed-weight buffers; never pickled / deep-copied
        engine.sink = self.__dict__.get("_grad_sink") if (self.block == "BasicBlock" and engine.up_mode == "trilinear") else None
        if x.device.index is not None and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):     # kernels launch on the current device's stream: follow the tensors
                out = _UNetFunction.apply(x.float(), engine, names, self.num_classes, torch.is_grad_enabled(), *params)
        else:
            out = _UNetFunction.apply(x.float(), engine, names, self.num_classes, torch.is_grad_enabled(), *params)
        # calculate_loss indexes model_output['segmentation'] (losses_foundation.py:859)
        return {"segmentation": out} if self.return_dict else out
