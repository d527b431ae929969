"""Python-side operator wrappers over the C-ABI (one function per exported kernel family).

Tensors are torch CUDA tensors used purely as device buffers.  Activations are channels-last
`[N, D, H, W, C]` views (possibly channel slices of a wider buffer: the channel pitch is
`t.stride(3)`), bf16 or fp32.  InstanceNorm statistics are fp32 `[N, C, 2]` (sum, sumsq).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import RSB_BF16, RSB_F32, check, lib

EPS_IN = 1e-4  # nn.InstanceNorm3d(ch, eps=1e-4) — reference conv_layers.py:39-42

# Instrumentation used by bench.py: LAUNCHES counts kernels launched through this module; when
# PROFILE is a list every call is bracketed by CUDA events on the launching stream and recorded as
# (kernel family, algorithmic FLOPs, start_event, end_event).
LAUNCHES = 0
PROFILE = None


def _call(family: str, nkern: int, flops: float, fn, *args, what: str, desc: str = ""):
    global LAUNCHES
    LAUNCHES += nkern
    if PROFILE is None:
        check(fn(*args), what)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rc = fn(*args)
    b.record()
    PROFILE.append((family, flops, a, b, what + (" " + desc if desc else "")))
    check(rc, what)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> C.c_void_p:
    """Current CUDA stream of the current device as a raw handle (the private getter is ~10x cheaper than building a
    torch.cuda.Stream object per launch: ~300 launches per train step)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(t) -> bool:
    """Single place that decides whether a tensor (or a device) may be handed to a kernel; tests/dryrun.py patches it for
    the host-logic dry run in the GPU-less container."""
    return t.is_cuda if isinstance(t, torch.Tensor) else torch.device(t).type == "cuda"


def _p(t: Optional[torch.Tensor]) -> Optional[C.c_void_p]:
    return None if t is None else C.c_void_p(t.data_ptr())


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return RSB_BF16
    if t.dtype == torch.float32:
        return RSB_F32
    raise TypeError(f"unsupported storage dtype {t.dtype}")


_CL_OK = {}   # (shape, strides, offset mod 8) -> channel pitch of layouts that already passed validation


def _check_cl(t: torch.Tensor, name: str) -> int:
    """Validate an NDHWC (possibly channel-sliced) activation view and return its channel pitch."""
    key = (t.shape, t.stride(), t.storage_offset() & 7, _on_device(t))
    pitch = _CL_OK.get(key)
    if pitch is not None:
        return pitch
    if t.dim() != 5 or not _on_device(t):
        raise ValueError(f"{name}: expected a 5-D CUDA tensor [N,D,H,W,C], got {tuple(t.shape)}")
    n, d, h, w, c = t.shape
    pitch = t.stride(3) if w > 1 else (t.stride(2) // max(w, 1) if h > 1 else t.stride(3))
    want = (d * h * w * pitch, h * w * pitch, w * pitch, pitch, 1)
    for i, (sz, st, wt) in enumerate(zip(t.shape, t.stride(), want)):
        if sz > 1 and st != wt:
            raise ValueError(f"{name}: not a dense NDHWC view with a channel pitch (strides {t.stride()})")
    if pitch % 8 or (t.storage_offset() % 8) or c % 8:
        raise ValueError(f"{name}: channel count/pitch/offset must be multiples of 8")
    if len(_CL_OK) < 4096:
        _CL_OK[key] = pitch
    return pitch


def _st(stats: Optional[torch.Tensor], act: torch.Tensor, name: str) -> Optional[C.c_void_p]:
    """Statistics pointer; the array must share the activation's channel pitch (see rsuper_b200.h)."""
    if stats is None:
        return None
    pitch = _check_cl(act, name)
    if stats.dtype != torch.float32 or stats.dim() != 3 or stats.shape[1] != act.shape[4] or stats.shape[2] != 2 \
            or stats.stride(2) != 1 or stats.stride(1) != 2 or (stats.shape[0] > 1 and stats.stride(0) != 2 * pitch):
        raise ValueError(f"{name}: statistics {tuple(stats.shape)}/{stats.stride()} do not match activation pitch {pitch}")
    return C.c_void_p(stats.data_ptr())


def new_act(n, d, h, w, c, dtype, device) -> torch.Tensor:
    return torch.empty((n, d, h, w, c), dtype=dtype, device=device)


def new_stats(n, c, device) -> torch.Tensor:
    return torch.zeros((n, c, 2), dtype=torch.float32, device=device)


# --------------------------------------------------------------------------------------------
# conv 3x3x3
# --------------------------------------------------------------------------------------------
def conv3_pack_weights(w: torch.Tensor, transpose_flip: bool = False, split=False) -> torch.Tensor:
    """fp32 OIDHW [Cout,Cin,3,3,3] -> packed bf16 UMMA image (uint8 buffer); split = True / 3 => [hi | hi | lo] parts,
    split = 6 => the three-piece image [hi | lo | hi | lo2 | hi | lo]."""
    assert w.dtype == torch.float32 and _on_device(w) and w.is_contiguous() and w.shape[2:] == (3, 3, 3)
    cout, cin = w.shape[0], w.shape[1]
    co_eff, ci_eff = (cin, cout) if transpose_flip else (cout, cin)
    parts = 6 if split == 6 else (3 if split else 1)
    nbytes = lib().rsb_conv3_packed_weight_bytes(co_eff, ci_eff, parts)
    out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    _call("pack_weights", 1, 0.0, lib().rsb_conv3_pack_weights, _p(w), _p(out), cout, cin, int(transpose_flip), parts, _stream(), what="conv3_pack_weights")
    return out


class PackPlan:
    """Persistent packed-weight images of a set of convs + the device job table that refreshes ALL of them in one launch.

    jobs: list of (key, w_a, w_b | None, transpose_flip); the logical OIDHW tensor of a job is cat([w_a, w_b], 0).
    The plan holds raw pointers of the parameter tensors: `key()` lets the owner detect re-allocated parameters."""

    def __init__(self, jobs, split=False):
        parts = 6 if split == 6 else (3 if split else 1)
        n = len(jobs)
        arr = (_lib.RsbPackJob * n)()
        self.images = {}
        self._keep = []
        dev = jobs[0][1].device
        for j, (key, w_a, w_b, flip) in enumerate(jobs):
            pointwise = tuple(w_a.shape[2:]) == (1, 1, 1)          # 1x1x1 kernel: embedded at the centre tap of a 3x3x3 image
            assert w_a.dtype == torch.float32 and _on_device(w_a) and w_a.is_contiguous() and (pointwise or w_a.shape[2:] == (3, 3, 3))
            cin = w_a.shape[1]
            cout = w_a.shape[0]
            if w_b is not None:
                assert w_b.dtype == torch.float32 and w_b.is_contiguous() and w_b.shape[1:] == w_a.shape[1:]
                cout += w_b.shape[0]
            co_eff, ci_eff = (cin, cout) if flip else (cout, cin)
            # 1x1x1 sources live at the centre tap of a 3x3x3 image whose other taps stay zero: zeroed here, once
            img = (torch.zeros if pointwise else torch.empty)(lib().rsb_conv3_packed_weight_bytes(co_eff, ci_eff, parts), dtype=torch.uint8, device=dev)
            self.images[key] = img
            self._keep.append((w_a, w_b))
            jb = arr[j]
            jb.w_a, jb.w_b, jb.packed = w_a.data_ptr(), (w_b.data_ptr() if w_b is not None else None), img.data_ptr()
            jb.rows_a, jb.Cout, jb.Cin, jb.transpose_flip, jb.parts = w_a.shape[0], cout, cin, int(flip), parts
            jb.pointwise = int(pointwise)
        nblk = C.c_uint(0)
        check(lib().rsb_conv3_pack_plan(arr, n, C.byref(nblk)), "conv3_pack_plan")
        self.n_jobs, self.n_blocks = n, nblk.value
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.ptr_key = self.pointer_key([(w_a, w_b) for _, w_a, w_b, _ in jobs])

    @staticmethod
    def pointer_key(pairs):
        return tuple((a.data_ptr(), 0 if b is None else b.data_ptr()) for a, b in pairs)

    def refresh(self):
        """Re-pack every image from the current parameter values (one launch)."""
        _call("pack_weights", 1, 0.0, lib().rsb_conv3_pack_weights_batched, _p(self.table), self.n_jobs, self.n_blocks, _stream(),
              what="conv3_pack_weights_batched")


def conv3_forward(a_op, w_packed, y, *, a_lo=None, a_lo2=None, slope=0.0, res=None, out_stats=None,
                  mask_x=None, mask_stats=None, bwd_sums=None, planes_per_item=0, max_ctas=0, eps=EPS_IN, pointwise=False):
    """y = conv3x3x3(a_op) [+ res]; a_op is the bf16 operand from norm_act (a_lo: its split-precision low part,
    with weights packed split=True); optional fused output statistics / dgrad masking epilogue."""
    a = _lib.RsbConv3Args()
    n, d, h, w_, cin = a_op.shape
    cout = y.shape[4]
    assert y.shape[:4] == a_op.shape[:4] and a_op.dtype == torch.bfloat16
    a.N, a.D, a.H, a.W, a.Cin, a.Cout = n, d, h, w_, cin, cout
    a.dtype = dtype_code(y)
    a.a, a.a_pitch = _p(a_op), _check_cl(a_op, "a")
    if a_lo is not None:
        assert a_lo.shape == a_op.shape and a_lo.dtype == torch.bfloat16 and _check_cl(a_lo, "a_lo") == a.a_pitch
        a.a_lo = _p(a_lo)
    if a_lo2 is not None:
        assert a_lo is not None and a_lo2.shape == a_op.shape and a_lo2.dtype == torch.bfloat16 and _check_cl(a_lo2, "a_lo2") == a.a_pitch
        a.a_lo2 = _p(a_lo2)
    a.eps, a.slope = eps, slope
    a.w_packed = _p(w_packed)
    a.y, a.y_pitch = _p(y), _check_cl(y, "y")
    if res is not None:
        assert res.shape == y.shape and res.dtype == y.dtype
        a.res, a.res_pitch = _p(res), _check_cl(res, "res")
    a.out_stats = _st(out_stats, y, "out_stats")
    if mask_x is not None:
        assert mask_x.shape == y.shape and mask_x.dtype == y.dtype
        a.mask_x, a.mask_x_pitch = _p(mask_x), _check_cl(mask_x, "mask_x")
        a.mask_stats, a.bwd_sums = _st(mask_stats, mask_x, "mask_stats"), _st(bwd_sums, mask_x, "bwd_sums")
    a.planes_per_item, a.max_ctas = planes_per_item, max_ctas
    a.pointwise = int(bool(pointwise))
    _call("conv3_igemm", 1, 2.0 * (1 if pointwise else 27) * cin * cout * n * d * h * w_, lib().rsb_conv3_forward, C.byref(a), _stream(),
          what="conv3_forward", desc=f"{cin}->{cout} {n}x{d}x{h}x{w_}{' 1x1x1' if pointwise else ''}{' dgrad' if mask_x is not None else ''}"
                                     f"{' res' if res is not None else ''}")
    return y


_WG_WS = {}


def _wgrad_workspace(cout: int, cin: int, device) -> torch.Tensor:
    key = (cout, cin, str(device))
    ws = _WG_WS.get(key)
    if ws is None:
        nbytes = lib().rsb_conv3_wgrad_workspace_bytes(cout, cin, 0)
        if nbytes == 0:
            raise RuntimeError(f"rsuper_b200: no wgrad tiling for Cout={cout} Cin={cin}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WG_WS[key] = ws
    return ws


def conv3_wgrad(a_op, dy, dw, *, accumulate=False, max_ctas=0):
    """dw[Cout,Cin,3,3,3] (fp32) (+)= wgrad(dy, a_op); both operands bf16 NDHWC (a_op from norm_act)."""
    a = _lib.RsbConv3WgradArgs()
    n, d, h, w_, cin = a_op.shape
    cout = dy.shape[4]
    assert dy.shape[:4] == a_op.shape[:4] and dy.dtype == torch.bfloat16 and a_op.dtype == torch.bfloat16
    assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (cout, cin, 3, 3, 3)
    a.N, a.D, a.H, a.W, a.Cin, a.Cout = n, d, h, w_, cin, cout
    a.a, a.a_pitch = _p(a_op), _check_cl(a_op, "a")
    a.dy, a.dy_pitch = _p(dy), _check_cl(dy, "dy")
    a.dw_oidhw, a.accumulate = _p(dw), int(accumulate)
    ws = _wgrad_workspace(cout, cin, a_op.device)
    a.workspace, a.workspace_bytes = _p(ws), ws.numel()
    a.max_ctas = max_ctas
    _call("conv3_wgrad", 2, 2.0 * 27 * cin * cout * n * d * h * w_, lib().rsb_conv3_wgrad, C.byref(a), _stream(), what="conv3_wgrad",
          desc=f"{cin}x{cout} {n}x{d}x{h}x{w_}")
    return dw


def norm_act(x, stats=None, *, slope=0.0, eps=EPS_IN, split=False, out=None, full=None):
    """Conv operand tensor(s): hi = bf16(act(instnorm(x))) [, lo = bf16(value - hi) [, lo2 = bf16(value - hi - lo)]];
    split = False | True (hi, lo) | 3 (hi, lo, lo2); stats=None => cast / split only.
    full: tensor of x's dtype that receives the activation itself (post-activation block output); with full given and
    out / split not requested no bf16 piece is written and `full` is returned."""
    n, d, h, w_, c = x.shape
    only_full = full is not None and out is None and not split
    hi = None if only_full else (out if out is not None else torch.empty((n, d, h, w_, c), dtype=torch.bfloat16, device=x.device))
    lo = torch.empty((n, d, h, w_, c), dtype=torch.bfloat16, device=x.device) if split else None
    lo2 = torch.empty((n, d, h, w_, c), dtype=torch.bfloat16, device=x.device) if split == 3 else None
    if full is not None:
        assert full.dtype == x.dtype and full.shape == x.shape
    _call("norm_act", 1, 0.0, lib().rsb_norm_act, _p(x), _check_cl(x, "x"), dtype_code(x), _st(stats, x, "stats"), eps, slope,
          _p(hi), _check_cl(hi, "hi") if hi is not None else 0, _p(lo), _check_cl(lo, "lo") if lo is not None else 0,
          _p(lo2), _check_cl(lo2, "lo2") if lo2 is not None else 0, _p(full), _check_cl(full, "full") if full is not None else 0,
          n, d, h, w_, c, _stream(), what="norm_act", desc=f"{c} {n}x{d}x{h}x{w_}")
    if only_full:
        return full
    if split == 3:
        return hi, lo, lo2
    return (hi, lo) if split else hi


def act_backward_stats(d, y, y_stats, bwd_sums, g, *, slope=0.0, eps=EPS_IN):
    """g = d * act'(instnorm(y)); bwd_sums += (sum g, sum g * yhat) — backward of a post-activation (SingleConv)."""
    n, dd, h, w_, c = y.shape
    assert d.shape == y.shape and g.shape == y.shape and d.dtype == y.dtype == g.dtype
    _call("instnorm_bwd", 1, 0.0, lib().rsb_act_backward_stats, _p(d), _check_cl(d, "d"), _p(y), _check_cl(y, "y"), _st(y_stats, y, "y_stats"),
          _st(bwd_sums, y, "bwd_sums"), _p(g), _check_cl(g, "g"), dtype_code(y), eps, slope, n, dd, h, w_, c, _stream(),
          what="act_backward_stats", desc=f"{c} {n}x{dd}x{h}x{w_}")
    return g


# --------------------------------------------------------------------------------------------
# stem / head
# --------------------------------------------------------------------------------------------
def stem_conv_forward(x, w, y, out_stats=None):
    """x fp32 [N,1,D,H,W] (== NDHWC with C=1), w fp32 [Cout,1,3,3,3] -> y NDHWC."""
    n, _, d, h, w_ = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 1
    assert w.dtype == torch.float32 and w.is_contiguous()
    cout = y.shape[4]
    _call("stem", 1, 2.0 * 27 * cout * n * d * h * w_, lib().rsb_stem_conv_forward, _p(x), _p(w), _p(y), _check_cl(y, "y"), dtype_code(y), _st(out_stats, y, "out_stats"),
                                      n, d, h, w_, cout, _stream(), what="stem_conv_forward")
    return y


def stem_conv_wgrad(x, dy, dw):
    n, _, d, h, w_ = x.shape
    cout = dy.shape[4]
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == cout * 27
    _call("stem", 1, 2.0 * 27 * cout * n * d * h * w_, lib().rsb_stem_conv_wgrad, _p(x), _p(dy), _check_cl(dy, "dy"), dtype_code(dy), _p(dw),
                                    n, d, h, w_, cout, _stream(), what="stem_conv_wgrad")
    return dw


def head_forward(x, w, bias, logits):
    """x NDHWC [N,D,H,W,Cin]; w fp32 [C,Cin(,1,1,1)]; logits fp32 NCDHW [N,C,D,H,W]."""
    n, d, h, w_, cin = x.shape
    c = logits.shape[1]
    assert logits.dtype == torch.float32 and logits.is_contiguous() and w.is_contiguous()
    _call("head", 1, 0.0, lib().rsb_head_forward, _p(x), _check_cl(x, "x"), dtype_code(x), _p(w), _p(bias), _p(logits),
                                 n, d, h, w_, cin, c, _stream(), what="head_forward")
    return logits


def head_backward(x, w, dlogits, dx, dw, db):
    n, d, h, w_, cin = x.shape
    c = dlogits.shape[1]
    assert dlogits.dtype == torch.float32 and dlogits.is_contiguous()
    assert dw.is_contiguous() and db.is_contiguous() and dw.dtype == torch.float32
    _call("head", 2, 0.0, lib().rsb_head_backward, _p(x), _check_cl(x, "x"), dtype_code(x), _p(w), _p(dlogits), _p(dx),
                                  _check_cl(dx, "dx"), _p(dw), _p(db), n, d, h, w_, cin, c, _stream(), what="head_backward")


# --------------------------------------------------------------------------------------------
# pool / upsample / instnorm backward / layout
# --------------------------------------------------------------------------------------------
def maxpool2_forward(x, y, out_stats=None):
    n, d, h, w_, c = x.shape
    _call("pool", 1, 0.0, lib().rsb_maxpool2_forward, _p(x), _check_cl(x, "x"), _p(y), _check_cl(y, "y"), dtype_code(x),
                                     _st(out_stats, y, "out_stats"), n, d, h, w_, c, _stream(), what="maxpool2_forward", desc=f"{c} {n}x{d}x{h}x{w_}")
    return y


def maxpool2_backward(x, dy, dx, dskip=None):
    n, d, h, w_, c = x.shape
    _call("pool", 1, 0.0, lib().rsb_maxpool2_backward, _p(x), _check_cl(x, "x"), _p(dy), _check_cl(dy, "dy"), _p(dskip),
                                      _check_cl(dskip, "dskip") if dskip is not None else 0, _p(dx),
                                      _check_cl(dx, "dx"), dtype_code(x), n, d, h, w_, c, _stream(), what="maxpool2_backward", desc=f"{c} {n}x{d}x{h}x{w_}")
    return dx


def upsample_forward(x, y, out_stats=None):
    n, di, hi, wi, c = x.shape
    _, do, ho, wo, _ = y.shape
    _call("upsample", 1, 0.0, lib().rsb_upsample_trilinear_forward, _p(x), _check_cl(x, "x"), _p(y), _check_cl(y, "y"),
                                               dtype_code(x), _st(out_stats, y, "out_stats"), n, di, hi, wi, do, ho, wo, c,
                                               _stream(), what="upsample_forward", desc=f"{c} {n}x{do}x{ho}x{wo}")
    return y


def depth_to_space2(q, bias, up, out_stats=None):
    """up[n, 2z+a, 2y+b, 2x+c, co] = q[n, z, y, x, (4a+2b+c) * C + co] + bias[co] (+ statistics of up): the rearrangement that
    turns a 1x1x1 conv to 8 * C channels into ConvTranspose3d(kernel = stride = 2)."""
    n, d, h, w_, c8 = q.shape
    c = up.shape[4]
    assert c8 == 8 * c and tuple(up.shape[:4]) == (n, 2 * d, 2 * h, 2 * w_) and q.dtype == up.dtype
    assert bias is None or (bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == c)
    _call("upsample", 1, 0.0, lib().rsb_depth_to_space2, _p(q), _check_cl(q, "q"), _p(bias), _p(up), _check_cl(up, "up"), dtype_code(q),
          _st(out_stats, up, "out_stats"), n, d, h, w_, c, _stream(), what="depth_to_space2", desc=f"{c} {n}x{2 * d}x{2 * h}x{2 * w_}")
    return up


def space_to_depth2(up, q):
    """q[n, z, y, x, (4a+2b+c) * C + co] = up[n, 2z+a, 2y+b, 2x+c, co] — the adjoint copy of depth_to_space2."""
    n, d, h, w_, c8 = q.shape
    c = up.shape[4]
    assert c8 == 8 * c and tuple(up.shape[:4]) == (n, 2 * d, 2 * h, 2 * w_) and q.dtype == up.dtype
    _call("upsample", 1, 0.0, lib().rsb_space_to_depth2, _p(up), _check_cl(up, "up"), _p(q), _check_cl(q, "q"), dtype_code(q),
          n, d, h, w_, c, _stream(), what="space_to_depth2", desc=f"{c} {n}x{2 * d}x{2 * h}x{2 * w_}")
    return q


def upsample_backward(dy, dx, two_pass=True):
    """Adjoint of upsample_forward.  two_pass (default): separable z / y / x adjoint passes through a scratch buffer;
    False: the single-pass 3-D gather."""
    n, do, ho, wo, c = dy.shape
    _, di, hi, wi, _ = dx.shape
    ws = torch.empty(n * di * (ho + hi) * wo * c, dtype=dy.dtype, device=dy.device) if two_pass else None
    _call("upsample", 3 if ws is not None else 1, 0.0, lib().rsb_upsample_trilinear_backward, _p(dy), _check_cl(dy, "dy"), _p(dx),
          _check_cl(dx, "dx"), dtype_code(dy), n, di, hi, wi, do, ho, wo, c, _p(ws), _stream(), what="upsample_backward",
          desc=f"{c} {n}x{do}x{ho}x{wo}")
    return dx


def instnorm_backward_apply(g, x, x_stats, bwd_sums, dx, add=None, eps=EPS_IN):
    n, d, h, w_, c = x.shape
    _call("instnorm_bwd", 1, 0.0, lib().rsb_instnorm_backward_apply, _p(g), _check_cl(g, "g"), _p(x), _check_cl(x, "x"), _st(x_stats, x, "x_stats"),
                                            _st(bwd_sums, x, "bwd_sums"), _p(add), _check_cl(add, "add") if add is not None else 0,
                                            _p(dx), _check_cl(dx, "dx"), dtype_code(x), eps, n, d, h, w_, c,
                                            _stream(), what="instnorm_backward_apply", desc=f"{c} {n}x{d}x{h}x{w_}")
    return dx


def ncdhw_to_ndhwc(src, dst):
    n, c, d, h, w_ = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    _call("layout", 1, 0.0, lib().rsb_ncdhw_to_ndhwc, _p(src), _p(dst), _check_cl(dst, "dst"), dtype_code(dst), n, c, d, h, w_,
                                   _stream(), what="ncdhw_to_ndhwc")
    return dst


def ndhwc_to_ncdhw(src, dst):
    n, d, h, w_, c = src.shape
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    _call("layout", 1, 0.0, lib().rsb_ndhwc_to_ncdhw, _p(src), _check_cl(src, "src"), dtype_code(src), _p(dst), n, c, d, h, w_,
                                   _stream(), what="ndhwc_to_ncdhw")
    return dst


def channel_stats(x, stats=None):
    n, d, h, w_, c = x.shape
    if stats is None:
        pitch = _check_cl(x, "x")
        c0 = x.storage_offset() % pitch
        stats = torch.zeros((n, pitch, 2), dtype=torch.float32, device=x.device)[:, c0:c0 + c]
    _call("stats", 1, 0.0, lib().rsb_channel_stats, _p(x), _check_cl(x, "x"), dtype_code(x), _st(stats, x, "stats"), n, d, h, w_, c,
                                  _stream(), what="channel_stats")
    return stats


# --------------------------------------------------------------------------------------------
# segmentation loss
# --------------------------------------------------------------------------------------------
class SegLossState:
    """Device workspaces of one fused BCE+Dice evaluation (kept for the backward pass)."""

    def __init__(self, logits, label_u8, known_u8, class_weights, wmap=None):
        b, c = logits.shape[:2]
        v = logits[0, 0].numel()
        dev = logits.device
        self.args = _lib.RsbSegLossArgs()
        self.keep = (logits, label_u8, known_u8, class_weights, wmap)
        self.partials = torch.empty(b * c * 4, dtype=torch.float32, device=dev)
        self.coef = torch.empty(b * c * 4, dtype=torch.float32, device=dev)
        self.loss_out = torch.empty(3, dtype=torch.float32, device=dev)
        a = self.args
        a.B, a.C, a.V = b, c, v
        a.logits, a.label, a.known = _p(logits), _p(label_u8), _p(known_u8)
        a.class_weights = _p(class_weights)
        a.bce_weight_map = _p(wmap)
        a.partials, a.coef, a.loss_out = _p(self.partials), _p(self.coef), _p(self.loss_out)


def seg_loss_forward(logits, label_u8, known_u8=None, class_weights=None, wmap=None) -> SegLossState:
    assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 5
    assert label_u8.dtype == torch.uint8 and label_u8.is_contiguous() and label_u8.shape == logits.shape
    if known_u8 is not None:
        assert known_u8.dtype == torch.uint8 and known_u8.is_contiguous() and known_u8.shape == logits.shape
    if class_weights is not None:
        assert class_weights.dtype == torch.float32 and class_weights.is_contiguous()
    if wmap is not None:
        assert wmap.dtype == torch.float32 and wmap.is_contiguous() and wmap.shape == logits.shape
    st = SegLossState(logits, label_u8, known_u8, class_weights, wmap)
    _call("seg_loss", 2, 0.0, lib().rsb_seg_loss_forward, C.byref(st.args), _stream(), what="seg_loss_forward")
    return st


def seg_loss_backward(st: SegLossState, grad_scale, dlogits, accumulate=False):
    """grad_scale: device fp32 [2] = (scale of the BCE term, scale of the Dice term)."""
    assert dlogits.dtype == torch.float32 and dlogits.is_contiguous()
    assert grad_scale.dtype == torch.float32 and grad_scale.numel() == 2 and grad_scale.is_contiguous()
    _call("seg_loss", 1, 0.0, lib().rsb_seg_loss_backward, C.byref(st.args), _p(grad_scale), _p(dlogits), int(accumulate),
                                      _stream(), what="seg_loss_backward")
    return dlogits


def dilate_ball(src_u8, kernel_size: int):
    """dilate_volume(volume, kernel_size) on uint8 0/1 volumes [..., D, H, W]."""
    assert src_u8.dtype == torch.uint8 and src_u8.is_contiguous() and src_u8.dim() >= 3
    d, h, w_ = src_u8.shape[-3:]
    nvol = src_u8.numel() // (d * h * w_)
    dst = torch.empty_like(src_u8)
    tmp = torch.empty_like(src_u8)
    _call("dilate", 1, 0.0, lib().rsb_dilate_ball, _p(src_u8), _p(dst), _p(tmp), nvol, d, h, w_, int(kernel_size), _stream(), what="dilate_ball")
    return dst


# --------------------------------------------------------------------------------------------
# report-supervised losses (Volume / Ball): thin wrappers, rows are contiguous [.., V] runs
# --------------------------------------------------------------------------------------------
def _rows(t: torch.Tensor, v: int) -> int:
    assert t.is_contiguous() and t.numel() % v == 0
    return t.numel() // v


def rows_gather(src, row_map, n_rows, v):
    """dst[r] = src.view(-1, V)[row_map[r]] — lesion-channel selection (get_lesion_channels); row_map int32 [n_rows], or
    int32 [n_rows, G] for groups that max-merge several channels (-1 = no such member)."""
    assert src.is_contiguous() and row_map.dtype == torch.int32 and src.element_size() in (1, 4)
    dst = torch.empty((n_rows, v), dtype=src.dtype, device=src.device)
    if row_map.dim() == 2 and row_map.shape[1] > 1:
        assert row_map.is_contiguous() and row_map.shape[0] == n_rows
        _call("report", 1, 0.0, lib().rsb_rows_gather_max, _p(src), _p(row_map), row_map.shape[1], _p(dst), n_rows, v, src.element_size(),
              _stream(), what="rows_gather_max")
        return dst
    _call("report", 1, 0.0, lib().rsb_rows_gather, _p(src), _p(row_map), _p(dst), n_rows, v, src.element_size(), _stream(), what="rows_gather")
    return dst


def rows_scatter_add(src, row_map, dst, v, x=None):
    """Backward of rows_gather for the logits; with a group table the gradient goes to the member that attained the maximum
    (x = the gathered tensor)."""
    assert src.dtype == torch.float32 and dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    if row_map.dim() == 2 and row_map.shape[1] > 1:
        assert x is not None and x.dtype == torch.float32 and x.is_contiguous() and x.numel() == dst.numel()
        _call("report", 1, 0.0, lib().rsb_rows_scatter_add_max, _p(src), _p(x), _p(row_map), row_map.shape[1], _p(dst), _rows(src, v), v,
              _stream(), what="rows_scatter_add_max")
        return dst
    _call("report", 1, 0.0, lib().rsb_rows_scatter_add, _p(src), _p(row_map), _p(dst), _rows(src, v), v, _stream(), what="rows_scatter_add")
    return dst


U8_OR, U8_AND, U8_ANDNOT, U8_NOR, U8_NOT = 0, 1, 2, 3, 4


def u8_binary(a, b, op, out=None):
    assert a.dtype == torch.uint8 and a.is_contiguous() and (b is None or (b.dtype == torch.uint8 and b.is_contiguous() and b.numel() == a.numel()))
    out = torch.empty_like(a) if out is None else out
    _call("report", 1, 0.0, lib().rsb_u8_binary, _p(a), _p(b), _p(out), op, a.numel(), _stream(), what="u8_binary")
    return out


def u8_row_count(a, v):
    assert a.dtype == torch.uint8 and a.is_contiguous()
    counts = torch.empty(_rows(a, v), dtype=torch.int64, device=a.device)
    _call("report", 1, 0.0, lib().rsb_u8_row_count, _p(a), _p(counts), counts.numel(), v, _stream(), what="u8_row_count")
    return counts


def masked_sigmoid_sum(x, mask, scale, v):
    assert x.dtype == torch.float32 and x.is_contiguous() and mask.dtype == torch.uint8 and mask.is_contiguous()
    sums = torch.empty(_rows(x, v), dtype=torch.float32, device=x.device)
    _call("report", 1, 0.0, lib().rsb_masked_sigmoid_sum, _p(x), _p(mask), _p(scale), _p(sums), sums.numel(), v, _stream(), what="masked_sigmoid_sum")
    return sums


def masked_sigmoid_grad(x, mask, scale, coef, dx, v, accumulate=False):
    assert coef.dtype == torch.float32 and coef.is_contiguous() and dx.dtype == torch.float32 and dx.is_contiguous()
    _call("report", 1, 0.0, lib().rsb_masked_sigmoid_grad, _p(x), _p(mask), _p(scale), _p(coef), _p(dx), int(accumulate), _rows(x, v), v,
          _stream(), what="masked_sigmoid_grad")
    return dx


def ball_prepare(x, seg):
    assert x.dtype == torch.float32 and x.is_contiguous() and seg.dtype == torch.uint8 and seg.is_contiguous()
    out = torch.empty_like(x)
    _call("report", 1, 0.0, lib().rsb_ball_prepare, _p(x), _p(seg), _p(out), x.numel(), _stream(), what="ball_prepare")
    return out


def ball_remove(x_iter, mask):
    _call("report", 1, 0.0, lib().rsb_ball_remove, _p(x_iter), _p(mask), x_iter.numel(), _stream(), what="ball_remove")
    return x_iter


def ball_correlate_argmax(x_iter, taps, kernel_half):
    """Flat index (device int64 scalar, packed) of the first maximum of corr(x_iter, ball taps)."""
    d, h, w_ = x_iter.shape
    ws = torch.empty(lib().rsb_ball_workspace_bytes(d, h, w_), dtype=torch.uint8, device=x_iter.device)
    out = torch.empty(1, dtype=torch.int64, device=x_iter.device)
    _call("report", 2, 0.0, lib().rsb_ball_correlate_argmax, _p(x_iter), _p(taps), taps.shape[0], kernel_half, _p(ws), _p(out), d, h, w_,
          _stream(), what="ball_correlate_argmax")
    return out


_SEP_WS = {}


def ball_correlate_argmax_sep(x_iter, gauss, wtab, reach, gauss_host=None, wtab_host=None):
    """Same packed argmax as ball_correlate_argmax through the rows -> discs -> planes decomposition of the truncated
    Gaussian ball (gauss: device fp32 [reach + 1], wtab: device int32 [(reach + 1)^2])."""
    d, h, w_ = x_iter.shape
    assert x_iter.dtype == torch.float32 and x_iter.is_contiguous() and gauss.dtype == torch.float32 and wtab.dtype == torch.int32
    key = (d, h, w_, reach, str(x_iter.device))
    ws = _SEP_WS.get(key)
    if ws is None:
        if len(_SEP_WS) > 8:
            _SEP_WS.clear()
        ws = _SEP_WS[key] = torch.empty(lib().rsb_ball_sep_workspace_bytes(d, h, w_, reach), dtype=torch.uint8, device=x_iter.device)
    out = torch.empty(1, dtype=torch.int64, device=x_iter.device)
    if gauss_host is not None:     # host copies of the two tables (numpy float32 / int32): the tiled disc stage takes them by value
        assert gauss_host.dtype.name == "float32" and wtab_host.dtype.name == "int32" and gauss_host.size == reach + 1
    gh = C.c_void_p(gauss_host.ctypes.data) if gauss_host is not None else None
    wh = C.c_void_p(wtab_host.ctypes.data) if wtab_host is not None else None
    _call("report", 4, 0.0, lib().rsb_ball_correlate_argmax_sep, _p(x_iter), _p(gauss), _p(wtab), gh, wh, reach, _p(ws), _p(out), d, h, w_,
          _stream(), what="ball_correlate_argmax_sep")
    return out


def ball_candidates(x, mask, mode, center, half, radius2, cand, n_cand, ball_out):
    d, h, w_ = x.shape
    cz, cy, cx = center
    _call("report", 1, 0.0, lib().rsb_ball_candidates, _p(x), _p(mask), mode, cz, cy, cx, half, float(radius2), _p(cand), _p(n_cand),
          cand.shape[0], _p(ball_out), d, h, w_, _stream(), what="ball_candidates")


def ball_rank_select(cand, n_cand, max_cand, ks, masks):
    _call("report", 1, 0.0, lib().rsb_ball_rank_select, _p(cand), _p(n_cand), max_cand, ks[0], ks[1], ks[2], _p(masks[0]), _p(masks[1]),
          _p(masks[2]), _stream(), what="ball_rank_select")


def ball_rank_gwrp(cand, n_cand, max_cand, concentration, wmap):
    _call("report", 1, 0.0, lib().rsb_ball_rank_gwrp, _p(cand), _p(n_cand), max_cand, float(concentration), _p(wmap), _stream(), what="ball_rank_gwrp")


def ball_weight_map(wmap, pseudo, dilated):
    _call("report", 1, 0.0, lib().rsb_ball_weight_map, _p(wmap), _p(pseudo), _p(dilated), wmap.numel(), _stream(), what="ball_weight_map")
    return wmap


# --------------------------------------------------------------------------------------------
# sliding-window inference post-processing, connected components, bit-packed masks
# --------------------------------------------------------------------------------------------
def _need(t: torch.Tensor, dtype, name: str):
    if not (_on_device(t) and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} {tuple(t.shape)} on {t.device}")


def sigmoid_window_accumulate(pred: Optional[torch.Tensor], out: torch.Tensor, count: torch.Tensor, start, window):
    """out[:, :, win] += sigmoid(pred); count[:, 0, win] += 1 (inference3d.py:80-100).  pred None = gated-out window."""
    _need(out, torch.float32, "out"); _need(count, torch.float32, "count")
    b, c, d, h, w_ = out.shape
    assert count.numel() == b * d * h * w_
    wd, wh, ww = window
    if pred is not None:
        _need(pred, torch.float32, "pred")
        assert tuple(pred.shape) == (b, c, wd, wh, ww), f"window logits {tuple(pred.shape)} != {(b, c, wd, wh, ww)}"
    _call("infer", 1, 0.0, lib().rsb_sigmoid_window_accumulate, _p(pred), _p(out), _p(count), b, c, d, h, w_, wd, wh, ww,
          int(start[0]), int(start[1]), int(start[2]), _stream(), what="sigmoid_window_accumulate")


def blend_finalize(acc: torch.Tensor, count: torch.Tensor, threshold: Optional[float] = None, in_place: bool = True):
    """prob = acc / count (in place by default); with a threshold also returns the uint8 mask prob > threshold."""
    _need(acc, torch.float32, "acc"); _need(count, torch.float32, "count")
    b, c = acc.shape[0], acc.shape[1]
    v = acc[0, 0].numel()
    prob = acc if in_place else torch.empty_like(acc)
    mask = torch.empty(acc.shape, dtype=torch.uint8, device=acc.device) if threshold is not None else None
    _call("infer", 1, 0.0, lib().rsb_blend_finalize, _p(acc), _p(count), _p(prob), _p(mask), float(threshold if threshold is not None else 0.5),
          b, c, v, _stream(), what="blend_finalize")
    return prob, mask


def dilate_box3(mask: torch.Tensor) -> torch.Tensor:
    """ndi.binary_dilation(mask, structure=np.ones((3, 3, 3))) per [D, H, W] volume of a uint8 [..., D, H, W] tensor."""
    _need(mask, torch.uint8, "mask")
    d, h, w_ = mask.shape[-3:]
    out = torch.empty_like(mask)
    _call("infer", 1, 0.0, lib().rsb_dilate_box3, _p(mask), _p(out), mask.numel() // (d * h * w_), d, h, w_, _stream(), what="dilate_box3")
    return out


def gate_by_mask(prob: torch.Tensor, organ: torch.Tensor) -> torch.Tensor:
    """prob *= organ (uint8 0/1), in place (predict_abdomenatlas.py:678-680)."""
    _need(prob, torch.float32, "prob"); _need(organ, torch.uint8, "organ")
    assert prob.numel() == organ.numel()
    _call("infer", 1, 0.0, lib().rsb_gate_by_mask, _p(prob), _p(organ), prob.numel(), _stream(), what="gate_by_mask")
    return prob


def cc_label(mask: torch.Tensor, keep_largest: bool = False):
    """Face-connected components of one uint8 [D, H, W] volume.  Returns (labels int32 [D, H, W] = smallest linear index of the
    voxel's component or -1, n_components int32[1] on the device, largest uint8 [D, H, W] or None)."""
    _need(mask, torch.uint8, "mask")
    if mask.dim() != 3:
        raise ValueError(f"cc_label: expected one [D, H, W] volume, got {tuple(mask.shape)}")
    d, h, w_ = mask.shape
    labels = torch.empty((d, h, w_), dtype=torch.int32, device=mask.device)
    n = torch.empty(1, dtype=torch.int32, device=mask.device)
    ws = largest = None
    if keep_largest:
        ws = torch.empty((lib().rsb_cc_workspace_bytes(d, h, w_) + 7) // 8, dtype=torch.int64, device=mask.device)
        largest = torch.empty_like(mask)
    _call("infer", 5 if keep_largest else 3, 0.0, lib().rsb_cc_label, _p(mask), _p(labels), _p(n), _p(largest), _p(ws), d, h, w_, _stream(),
          what="cc_label")
    return labels, n, largest


def unpack_masks(packed: torch.Tensor, num_classes: int, invert: bool = False) -> torch.Tensor:
    """np.unpackbits(packed, axis=0)[:C] per sample on the device: uint8 [B, ceil(C/8), D, H, W] -> uint8 [B, C, D, H, W]."""
    _need(packed, torch.uint8, "packed")
    if packed.dim() != 5 or packed.shape[1] != (num_classes + 7) // 8:
        raise ValueError(f"unpack_masks: expected [B, {(num_classes + 7) // 8}, D, H, W] for {num_classes} classes, got {tuple(packed.shape)}")
    b, _, d, h, w_ = packed.shape
    out = torch.empty((b, num_classes, d, h, w_), dtype=torch.uint8, device=packed.device)
    _call("assemble", 1, 0.0, lib().rsb_unpack_masks, _p(packed), _p(out), b, num_classes, d * h * w_, int(invert), _stream(), what="unpack_masks")
    return out


# --------------------------------------------------------------------------------------------
# online intensity augmentations: deterministic kernels, the draws come from the host mirror
# --------------------------------------------------------------------------------------------
def aug_stats(x: torch.Tensor) -> torch.Tensor:
    """Device float32[4] = (min, max, mean, unbiased std) of the whole tensor."""
    _need(x, torch.float32, "x")
    ws = torch.empty((lib().rsb_aug_workspace_bytes() + 7) // 8, dtype=torch.int64, device=x.device)
    out = torch.empty(4, dtype=torch.float32, device=x.device)
    _call("augment", 2, 0.0, lib().rsb_aug_stats, _p(x), x.numel(), _p(ws), _p(out), _stream(), what="aug_stats")
    return out


def aug_affine(x: torch.Tensor, mul: Optional[float] = None, add: Optional[float] = None, noise: Optional[torch.Tensor] = None,
               noise_std: float = 0.0) -> torch.Tensor:
    _need(x, torch.float32, "x")
    if noise is not None:
        _need(noise, torch.float32, "noise")
        assert noise.numel() == x.numel()
    y = torch.empty_like(x)
    _call("augment", 1, 0.0, lib().rsb_aug_affine, _p(x), _p(y), x.numel(), float(mul if mul is not None else 1.0), int(mul is not None),
          float(add if add is not None else 0.0), int(add is not None), _p(noise), float(noise_std), _stream(), what="aug_affine")
    return y


def aug_gamma(x: torch.Tensor, gamma: float, retain_stats: bool = True) -> torch.Tensor:
    _need(x, torch.float32, "x")
    sx = aug_stats(x)
    y = torch.empty_like(x)
    _call("augment", 1, 0.0, lib().rsb_aug_gamma, _p(x), _p(y), x.numel(), _p(sx), float(gamma), _stream(), what="aug_gamma")
    if retain_stats:
        sy = aug_stats(y)
        _call("augment", 1, 0.0, lib().rsb_aug_renorm, _p(y), y.numel(), _p(sy), _p(sx), _stream(), what="aug_renorm")
    return y


def aug_contrast(x: torch.Tensor, factor: float) -> torch.Tensor:
    _need(x, torch.float32, "x")
    sx = aug_stats(x)
    y = torch.empty_like(x)
    _call("augment", 1, 0.0, lib().rsb_aug_contrast, _p(x), _p(y), x.numel(), _p(sx), float(factor), _stream(), what="aug_contrast")
    return y


def aug_blur(x: torch.Tensor, taps) -> torch.Tensor:
    """Separable zero-padded blur of every [D, H, W] volume of x with the 1-D kernel `taps` (sequence of floats, odd length)."""
    _need(x, torch.float32, "x")
    d, h, w_ = x.shape[-3:]
    n_vol = x.numel() // (d * h * w_)
    arr = (C.c_float * len(taps))(*[float(t) for t in taps])
    a, b = torch.empty_like(x), torch.empty_like(x)
    src = x
    for axis, dst in ((0, a), (1, b), (2, a)):
        _call("augment", 1, 0.0, lib().rsb_aug_blur_axis, _p(src), _p(dst), n_vol, d, h, w_, axis, arr, len(taps), _stream(), what="aug_blur_axis")
        src = dst
    return a


# --------------------------------------------------------------------------------------------
# MedFormer voxel-side kernels (csrc/medformer.cu)
# --------------------------------------------------------------------------------------------
def dwconv3(a, w, y=None, flip=False):
    """Depthwise 3x3x3 conv (padding 1, groups = C, no bias): a NDHWC, w fp32 [C,1,3,3,3]; flip=True = its data gradient."""
    n, d, h, w_, c = a.shape
    assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == c * 27
    if y is None:
        y = torch.empty_like(a)
    assert y.shape == a.shape and y.dtype == a.dtype
    _call("medformer", 1, 2.0 * 27 * c * n * d * h * w_, lib().rsb_dwconv3_forward, _p(a), _check_cl(a, "a"), _p(w), _p(y), _check_cl(y, "y"),
          dtype_code(a), int(flip), n, d, h, w_, c, _stream(), what="dwconv3_forward", desc=f"{c} {n}x{d}x{h}x{w_}")
    return y


def dwconv3_wgrad(a, dy, dw=None):
    n, d, h, w_, c = a.shape
    assert dy.shape == a.shape and dy.dtype == a.dtype
    if dw is None:
        dw = torch.empty((c, 1, 3, 3, 3), dtype=torch.float32, device=a.device)
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == c * 27
    _call("medformer", 1, 2.0 * 27 * c * n * d * h * w_, lib().rsb_dwconv3_wgrad, _p(a), _check_cl(a, "a"), _p(dy), _check_cl(dy, "dy"),
          dtype_code(a), _p(dw), n, d, h, w_, c, _stream(), what="dwconv3_wgrad", desc=f"{c} {n}x{d}x{h}x{w_}")
    return dw


def scale_channels(x, s, y=None):
    """y[n,...,c] = x[n,...,c] * s[n,c] (SEBlock)."""
    n, c = x.shape[0], x.shape[4]
    v = x.shape[1] * x.shape[2] * x.shape[3]
    assert s.dtype == torch.float32 and s.is_contiguous() and tuple(s.shape) == (n, c)
    if y is None:
        y = torch.empty_like(x)
    _call("medformer", 1, 0.0, lib().rsb_scale_channels, _p(x), _check_cl(x, "x"), _p(s), _p(y), _check_cl(y, "y"), dtype_code(x), n, v, c,
          _stream(), what="scale_channels")
    return y


def channel_dot(a, b):
    """out[n,c] = sum over the voxels of a * b (fp32)."""
    n, c = a.shape[0], a.shape[4]
    v = a.shape[1] * a.shape[2] * a.shape[3]
    assert b.shape == a.shape and b.dtype == a.dtype
    out = torch.empty((n, c), dtype=torch.float32, device=a.device)
    _call("medformer", 1, 0.0, lib().rsb_channel_dot, _p(a), _check_cl(a, "a"), _p(b), _check_cl(b, "b"), dtype_code(a), _p(out), n, v, c,
          _stream(), what="channel_dot")
    return out


def _colstats_ws(rows, device):
    return torch.empty(lib().rsb_colstats_workspace_floats(rows), dtype=torch.float32, device=device)


def softmax_pool_forward(feat, logit, k):
    """SemanticMapGeneration's softmax over the voxels + pooling: -> (smap [N,C,k] fp32, ms [N,64] column statistics)."""
    n, c = feat.shape[0], feat.shape[4]
    v = feat.shape[1] * feat.shape[2] * feat.shape[3]
    assert logit.shape[:4] == feat.shape[:4] and logit.dtype == feat.dtype and logit.shape[4] >= k
    ms = torch.empty((n, 64), dtype=torch.float32, device=feat.device)
    smap = torch.empty((n, c, k), dtype=torch.float32, device=feat.device)
    ws = _colstats_ws(n, feat.device)
    _call("medformer", 4, 2.0 * n * v * c * k, lib().rsb_softmax_pool_forward, _p(feat), _check_cl(feat, "feat"), _p(logit), _check_cl(logit, "logit"),
          dtype_code(feat), _p(ms), _p(ws), _p(smap), n, v, c, k, _stream(), what="softmax_pool_forward")
    return smap, ms


def softmax_pool_backward(feat, logit, ms, d_smap, tk):
    n, c = feat.shape[0], feat.shape[4]
    v = feat.shape[1] * feat.shape[2] * feat.shape[3]
    k, kp = d_smap.shape[2], logit.shape[4]
    assert d_smap.dtype == torch.float32 and d_smap.is_contiguous() and tk.dtype == torch.float32 and tk.is_contiguous()
    dfeat, dlogit = torch.empty_like(feat), torch.empty_like(logit)
    _call("medformer", 1, 4.0 * n * v * c * k, lib().rsb_softmax_pool_backward, _p(feat), _check_cl(feat, "feat"), _p(logit), _check_cl(logit, "logit"),
          dtype_code(feat), _p(ms), _p(d_smap), _p(tk), _p(dfeat), _check_cl(dfeat, "dfeat"), _p(dlogit), _check_cl(dlogit, "dlogit"),
          n, v, c, k, kp, _stream(), what="softmax_pool_backward")
    return dfeat, dlogit


def biattention_forward(qv, mq, mv, heads):
    """qv NDHWC [N,D,H,W,2C] (query channels then value channels); mq / mv fp32 [N,heads,J,dh] -> (feat_out NDHWC [.., C],
    map_out fp32 [N,heads,J,dh], ms)."""
    n, d, h, w_, c2 = qv.shape
    c = c2 // 2
    dh, j = c // heads, mq.shape[2]
    v = d * h * w_
    assert tuple(mq.shape) == (n, heads, j, dh) and mq.dtype == torch.float32 and mq.is_contiguous() and mv.shape == mq.shape and mv.is_contiguous()
    fo = torch.empty((n, d, h, w_, c), dtype=qv.dtype, device=qv.device)
    mo = torch.empty_like(mq)
    ms = torch.empty((n * heads, 64), dtype=torch.float32, device=qv.device)
    ws = _colstats_ws(n * heads, qv.device)
    _call("medformer", 4, 6.0 * n * v * c * j, lib().rsb_biattention_forward, _p(qv), _p(qv[..., c:]), _check_cl(qv, "qv"), dtype_code(qv), _p(mq), _p(mv),
          _p(ms), _p(ws), _p(fo), _check_cl(fo, "fo"), _p(mo), n, v, heads, dh, j, float(dh) ** -0.5, _stream(),
          what="biattention_forward")
    return fo, mo, ms


def biattention_backward(qv, mq, mv, ms, dfo, dmo, tj, heads):
    n, d, h, w_, c2 = qv.shape
    c = c2 // 2
    dh, j = c // heads, mq.shape[2]
    v = d * h * w_
    assert dfo.dtype == qv.dtype and dmo.dtype == torch.float32 and dmo.is_contiguous() and tj.is_contiguous()
    dqv = torch.empty_like(qv)
    dmq, dmv = torch.empty_like(mq), torch.empty_like(mv)
    _call("medformer", 1, 12.0 * n * v * c * j, lib().rsb_biattention_backward, _p(qv), _p(qv[..., c:]), _check_cl(qv, "qv"), dtype_code(qv), _p(mq), _p(mv),
          _p(ms), _p(dfo), _check_cl(dfo, "dfo"), _p(dmo), _p(tj), _p(dqv), _p(dqv[..., c:]), _check_cl(dqv, "dqv"), _p(dmq), _p(dmv), n, v, heads, dh, j,
          float(dh) ** -0.5, _stream(), what="biattention_backward")
    return dqv, dmq, dmv
