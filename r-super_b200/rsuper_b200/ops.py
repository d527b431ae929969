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


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[C.c_void_p]:
    return None if t is None else C.c_void_p(t.data_ptr())


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return RSB_BF16
    if t.dtype == torch.float32:
        return RSB_F32
    raise TypeError(f"unsupported storage dtype {t.dtype}")


def _check_cl(t: torch.Tensor, name: str) -> int:
    """Validate an NDHWC (possibly channel-sliced) activation view and return its channel pitch."""
    if t.dim() != 5 or not t.is_cuda:
        raise ValueError(f"{name}: expected a 5-D CUDA tensor [N,D,H,W,C], got {tuple(t.shape)}")
    n, d, h, w, c = t.shape
    pitch = t.stride(3)
    if t.stride(4) != 1 or t.stride(2) != w * pitch or t.stride(1) != h * w * pitch or t.stride(0) != d * h * w * pitch:
        raise ValueError(f"{name}: not a dense NDHWC view with a channel pitch (strides {t.stride()})")
    if pitch % 8 or (t.storage_offset() % 8):
        raise ValueError(f"{name}: channel pitch/offset must be multiples of 8")
    return pitch


def new_act(n, d, h, w, c, dtype, device) -> torch.Tensor:
    return torch.empty((n, d, h, w, c), dtype=dtype, device=device)


def new_stats(n, c, device) -> torch.Tensor:
    return torch.zeros((n, c, 2), dtype=torch.float32, device=device)


# --------------------------------------------------------------------------------------------
# conv 3x3x3
# --------------------------------------------------------------------------------------------
def conv3_pack_weights(w: torch.Tensor, transpose_flip: bool = False) -> torch.Tensor:
    """fp32 OIDHW [Cout,Cin,3,3,3] -> packed bf16 UMMA image (uint8 buffer)."""
    assert w.dtype == torch.float32 and w.is_cuda and w.is_contiguous() and w.shape[2:] == (3, 3, 3)
    cout, cin = w.shape[0], w.shape[1]
    co_eff, ci_eff = (cin, cout) if transpose_flip else (cout, cin)
    nbytes = lib().rsb_conv3_packed_weight_bytes(co_eff, ci_eff)
    out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    check(lib().rsb_conv3_pack_weights(_p(w), _p(out), cout, cin, int(transpose_flip), _stream()),
          "conv3_pack_weights")
    return out


def conv3_forward(x, w_packed, y, *, in_stats=None, slope=0.0, res=None, out_stats=None,
                  mask_x=None, mask_stats=None, bwd_sums=None, planes_per_item=0, n_tile=0,
                  max_ctas=0, eps=EPS_IN):
    """y = conv3x3x3(act(instnorm(x))) [+ res]; optional fused statistics / dgrad masking epilogue."""
    a = _lib.RsbConv3Args()
    n, d, h, w_, cin = x.shape
    cout = y.shape[4]
    assert y.shape[:4] == x.shape[:4] and y.dtype == x.dtype
    a.N, a.D, a.H, a.W, a.Cin, a.Cout = n, d, h, w_, cin, cout
    a.dtype = dtype_code(x)
    a.x, a.x_pitch = _p(x), _check_cl(x, "x")
    a.in_stats = _p(in_stats)
    a.eps, a.slope = eps, slope
    a.w_packed = _p(w_packed)
    a.y, a.y_pitch = _p(y), _check_cl(y, "y")
    if res is not None:
        assert res.shape == y.shape and res.dtype == y.dtype
        a.res, a.res_pitch = _p(res), _check_cl(res, "res")
    a.out_stats = _p(out_stats)
    if mask_x is not None:
        assert mask_x.shape == y.shape and mask_x.dtype == y.dtype
        a.mask_x, a.mask_x_pitch = _p(mask_x), _check_cl(mask_x, "mask_x")
        a.mask_stats, a.bwd_sums = _p(mask_stats), _p(bwd_sums)
    a.planes_per_item, a.n_tile, a.max_ctas = planes_per_item, n_tile, max_ctas
    check(lib().rsb_conv3_forward(C.byref(a), _stream()), "conv3_forward")
    return y
