"""ctypes binding of librsuper_b200.so (the C-ABI declared in include/rsuper_b200.h).

PyTorch is used only for device memory and streams: every call passes raw device pointers,
sizes and the current CUDA stream.  There is NO fallback: if the library is missing or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RSB_LIB: an alternative build of the same library (A/B measurements of a kernel change on one box); default = the in-tree .so
LIB_PATH = os.environ.get("RSB_LIB") or os.path.join(_HERE, "librsuper_b200.so")

RSB_BF16 = 0
RSB_F32 = 1

c_void_p = C.c_void_p
c_int = C.c_int
c_float = C.c_float
c_size_t = C.c_size_t
c_ll = C.c_longlong


class RsbConv3Args(C.Structure):
    _fields_ = [
        ("N", c_int), ("D", c_int), ("H", c_int), ("W", c_int),
        ("Cin", c_int), ("Cout", c_int),
        ("dtype", c_int),
        ("a", c_void_p), ("a_pitch", c_int), ("a_lo", c_void_p), ("a_lo2", c_void_p),
        ("w_packed", c_void_p),
        ("y", c_void_p), ("y_pitch", c_int),
        ("res", c_void_p), ("res_pitch", c_int),
        ("out_stats", c_void_p),
        ("mask_x", c_void_p), ("mask_x_pitch", c_int),
        ("mask_stats", c_void_p),
        ("bwd_sums", c_void_p),
        ("eps", c_float), ("slope", c_float),
        ("planes_per_item", c_int), ("max_ctas", c_int),
        ("pointwise", c_int),
    ]


class RsbPackJob(C.Structure):
    _fields_ = [
        ("w_a", c_void_p), ("w_b", c_void_p), ("packed", c_void_p),
        ("rows_a", c_int), ("Cout", c_int), ("Cin", c_int),
        ("transpose_flip", c_int), ("parts", c_int),
        ("co_eff", c_int), ("ci_eff", c_int), ("NT", c_int), ("ntiles", c_int), ("nchunks", c_int),
        ("block_begin", C.c_uint),
        ("total", C.c_ulonglong),
        ("pointwise", c_int),
    ]


class RsbConv3WgradArgs(C.Structure):
    _fields_ = [
        ("N", c_int), ("D", c_int), ("H", c_int), ("W", c_int),
        ("Cin", c_int), ("Cout", c_int),
        ("a", c_void_p), ("a_pitch", c_int),
        ("dy", c_void_p), ("dy_pitch", c_int),
        ("dw_oidhw", c_void_p), ("accumulate", c_int),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("max_ctas", c_int),
    ]


class RsbSegLossArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("C", c_int), ("V", c_ll),
        ("logits", c_void_p), ("label", c_void_p), ("known", c_void_p),
        ("class_weights", c_void_p), ("bce_weight_map", c_void_p),
        ("partials", c_void_p), ("coef", c_void_p), ("loss_out", c_void_p),
    ]


class RsbOptTensor(C.Structure):
    _fields_ = [
        ("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("ema", c_void_p),
        ("n", c_ll), ("chunk_begin", c_ll),
    ]


c_double = C.c_double

# name -> (restype, argtypes); every symbol include/rsuper_b200.h declares
SIGNATURES = {
    "rsb_version": (C.c_char_p, []),
    "rsb_last_error": (C.c_char_p, []),
    "rsb_num_sms": (c_int, []),
    "rsb_conv3_n_tile": (c_int, [c_int]),
    "rsb_conv3_packed_weight_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rsb_conv3_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_conv3_pack_plan": (c_int, [C.POINTER(RsbPackJob), c_int, C.POINTER(C.c_uint)]),
    "rsb_conv3_pack_weights_batched": (c_int, [c_void_p, c_int, C.c_uint, c_void_p]),
    "rsb_conv3_forward": (c_int, [C.POINTER(RsbConv3Args), c_void_p]),
    "rsb_debug_set_timing_buffer": (c_int, [c_void_p]),
    "rsb_debug_set_wgrad_timing_buffer": (c_int, [c_void_p]),
    "rsb_conv3_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rsb_conv3_wgrad": (c_int, [C.POINTER(RsbConv3WgradArgs), c_void_p]),
    "rsb_norm_act": (c_int, [c_void_p, c_int, c_int, c_void_p, c_float, c_float, c_void_p, c_int, c_void_p, c_int,
                             c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_act_backward_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_float, c_float, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_stem_conv_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                      c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_stem_conv_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_head_forward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_head_backward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p]),
    "rsb_maxpool2_forward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                     c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_maxpool2_backward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                      c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p]),
    "rsb_depth_to_space2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_space_to_depth2": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_upsample_trilinear_forward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                               c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                               c_int, c_void_p]),
    "rsb_upsample_trilinear_backward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int,
                                                c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                c_int, c_void_p, c_void_p]),
    "rsb_instnorm_backward_apply": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                            c_void_p, c_int, c_void_p, c_int, c_int, c_float,
                                            c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_ncdhw_to_ndhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p]),
    "rsb_ndhwc_to_ncdhw": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p]),
    "rsb_channel_stats": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p]),
    "rsb_seg_loss_forward": (c_int, [C.POINTER(RsbSegLossArgs), c_void_p]),
    "rsb_seg_loss_backward": (c_int, [C.POINTER(RsbSegLossArgs), c_void_p, c_void_p, c_int,
                                      c_void_p]),
    "rsb_dilate_ball": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_void_p]),
    "rsb_rows_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "rsb_rows_scatter_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "rsb_rows_gather_max": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "rsb_rows_scatter_add_max": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_ll, c_void_p]),
    "rsb_u8_binary": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "rsb_u8_row_count": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "rsb_masked_sigmoid_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    "rsb_masked_sigmoid_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p]),
    "rsb_ball_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "rsb_ball_remove": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "rsb_ball_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rsb_ball_correlate_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rsb_ball_sep_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rsb_ball_correlate_argmax_sep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_void_p]),
    "rsb_ball_candidates": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int,
                                    c_void_p, c_int, c_int, c_int, c_void_p]),
    "rsb_ball_rank_select": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rsb_ball_rank_gwrp": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "rsb_ball_weight_map": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "rsb_opt_chunk_elems": (c_ll, []),
    "rsb_opt_max_blocks": (c_int, []),
    "rsb_clip_adamw_ema_step": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_double, c_double, c_double, c_double,
                                        c_double, c_double, c_ll, c_double, c_void_p, c_void_p]),
    "rsb_opt_hyper_floats": (c_int, []),
    "rsb_opt_fill_hyper": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_double, c_double, c_ll, c_double]),
    "rsb_sigmoid_window_accumulate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                              c_int, c_int, c_int, c_void_p]),
    "rsb_blend_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_ll, c_void_p]),
    "rsb_dilate_box3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_gate_by_mask": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "rsb_cc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rsb_cc_label": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rsb_unpack_masks": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p]),
    "rsb_aug_workspace_bytes": (c_size_t, []),
    "rsb_aug_stats": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "rsb_aug_affine": (c_int, [c_void_p, c_void_p, c_ll, c_float, c_int, c_float, c_int, c_void_p, c_float, c_void_p]),
    "rsb_aug_gamma": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_float, c_void_p]),
    "rsb_aug_renorm": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "rsb_aug_contrast": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_float, c_void_p]),
    "rsb_aug_blur_axis": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float), c_int, c_void_p]),
    "rsb_dwconv3_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_dwconv3_wgrad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rsb_scale_channels": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_int, c_void_p]),
    "rsb_channel_dot": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "rsb_colstats_workspace_floats": (c_size_t, [c_int]),
    "rsb_softmax_pool_forward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_int,
                                         c_void_p]),
    "rsb_softmax_pool_backward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                          c_int, c_int, c_ll, c_int, c_int, c_int, c_void_p]),
    "rsb_biattention_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                        c_int, c_ll, c_int, c_int, c_int, c_float, c_void_p]),
    "rsb_biattention_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_float, c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly if it is missing (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python r-super_b200/build.py` "
                "(rsuper_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        lax = os.environ.get("RSB_LOADER_LAX") == "1"  # bring-up probes only
        for name, (res, args) in SIGNATURES.items():
            if lax and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rsb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"rsuper_b200: {what} failed (rc={rc}): {msg}")
