"""CPU: the C-ABI library builds/loads and exports every symbol include/rsuper_b200.h declares, the
ctypes signature table covers them all, and argument validation fails loudly (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "rsuper_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from rsuper_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    syms = _header_symbols()
    assert len(syms) >= 23
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/rsuper_b200.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == syms, "ctypes signature table out of sync with the header"
    lib = _lib.lib()
    assert b"sm_100a" in lib.rsb_version()


def test_argument_validation_returns_errors_without_a_gpu():
    from rsuper_b200 import _lib
    lib = _lib.lib()
    a = _lib.RsbConv3Args()
    assert lib.rsb_conv3_forward(ctypes.byref(a), None) != 0
    assert b"null" in lib.rsb_last_error()
    assert lib.rsb_conv3_forward(None, None) != 0
    assert lib.rsb_dilate_ball(None, None, None, 1, 4, 4, 4, 5, None) != 0
    assert lib.rsb_conv3_packed_weight_bytes(32, 32, 1) == 27 * 32 * 64
    assert lib.rsb_conv3_packed_weight_bytes(40, 24, 3) == 3 * 27 * 48 * 64
    assert lib.rsb_conv3_packed_weight_bytes(32, 64, 6) == 6 * 2 * 27 * 32 * 64
    assert [lib.rsb_conv3_n_tile(c) for c in (32, 64, 96, 192, 320, 576, 640)] == [32, 64, 96, 96, 80, 96, 128]
    jobs = (_lib.RsbPackJob * 2)()
    nblk = ctypes.c_uint(0)
    assert lib.rsb_conv3_pack_plan(jobs, 2, ctypes.byref(nblk)) != 0 and b"null" in lib.rsb_last_error()
    for jb, (co, ci, flip) in zip(jobs, ((64, 32, 0), (64, 32, 1))):
        jb.w_a, jb.packed, jb.rows_a, jb.Cout, jb.Cin, jb.transpose_flip, jb.parts = 4096, 8192, co, co, ci, flip, 1
    assert lib.rsb_conv3_pack_plan(jobs, 2, ctypes.byref(nblk)) == 0   # host-only planning: no device access
    assert (jobs[0].co_eff, jobs[0].ci_eff, jobs[0].NT, jobs[0].nchunks) == (64, 32, 64, 1)
    assert (jobs[1].co_eff, jobs[1].ci_eff, jobs[1].NT, jobs[1].nchunks) == (32, 64, 32, 2)
    assert jobs[0].total == 27 * 64 * 32 and jobs[1].block_begin == (64 // 8) * 1   # one block per (8 co, 32 ci) tile
    assert nblk.value == jobs[1].block_begin + (32 // 8) * 2
    assert lib.rsb_conv3_pack_weights_batched(None, 2, nblk, None) != 0
    w = _lib.RsbConv3WgradArgs()
    assert lib.rsb_conv3_wgrad(ctypes.byref(w), None) != 0 and b"null" in lib.rsb_last_error()
    assert lib.rsb_conv3_wgrad_workspace_bytes(64, 96, 148) > 0
    assert lib.rsb_norm_act(None, 8, 0, None, 1e-4, 0.0, None, 8, None, 0, None, 0, None, 0, 1, 4, 4, 4, 8, None) != 0
    assert lib.rsb_act_backward_stats(None, 8, None, 8, None, None, None, 8, 0, 1e-4, 0.0, 1, 4, 4, 4, 8, None) != 0
    with pytest.raises(RuntimeError):
        _lib.check(-1, "unit-test")


def test_host_module_mirrors_reference_state_dict_on_cpu():
    """B200UNet can be constructed, saved and loaded without a GPU; forward refuses CPU tensors."""
    import torch
    from oracle.unet_ref import synthetic_state_dict
    from rsuper_b200.unet import B200UNet
    net = B200UNet(1, 8, num_classes=3)
    sd = synthetic_state_dict(8, 3)
    assert [k for k, _ in net.named_parameters()] == list(sd.keys())
    assert all(p.shape == sd[k].shape for k, p in net.named_parameters())
    net.load_state_dict(sd)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 32, 32, 32))


def test_engine_layer_tables_match_the_state_dict_names():
    """Host logic, no GPU: the engine's per-block key tables address exactly the parameters of the module tree
    (a mismatch would only show up as a KeyError on the GPU box)."""
    from oracle.unet_ref import synthetic_state_dict
    from rsuper_b200.unet import B200UNet, _Engine
    sd = synthetic_state_dict(8, 2)
    names = set(sd.keys())
    used = {"inc.conv1.weight", "outc.weight", "outc.bias"}
    for pre, has_sc in _Engine.BLOCKS:
        used |= {pre + "conv1.conv.weight", pre + "conv2.conv.weight"}
        if has_sc:
            used.add(pre + "shortcut.conv.weight")
    assert used == names
    sd1 = synthetic_state_dict(8, 2, block="SingleConv")
    used1 = {"inc.conv1.weight", "outc.weight", "outc.bias"} | {pre + "conv.conv.weight" for pre in _Engine.SINGLE_CONVS}
    assert used1 == set(sd1.keys())
    net = B200UNet(1, 8, num_classes=2, block="SingleConv")
    assert [k for k, _ in net.named_parameters()] == list(sd1.keys())
    # channel bookkeeping of the SingleConv tables: consecutive convs chain (out of one == in of the next within a stage)
    for pre in _Engine.SINGLE_CONVS:
        w = sd1[pre + "conv.conv.weight"]
        assert w.shape[2:] == (3, 3, 3) and w.shape[0] % 8 == 0 and w.shape[1] % 8 == 0


def _header_prototypes():
    """name -> (return type, [parameter types]) parsed from include/rsuper_b200.h."""
    txt = open(os.path.join(ROOT, "include", "rsuper_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for ret, name, params in re.findall(r"\b(const char\*|int|size_t|long long)\s+(rsb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = (ret, plist)
    return protos


def _kind(ctype_param: str) -> str:
    if "*" in ctype_param:
        return "ptr"
    for key, kind in (("double", "f64"), ("float", "f32"), ("long long", "i64"), ("size_t", "size"), ("unsigned int", "u32"), ("int", "i32")):
        if key in ctype_param:
            return kind
    raise AssertionError(f"unparsed parameter type: {ctype_param!r}")


def test_ctypes_signatures_match_the_header_prototypes():
    """Every ctypes argtypes list has the arity and the scalar/pointer kinds of its C prototype (ctypes itself never
    checks this; a mismatch would silently corrupt the arguments on the GPU box)."""
    from rsuper_b200 import _lib
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "i32", ctypes.c_uint: "u32", ctypes.c_float: "f32",
             ctypes.c_double: "f64", ctypes.c_longlong: "i64", ctypes.c_size_t: "size", ctypes.c_ulonglong: "u64"}
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    rets = {"const char*": ctypes.c_char_p, "int": ctypes.c_int, "size_t": ctypes.c_size_t, "long long": ctypes.c_longlong}
    for name, (ret, params) in protos.items():
        res, argtypes = _lib.SIGNATURES[name]
        assert res is rets[ret], name
        assert len(argtypes) == len(params), f"{name}: {len(argtypes)} ctypes arguments vs {len(params)} in the header"
        for i, (a, p) in enumerate(zip(argtypes, params)):
            got = kinds.get(a, "ptr" if hasattr(a, "contents") or issubclass(a, ctypes._Pointer) else None)
            assert got == _kind(p), f"{name}: argument {i} is {a} in _lib.py but `{p}` in the header"


def test_struct_layouts_match_the_header():
    """ctypes Structures mirror the header's structs field by field (names and order)."""
    from rsuper_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "rsuper_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for cname, cls in (("RsbPackJob", _lib.RsbPackJob), ("RsbConv3Args", _lib.RsbConv3Args), ("RsbConv3WgradArgs", _lib.RsbConv3WgradArgs),
                       ("RsbSegLossArgs", _lib.RsbSegLossArgs), ("RsbOptTensor", _lib.RsbOptTensor)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), txt, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.findall(r"[A-Za-z_0-9]+", first)[-1])
            names += [re.findall(r"[A-Za-z_0-9]+", r)[-1] for r in rest]
        assert names == [f for f, _ in cls._fields_], cname


def test_widened_entry_points_validate_arguments_without_a_gpu():
    """SURVEY §8f rows (N2 / N3 / N4): null pointers, empty shapes and windows outside the volume are refused with an error
    code + message before any launch."""
    from rsuper_b200 import _lib
    lib = _lib.lib()
    assert lib.rsb_opt_chunk_elems() == 4096 and lib.rsb_opt_max_blocks() >= 148
    assert lib.rsb_clip_adamw_ema_step(None, 0, 0, 1, None, None, 1.0, 6e-4, 0.9, 0.999, 1e-5, 0.05, 1, 0.99, None, None) != 0
    assert b"empty tensor table" in lib.rsb_last_error()
    assert lib.rsb_clip_adamw_ema_step(4096, 1, 1, 1, None, None, 1.0, 6e-4, 0.9, 0.999, 1e-5, 0.05, 0, 0.99, None, None) != 0
    assert b"step counts from 1" in lib.rsb_last_error()
    assert lib.rsb_clip_adamw_ema_step(4096, 1, 1, 1, None, None, 1.0, 6e-4, 0.9, 0.999, 1e-5, 0.05, 1, 0.99, None, None) != 0
    assert b"partials" in lib.rsb_last_error()
    n = lib.rsb_opt_hyper_floats()
    buf = (ctypes.c_float * n)()
    assert n == 10 and lib.rsb_opt_fill_hyper(buf, 1.0, 6e-4, 0.9, 0.999, 1e-5, 0.05, 2, 0.5) == 0      # host-only
    assert buf[0] == 1.0 and abs(buf[1] - (1 - 6e-4 * 0.05)) < 1e-7 and abs(buf[5] - 6e-4 / (1 - 0.81)) < 1e-9 and buf[8] == 0.5
    assert lib.rsb_opt_fill_hyper(buf, 1.0, 6e-4, 0.9, 0.999, 1e-5, 0.05, 0, 0.5) != 0
    assert lib.rsb_sigmoid_window_accumulate(None, 4096, 4096, 1, 2, 16, 16, 16, 8, 8, 8, 9, 0, 0, None) != 0
    assert b"leaves the volume" in lib.rsb_last_error()
    assert lib.rsb_blend_finalize(None, None, None, None, 0.5, 1, 1, 8, None) != 0
    assert lib.rsb_dilate_box3(4096, 4096, 1, 4, 4, 4, None) != 0 and b"distinct" in lib.rsb_last_error()
    assert lib.rsb_gate_by_mask(None, None, 8, None) != 0
    assert lib.rsb_cc_workspace_bytes(4, 5, 6) == 4 * 5 * 6 * 4 + 16
    assert lib.rsb_cc_label(4096, 4096, 4096, 4096, None, 4, 4, 4, None) != 0 and b"workspace" in lib.rsb_last_error()
    assert lib.rsb_cc_label(4096, 4096, 4096, None, None, 2048, 2048, 2048, None) != 0
    assert lib.rsb_unpack_masks(None, None, 1, 2, 8, 0, None) != 0
    assert lib.rsb_aug_workspace_bytes() >= 148 * 8 * 32
    assert lib.rsb_aug_stats(None, 8, None, None, None) != 0 and lib.rsb_aug_stats(4096, 0, 4096, 4096, None) != 0
    assert lib.rsb_aug_affine(None, None, 8, 1.0, 1, 0.0, 0, None, 0.0, None) != 0
    assert lib.rsb_aug_gamma(4096, 4096, 8, None, 1.0, None) != 0 and lib.rsb_aug_renorm(None, 8, None, None, None) != 0
    assert lib.rsb_aug_contrast(4096, 4096, 8, None, 1.0, None) != 0
    taps = (ctypes.c_float * 4)(0.25, 0.25, 0.25, 0.25)
    assert lib.rsb_aug_blur_axis(4096, 8192, 1, 4, 4, 4, 0, taps, 4, None) != 0 and b"odd" in lib.rsb_last_error()
    assert lib.rsb_aug_blur_axis(4096, 4096, 1, 4, 4, 4, 0, taps, 3, None) != 0 and b"distinct" in lib.rsb_last_error()


def test_widened_host_wrappers_marshal_their_arguments(monkeypatch):
    """Dry run of the new Python wrappers on CPU tensors with the device checks patched out: every call must get through
    ctypes marshalling and the C-side validation and fail only at the kernel LAUNCH (there is no GPU here) — a wrong
    argument count / order would raise ctypes.ArgumentError or a validation message instead."""
    import torch
    from rsuper_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("dry run is for the GPU-less container")
    monkeypatch.setattr(ops, "_need", lambda t, dtype, name: None)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    out, cnt = torch.zeros(1, 2, 16, 16, 16), torch.zeros(1, 1, 16, 16, 16)
    calls = [
        lambda: ops.sigmoid_window_accumulate(torch.zeros(1, 2, 8, 8, 8), out, cnt, (8, 0, 4), (8, 8, 8)),
        lambda: ops.sigmoid_window_accumulate(None, out, cnt, (8, 0, 4), (8, 8, 8)),
        lambda: ops.blend_finalize(out, cnt, threshold=0.5),
        lambda: ops.dilate_box3(torch.zeros(2, 4, 5, 6, dtype=torch.uint8)),
        lambda: ops.gate_by_mask(torch.zeros(4, 5, 6), torch.zeros(4, 5, 6, dtype=torch.uint8)),
        lambda: ops.cc_label(torch.zeros(4, 5, 6, dtype=torch.uint8), keep_largest=True),
        lambda: ops.cc_label(torch.zeros(4, 5, 6, dtype=torch.uint8)),
        lambda: ops.unpack_masks(torch.zeros(2, 2, 4, 5, 6, dtype=torch.uint8), 11),
        lambda: ops.aug_stats(torch.zeros(4, 5, 6)),
        lambda: ops.aug_affine(torch.zeros(4, 5, 6), mul=1.1, add=0.2, noise=torch.zeros(4, 5, 6), noise_std=0.1),
        lambda: ops.aug_blur(torch.zeros(1, 1, 4, 5, 6), [0.25, 0.5, 0.25]),
    ]
    for i, fn in enumerate(calls):
        with pytest.raises(RuntimeError, match="launch failed"):
            fn()
