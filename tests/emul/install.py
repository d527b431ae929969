"""TEST INFRASTRUCTURE: route rsuper_b200's host code to the emulated kernels (tests/emul/_build/librsb_emul.so) for the whole
process — what the `emulated` fixture of tests/test_emulated_kernels.py does per test with monkeypatch, for spawned workers
(tests/test_ddp_gloo.py) that have no fixture.  The product never imports this."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import build_emul
    from rsuper_b200 import _lib, ops
    h = ctypes.CDLL(build_emul.build())
    names = set()
    for name, (res, args) in _lib.SIGNATURES.items():
        if hasattr(h, name):
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
            names.add(name)
    real = _lib.lib()

    class Lib:
        def __getattr__(self, name):
            if name in names:
                return getattr(h, name)
            if name in ("rsb_conv3_n_tile", "rsb_conv3_packed_weight_bytes", "rsb_conv3_pack_plan", "rsb_conv3_wgrad_workspace_bytes",
                        "rsb_ball_workspace_bytes", "rsb_ball_sep_workspace_bytes", "rsb_version"):
                return getattr(real, name)
            raise AttributeError(f"{name} is not emulated")

    lib = Lib()

    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"rsuper_b200: {what} failed (rc={rc}): {h.rsb_last_error().decode()}")

    ops.lib = lambda: lib
    _lib.lib = lambda: lib
    _lib.check = check
    ops.check = check
    ops._stream = lambda: None
    ops._on_device = lambda t: True
    return lib
