"""Test helper (host logic, no GPU): run the product's Python host code on CPU tensors with every kernel LAUNCH replaced by
a recorder.  Host-only entry points of the real library (packing plans, tile sizes, workspace sizes) still run, so the
recorded sequence is exactly what the GPU box would be asked to launch — names, order and scalar arguments — and every
Python branch of the engine gets executed in the GPU-less container.  Outputs are uninitialised memory: only the call
sequence means anything."""
from __future__ import annotations

import contextlib
import ctypes

import torch

HOST_ONLY = {"rsb_version", "rsb_last_error", "rsb_conv3_n_tile", "rsb_conv3_packed_weight_bytes", "rsb_conv3_pack_plan",
             "rsb_conv3_wgrad_workspace_bytes", "rsb_ball_workspace_bytes", "rsb_cc_workspace_bytes", "rsb_opt_chunk_elems",
             "rsb_opt_max_blocks", "rsb_aug_workspace_bytes", "rsb_opt_hyper_floats", "rsb_opt_fill_hyper"}


class RecordingLib:
    def __init__(self, real):
        self._real = real
        self.calls = []          # (name, tuple of scalar args; pointers -> 'p' / None, structs -> dict of scalar fields)

    @staticmethod
    def _norm(a):
        if a is None:
            return None
        if isinstance(a, ctypes.c_void_p):
            return "p" if a.value else None
        if hasattr(a, "_obj"):                      # byref(struct)
            s = a._obj
            out = {}
            for f, t in s._fields_:
                v = getattr(s, f)
                out[f] = ("p" if v else None) if t is ctypes.c_void_p else v
            return out
        if isinstance(a, (int, float)):
            return a
        return type(a).__name__

    def __getattr__(self, name):
        if name in HOST_ONLY:
            return getattr(self._real, name)
        if name == "rsb_num_sms":
            return lambda: 148

        def launch(*args):
            self.calls.append((name, tuple(self._norm(a) for a in args)))
            return 0
        return launch


@contextlib.contextmanager
def recording(monkeypatch):
    """Patch rsuper_b200 so that its host code runs on CPU tensors; yields the RecordingLib."""
    from rsuper_b200 import _lib, ops
    rec = RecordingLib(_lib.lib())
    monkeypatch.setattr(ops, "lib", lambda: rec)
    monkeypatch.setattr(_lib, "lib", lambda: rec)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_need", lambda t, dtype, name: None)
    monkeypatch.setattr(ops, "_on_device", lambda t: True)
    ops._CL_OK.clear()
    yield rec
    ops._CL_OK.clear()
