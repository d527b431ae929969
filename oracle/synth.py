"""ORACLE-side name for the synthetic batch generators (test infrastructure).  The generators themselves live in
`rsuper_b200.synthetic` — they are data construction, shared by the product's bench / smoke inputs and by
the tests, so that both sides of every comparison see the same tensors; nothing here computes what is being checked."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "r-super_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from rsuper_b200.synthetic import (  # noqa: E402,F401
    _LCG, _ellipsoid, default_loss_args, synthetic_image, lesion_channel_indices, make_sample, make_batch, synthetic_logits, pack_masks, unpack_masks, packed_bit)
