"""Golden vectors for the optimizer end of the train step from the REAL reference functions:
`training.utils.get_optimizer` (AdamW, eps=1e-5) and `training.utils.update_ema_variables`, driven in the order of
train_epoch (train_ddp.py:352-357) for four steps on formula-generated parameters / gradients (oracle/train_glue_ref.py).

Run in the build container only:  python tests/golden/make_golden_glue.py   ->  tests/golden/reference_train_glue.npz
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402
from oracle.train_glue_ref import glue_inputs  # noqa: E402

STEPS = 4
KEEP = (1, 3)   # steps whose tensors are stored (clipping active / inactive); every step's norm is stored
HYPER = dict(base_lr=6e-4, betas=(0.9, 0.999), weight_decay=0.05)   # train_ddp.py:429-465 defaults / resunet yaml


class _Net(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


def main():
    os.chdir(tempfile.mkdtemp())
    import_reference()
    tu = importlib.import_module("training.utils")
    net, ema = _Net(glue_inputs(-1)), _Net(glue_inputs(-1))
    args = types.SimpleNamespace(optimizer="adamw", momentum=0.9, **HYPER)
    opt = tu.get_optimizer(args, net)
    out = {}
    for step in range(STEPS):
        opt.zero_grad()
        for p, g in zip(net.ps, glue_inputs(step)):
            p.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)       # train_ddp.py:352
        opt.step()                                                          # :353
        tu.update_ema_variables(net, ema, 0.99, step)                       # :356-357 (ema_alpha: 0.99, config/abdomenatlas/resunet_3d.yaml:44)
        out[f"norm_{step}"] = np.float32(norm.item())
        if step not in KEEP:
            continue
        for i, (p, e) in enumerate(zip(net.ps, ema.ps)):
            out[f"p_{step}_{i}"] = p.detach().numpy().copy()
            out[f"ema_{step}_{i}"] = e.detach().numpy().copy()
            out[f"g_{step}_{i}"] = p.grad.numpy().copy()
    for i, p in enumerate(net.ps):
        out[f"exp_avg_{i}"] = opt.state[p]["exp_avg"].numpy().copy()
        out[f"exp_avg_sq_{i}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "reference_train_glue.npz"), **out)
    print(f"wrote {len(out)} arrays")


if __name__ == "__main__":
    main()
