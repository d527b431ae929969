"""Probe: end-to-end loop variants of the graph-replayed step (cfg2 shape) — where does the time of the input upload go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))
import torch
from oracle import losses_ref as LR, synth
from oracle.unet_ref import synthetic_image
from rsuper_b200 import losses
from rsuper_b200.optim import B200AdamW
from rsuper_b200.train_step import B200TrainStep
from rsuper_b200.unet import B200UNet
dev = torch.device("cuda:0")
classes = ["organ", "pancreatic_lesion"]
S, B, K = 128, 2, 20
net = B200UNet(1, 32, num_classes=2, precision="bf16").to(dev)
x = synthetic_image(B, S, S, S, seed=3).pin_memory()
lab = synth.make_batch(["mask"] * B, classes, (S, S, S), seed=5, device="cpu")["label"].contiguous().pin_memory()
args = LR.default_args(report_volume_loss_basic=0.0); args.nan_check = False
loss_fn = lambda out, lb: losses.calculate_loss(out, lb, None, args, None, None, None, None, classes)["overall"]
params = list(net.parameters())
opt = B200AdamW(params, lr=1e-4, weight_decay=0.05, max_norm=1.0, ema_params=[p.detach().clone() for p in params], capturable=True)
step = B200TrainStep(net, loss_fn, opt, [x.to(dev), lab.to(dev)], schedule="graph", warmup=3)
xd, ld = x.to(dev), lab.to(dev)

def run(name, body):
    for _ in range(3):
        body(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); body(K); torch.cuda.synchronize()
    print(f"{name:54s} {(time.perf_counter() - t0) * 1e3 / K:7.2f} ms/step", flush=True)

def v_dev(n):
    for _ in range(n): step(xd, ld)
def v_dev_item(n):
    for _ in range(n): step(xd, ld).item()
def v_direct(n):
    for _ in range(n): step(x, lab).item()
def v_prefetch(n):
    step.prefetch(x, lab)
    for i in range(n):
        loss = step()
        if i + 1 < n: step.prefetch(x, lab)
        loss.item()
def v_prefetch_noitem(n):
    step.prefetch(x, lab)
    for i in range(n):
        loss = step()
        if i + 1 < n: step.prefetch(x, lab)
def v_prefetch_late(n):      # prefetch issued after the loss read-back (no overlap with the step; copy stream still used)
    step.prefetch(x, lab)
    for i in range(n):
        loss = step(); loss.item()
        if i + 1 < n: step.prefetch(x, lab)
def v_prefetch_async(n):      # loss read one step late through a pinned 4-byte copy + event
    step.prefetch(x, lab)
    pending = None
    for i in range(n):
        step(); h = step.loss_async()
        if i + 1 < n: step.prefetch(x, lab)
        if pending is not None: pending.get()
        pending = h
    pending.get()
def v_dev_async(n):           # device inputs, async loss
    pending = None
    for i in range(n):
        step(xd, ld); h = step.loss_async()
        if pending is not None: pending.get()
        pending = h
    pending.get()
def v_h2d_only(n):
    for _ in range(n):
        xd.copy_(x, non_blocking=True); ld.copy_(lab, non_blocking=True)
    torch.cuda.synchronize()

run("device inputs, no read-back", v_dev)
run("device inputs, loss.item() every step", v_dev_item)
run("pinned host inputs through step(...), .item()", v_direct)
run("prefetch on copy stream, .item()", v_prefetch)
run("prefetch on copy stream, no read-back", v_prefetch_noitem)
run("prefetch issued after .item()", v_prefetch_late)
run("prefetch on copy stream, loss_async one step late", v_prefetch_async)
run("device inputs, loss_async one step late", v_dev_async)
run("prefetch on copy stream, .item() (again)", v_prefetch)
run("device inputs, no read-back (again)", v_dev)
run("H2D copies alone (24 MB)", v_h2d_only)
