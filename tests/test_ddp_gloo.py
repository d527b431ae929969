"""CPU, world_size = 2, gloo: the N > 1 host path of the train step.

The data path has exactly one exchange — DDP's gradient all-reduce — so what can be checked without a GPU is the
host-side contract: B200UNet's parameter set survives DistributedDataParallel(find_unused_parameters=False) (every
parameter receives a gradient every step), the all-reduced gradients equal the mean of the per-rank gradients, ranks
stay in lock-step after an optimizer step, and bench.py's per-rank seeding / whole-job accounting.  The CUDA kernels
cannot run here, so a test-only subclass evaluates the same parameters through the oracle (tests may use the oracle;
the product has no CPU path)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = ["organ", "pancreatic_lesion"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_net():
    from oracle.unet_ref import synthetic_state_dict, unet_forward
    from rsuper_b200.unet import B200UNet

    class OracleBackedUNet(B200UNet):
        """Same module tree / parameters; forward through the oracle (TEST ONLY)."""

        def forward(self, x):
            return {"segmentation": unet_forward(x, dict(self.named_parameters()))}

    net = OracleBackedUNet(1, 8, num_classes=len(CLASSES))
    net.load_state_dict(synthetic_state_dict(8, len(CLASSES)))
    return net


def _rank_batch(rank):
    from oracle import synth
    from oracle.unet_ref import synthetic_image
    x = synthetic_image(1, 32, 32, 32, seed=1234 + rank)
    lab = synth.make_batch(["mask"], CLASSES, (32, 32, 32), seed=4321 + rank)["label"].long()
    return x, lab


def _loss(net, x, lab):
    from oracle import losses_ref as LR
    args = LR.default_args(report_volume_loss_basic=0.0)
    return LR.calculate_loss(net(x), lab, None, args, None, None, None, None, CLASSES)["overall"]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "r-super_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        net = _make_net()
        ddp = torch.nn.parallel.DistributedDataParallel(net, find_unused_parameters=False)
        opt = torch.optim.AdamW(net.parameters(), lr=6e-4, betas=(0.9, 0.999), weight_decay=0.05, eps=1e-5)
        x, lab = _rank_batch(rank)
        _loss(ddp, x, lab).backward()           # would raise on an unused parameter
        grads = {k: p.grad.clone() for k, p in net.named_parameters()}
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        torch.save({"grads": grads, "params": {k: p.detach().clone() for k, p in net.named_parameters()}},
                   os.path.join(out_dir, f"rank{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_ddp_world2_gradients_are_rank_means(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    # reference: per-rank gradients computed without DDP
    local = []
    for rank in range(world):
        net = _make_net()
        x, lab = _rank_batch(rank)
        _loss(net, x, lab).backward()
        local.append({k: p.grad.clone() for k, p in net.named_parameters()})
    assert len(r0["grads"]) == 45
    for k in r0["grads"]:
        mean = (local[0][k] + local[1][k]) / world       # DDP averages, like the reference (train_ddp.py:661-671)
        scale = mean.abs().max().item() + 1e-12
        # 1e-3: the per-rank reference runs with a different CPU thread count (fp32 re-association through 44 layers);
        # a missing average or a dropped rank would be off by O(1)
        assert (r0["grads"][k] - mean).abs().max().item() <= 1e-3 * scale, k
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k          # identical on every rank
        assert torch.equal(r0["params"][k], r1["params"][k]), k        # ranks stay in lock-step after the step


def _step_worker(rank, world, port, out_dir):
    """B200TrainStep(schedule='eager', process_group=gloo): the product's own train-step object — flat gradient buffer,
    one all-reduce, fused optimizer — with the engine running on the emulated kernels (tests/emul)."""
    for p in (ROOT, os.path.join(ROOT, "r-super_b200"), os.path.join(ROOT, "tests", "emul")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(2)
    import install
    install.install()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from oracle import losses_ref as LR
        from oracle.unet_ref import synthetic_state_dict
        from rsuper_b200 import losses
        from rsuper_b200.optim import B200AdamW
        from rsuper_b200.train_step import B200TrainStep
        from rsuper_b200.unet import B200UNet
        net = B200UNet(1, 8, num_classes=len(CLASSES), precision="fp32")
        net.load_state_dict(synthetic_state_dict(8, len(CLASSES)))
        params = list(net.parameters())
        ema = [p.detach().clone() for p in params]
        opt = B200AdamW(params, lr=6e-4, weight_decay=0.05, max_norm=1.0, ema_params=ema, capturable=True)
        args = LR.default_args(report_volume_loss_basic=0.0)
        args.nan_check = False
        x, lab = _rank_batch(rank)
        loss_fn = lambda out, lb: losses.calculate_loss(out, lb, None, args, None, None, None, None, CLASSES)["overall"]
        loss_fn(net(x), lab.to(torch.uint8)).backward()           # this rank's LOCAL gradient through the same engine
        local = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        step = B200TrainStep(net, loss_fn, opt, (x, lab.to(torch.uint8)), schedule="eager", process_group=dist.group.WORLD)
        assert all(p.grad.data_ptr() >= step.flat_grad.data_ptr() for p in params)       # every grad is a view of the flat buffer
        before = [p.detach().clone() for p in params]
        loss = step(x, lab.to(torch.uint8))
        torch.save({"local": local, "loss": loss.item(), "grad_norm": float(opt.last_grad_norm), "flat": step.flat_grad.clone(), "grads": [p.grad.clone() for p in params], "buckets": sorted(step.sink.done),
                    "params": [p.detach().clone() for p in params], "before": before, "ema": ema, "step": opt.global_step},
                   os.path.join(out_dir, f"step{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_train_step_world2_allreduces_the_flat_gradient(tmp_path):
    """N > 1 path of B200TrainStep on CPU (gloo, world 2): after one step both ranks hold the same averaged (then clipped)
    flat gradient, the same parameters and EMA; the average equals the mean of the two ranks' local gradients (oracle)."""
    world = 2
    mp.spawn(_step_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "step0.pt")
    r1 = torch.load(tmp_path / "step1.pt")
    assert r0["step"] == r1["step"] == 1 and r0["loss"] != r1["loss"]                    # different shards, same update
    assert torch.equal(r0["flat"], r1["flat"])
    for a, b in zip(r0["params"] + r0["ema"], r1["params"] + r1["ema"]):
        assert torch.equal(a, b)
    assert any(not torch.equal(a, b) for a, b in zip(r0["params"], r0["before"]))       # the step really updated the weights
    mean = [(a + b) / world for a, b in zip(r0["local"], r1["local"])]   # DDP averages, like the reference (train_ddp.py:661-671)
    norm = torch.sqrt(sum((g.double() ** 2).sum() for g in mean)).item()
    # (tolerances: two passes of the engine differ by the order of their fp32 atomics, amplified by this 2^3-bottom toy net;
    #  a missing average or a dropped rank is off by O(1))
    assert abs(r0["grad_norm"] - norm) <= 2e-2 * norm                                    # norm of the AVERAGED gradient, before clipping
    clip = min(1.0, 1.0 / (norm + 1e-6))
    # whole-vector comparison: single tensors of this toy net move by several per cent between two passes of the engine
    num = den = 0.0
    for g, got in zip(mean, r0["grads"]):                # p.grad = views of the flat buffer (gradient-production order)
        num += (got.double() - g.double() * clip).pow(2).sum().item()
        den += (g.double() * clip).pow(2).sum().item()
    assert r0["buckets"] == [0, 1, 2, 3]                 # the engine reported all four buckets: each went out as its own all-reduce
    assert (num / den) ** 0.5 <= 5e-2


def test_bench_job_accounting():
    """bench.py: per-rank seeds differ, the whole-job metric counts every rank's voxels (weak scaling)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.rank_seeds(0) != bench.rank_seeds(1)
    assert bench.job_voxels(world=8, batch=2, size=128) == 8 * 2 * 128 ** 3
    line = bench.mvox_per_s(bench.job_voxels(2, 2, 128), ms=33.0)
    assert abs(line - 2 * 2 * 128 ** 3 / 0.033 / 1e6) < 1e-6
