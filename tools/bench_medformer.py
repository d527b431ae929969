#!/usr/bin/env python
"""MedFormer train-step timing on one B200: B200MedFormer (yaml configuration, base 32) forward + calculate_loss on
[final, aux] + backward, against the same graph in stock PyTorch (the oracle restatement of the reference module, cuDNN / cuBLAS)
on the same GPU.  Prints one JSON line.  Not part of bench.py: the north-star metric is the UNet step."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "r-super_b200"))

import torch  # noqa: E402

FULL = dict(base_chan=32, chan_num=[64, 128, 256, 320, 256, 128, 64, 32], conv_num=[2, 0, 0, 0, 0, 0, 2, 2], trans_num=[0, 2, 4, 6, 4, 2, 0, 0],
            num_heads=[1, 4, 8, 10, 8, 4, 1, 1], map_size=[3, 3, 3], expansion=4, fusion_depth=2, fusion_dim=320, fusion_heads=10, aux_loss=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--side", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--schedule", default="graph", choices=["graph", "eager", "both"])
    ap.add_argument("--trace", default="")
    a = ap.parse_args()
    from rsuper_b200 import synthetic as synth
    from rsuper_b200 import losses, ops
    from rsuper_b200.medformer import B200MedFormer
    # under torchrun (WORLD_SIZE > 1): one rank per GPU, weak scaling, the flat-gradient all-reduce of B200TrainStep inside the graph
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
        a.no_torch, a.schedule = True, "graph"
    torch.manual_seed(0)
    c = FULL
    classes = ["organ", "pancreatic_lesion"]
    net = B200MedFormer(1, len(classes), base_chan=c["base_chan"], map_size=c["map_size"], conv_num=c["conv_num"], trans_num=c["trans_num"],
                        chan_num=c["chan_num"], num_heads=c["num_heads"], fusion_depth=c["fusion_depth"], fusion_dim=c["fusion_dim"],
                        fusion_heads=c["fusion_heads"], expansion=c["expansion"], aux_loss=True, precision=a.precision).to(dev)
    S = a.side
    x = synth.synthetic_image(a.batch, S, S, S, seed=3 + rank, device=dev)
    batch = synth.make_batch(["mask"] * a.batch, classes, (S, S, S), seed=5 + rank, device=dev)
    args = synth.default_loss_args(report_volume_loss_basic=0.0)
    args.nan_check = False

    def step_ours():
        for p in net.parameters():
            p.grad = None
        out = net(x)
        loss = losses.calculate_loss(out, batch["label"], None, args, None, None, None, None, classes)["overall"]
        loss.backward()
        return loss

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps, float(loss)

    vox = a.batch * S ** 3
    res = {"model": "MedFormer base 32 (yaml configuration)", "batch": a.batch, "side": S, "precision": a.precision}
    if a.schedule in ("eager", "both"):
        before = ops.LAUNCHES
        step_ours()
        launches = ops.LAUNCHES - before
        ms, loss = timed(step_ours)
        res["eager_fwd_loss_bwd"] = {"ms_per_step": round(ms, 2), "mvox_per_s": round(vox / ms / 1e3, 2), "loss": loss, "gpu_launches": launches,
                                     "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    if a.trace:
        # one eager step with every launch of ours bracketed by CUDA events: where the GPU time of our kernels goes
        step_ours()
        torch.cuda.synchronize()
        ops.PROFILE = []
        step_ours()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for fam, flops, e0, e1, what in prof:
            key = what.split(" ")[0] + (" 1x1x1" if "1x1x1" in what else "")
            t = agg.setdefault(key, [0, 0.0, 0.0])
            t[0] += 1; t[1] += e0.elapsed_time(e1); t[2] += flops
        tot = sum(v[1] for v in agg.values())
        with open(a.trace, "w") as f:
            f.write(f"# B200MedFormer batch {a.batch} x {S}^3 {a.precision}: one eager step, every launch of librsuper_b200 bracketed by CUDA events (torch map-side ops not included)\n")
            f.write(f"# launches {len(prof)}  total {tot:.2f} ms\n")
            for k, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{k:36s} {n:5d} launches {t:8.3f} ms {t / tot:6.3f}" + (f" {fl / t / 1e9:8.1f} TF/s" if fl else "") + "\n")
            f.write("# the 40 longest launches\n")
            for fam, flops, e0, e1, what in sorted(prof, key=lambda r: -r[2].elapsed_time(r[3]))[:40]:
                f.write(f"{e0.elapsed_time(e1) * 1e3:9.1f} us  {what}\n")
    if a.schedule in ("graph", "both"):
        # the whole step (forward, loss, backward, clip + AdamW + EMA) as one CUDA graph: B200TrainStep is model-agnostic
        from rsuper_b200.optim import B200AdamW
        from rsuper_b200.train_step import B200TrainStep
        params = list(net.parameters())
        ema = [p.detach().clone() for p in params]
        opt = B200AdamW(params, lr=1e-4, weight_decay=0.05, max_norm=1.0, ema_params=ema, capturable=True)

        def loss_fn(out, lab):
            return losses.calculate_loss(out, lab, None, args, None, None, None, None, classes)["overall"]
        torch.cuda.reset_peak_memory_stats()
        step = B200TrainStep(net, loss_fn, opt, [x, batch["label"]], schedule="graph", warmup=2, process_group=pg)
        ms, loss = timed(lambda: step(x, batch["label"]))
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)          # a step takes as long as the slowest rank
            ms = float(t.item())
            vox *= world
            res["n_gpus"] = world
        res["graph_full_step"] = {"ms_per_step": round(ms, 2), "mvox_per_s": round(vox / ms / 1e3, 2), "loss": loss,
                                  "gpu_launches": step.launches_per_step, "includes": "clip + AdamW + EMA",
                                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    if not a.no_torch:
        # the stock-PyTorch arm: the oracle restatement of the reference module (test infrastructure, only imported here)
        from oracle import losses_ref as LR
        from oracle.medformer_ref import medformer_forward
        torch.manual_seed(0)
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters()}
        for label, ctx in (("torch_fp32", None), ("torch_bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            def step_torch():
                for v in sd.values():
                    v.grad = None
                if ctx is None:
                    out = medformer_forward(x, sd, c)
                else:
                    with ctx:
                        out = medformer_forward(x, sd, c)
                    out = {"segmentation": [t.float() for t in out["segmentation"]]}
                loss = LR.calculate_loss(out, batch["label"].long(), None, args, None, None, None, None, classes)["overall"]
                loss.backward()
                return loss
            try:
                torch.cuda.reset_peak_memory_stats()
                t_ms, t_loss = timed(step_torch)
                res[label] = {"ms_per_step": round(t_ms, 2), "mvox_per_s": round(vox / t_ms / 1e3, 2), "loss": t_loss,
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
            except Exception as e:  # noqa: BLE001
                res[label] = {"error": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        # same teardown as bench.py: a captured NCCL collective keeps the communicator busy in destroy_process_group
        del step
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
