"""GPU tests of the rows SURVEY §8(f) widens into (N2 batch assembly from bit-packed masks, N3 sliding-window inference +
connected components, N4 fused clip/AdamW/EMA) and the FULL-SIZE property tests of the hot path at BASELINE.json
configs[1] (2 x 128^3, base 32).

Everything here calls through the C-ABI and is checked against the oracle (oracle/*.py), the committed golden vectors of
the REAL reference (tests/golden) or plain torch.  The file sorts after the established suites on purpose: its kernels
(csrc/train_glue.cu, csrc/infer.cu) and its full-size tolerances were written after round 1's GPU budget was spent, so the
whole file is marked `staged` (tests/conftest.py: non-strict xfail until the first run on a B200 — an XPASS in the report
is that first run; the marker comes off once it has been seen green).
"""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


# ------------------------------------------------------------------------------------------------------------------
# N4: clip_grad_norm_ + AdamW + EMA in two launches
# ------------------------------------------------------------------------------------------------------------------
def _param_set(dev, seed):
    """Tensors of awkward sizes: several 4096-element chunks + ragged tail, tiny, 2-D, and one whose storage is offset by
    one float (not 16-byte aligned: the scalar path, as for DDP bucket views)."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(3 * 4096 + 5,), (1000,), (37, 5), (8197,), (4096,), (1,)]
    out = []
    for i, s in enumerate(shapes):
        n = int(np.prod(s))
        buf = torch.randn(n + 1, generator=g).to(dev)
        out.append(buf[1:1 + n].view(s) if i == 3 else buf[:n].view(s).clone())
    return out


@pytest.mark.parametrize("with_ema,max_norm", [(True, 1.0), (False, 1.0), (True, None)])
def test_fused_clip_adamw_ema_matches_torch(cuda_dev, with_ema, max_norm):
    """Four steps (large gradients: clipping active; small ones: inactive) against torch.nn.utils.clip_grad_norm_ +
    torch.optim.AdamW(eps=1e-5) + update_ema_variables' formula (train_ddp.py:352-357, training/utils.py:46-51,154-158)."""
    from rsuper_b200.optim import B200AdamW
    init = _param_set(cuda_dev, 0)
    ours = [torch.nn.Parameter(t) for t in init]                      # keeps the misaligned view
    ref = [torch.nn.Parameter(t.clone()) for t in init]
    assert ours[3].data_ptr() % 16 != 0
    ema_o = [p.detach().clone() for p in ours] if with_ema else None
    ema_r = [p.detach().clone() for p in ref]
    hyper = dict(lr=6e-4, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.05)
    opt_o = B200AdamW(ours, max_norm=max_norm, ema_params=ema_o, ema_alpha=0.99, **hyper)
    opt_r = torch.optim.AdamW(ref, foreach=False, fused=False, **hyper)
    for step in range(4):
        grads = [t * (3.0 if step < 2 else 1e-3) for t in _param_set(cuda_dev, 10 + step)]
        for p, q, g in zip(ours, ref, grads):
            p.grad = g.clone() if p.data_ptr() % 16 == 0 else torch.cat([g.new_zeros(1), g.reshape(-1)])[1:].view(g.shape)
            q.grad = g.clone()
        norm_r = torch.nn.utils.clip_grad_norm_(ref, max_norm) if max_norm is not None else None
        opt_r.step()
        alpha = min(1 - 1 / (step + 1), 0.99)
        with torch.no_grad():
            for e, q in zip(ema_r, ref):
                e.mul_(alpha).add_(q.detach(), alpha=1 - alpha)
        opt_o.step()
        if cuda_dev.type == "cuda":
            torch.cuda.synchronize()
        if max_norm is not None:
            torch.testing.assert_close(opt_o.last_grad_norm.reshape(()), norm_r.reshape(()), rtol=1e-5, atol=0)
        for i, (p, q) in enumerate(zip(ours, ref)):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-6, atol=2e-7, msg=lambda m: f"step {step} param {i}: {m}")
            torch.testing.assert_close(p.grad, q.grad, rtol=2e-6, atol=1e-8, msg=lambda m: f"step {step} grad {i}: {m}")
            so, sr = opt_o.state[p], opt_r.state[q]
            # exp_avg = m + (g - m) * 0.1 cancels where g changes sign: absolute rounding noise ~ eps * |g| (|g| up to ~10 unclipped)
            torch.testing.assert_close(so["exp_avg"], sr["exp_avg"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(so["exp_avg_sq"], sr["exp_avg_sq"], rtol=1e-5, atol=1e-10)
            assert float(so["step"]) == float(sr["step"]) == step + 1
            if with_ema:
                torch.testing.assert_close(ema_o[i], ema_r[i], rtol=2e-6, atol=2e-7)
    # checkpoints are interchangeable with torch's AdamW (same per-parameter state keys)
    assert set(opt_o.state_dict()["state"][0].keys()) == set(opt_r.state_dict()["state"][0].keys())


def test_fused_clip_adamw_ema_matches_reference_golden(cuda_dev):
    """Same four steps the REAL reference's get_optimizer / update_ema_variables ran for tests/golden/reference_train_glue.npz
    (inputs regenerated from oracle/train_glue_ref.glue_inputs)."""
    import os
    from oracle.train_glue_ref import glue_inputs
    from rsuper_b200.optim import B200AdamW
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_train_glue.npz"))
    params = [torch.nn.Parameter(t) for t in glue_inputs(-1, device=cuda_dev)]
    ema = [t.clone() for t in glue_inputs(-1, device=cuda_dev)]
    opt = B200AdamW(params, lr=6e-4, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.05, max_norm=1.0, ema_params=ema, ema_alpha=0.99)
    for step in range(4):
        opt.zero_grad()
        for p, g in zip(params, glue_inputs(step, device=cuda_dev)):
            p.grad = g.clone()
        opt.step()
        assert abs(opt.last_grad_norm.item() - float(gold[f"norm_{step}"])) <= 2e-6 * float(gold[f"norm_{step}"])
        if f"p_{step}_0" not in gold:
            continue
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f"p_{step}_{i}"], rtol=2e-6, atol=2e-7)
            np.testing.assert_allclose(ema[i].cpu().numpy(), gold[f"ema_{step}_{i}"], rtol=2e-6, atol=2e-7)
            np.testing.assert_allclose(p.grad.cpu().numpy(), gold[f"g_{step}_{i}"], rtol=2e-6, atol=1e-9)
    for i, p in enumerate(params):
        np.testing.assert_allclose(opt.state[p]["exp_avg"].cpu().numpy(), gold[f"exp_avg_{i}"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(opt.state[p]["exp_avg_sq"].cpu().numpy(), gold[f"exp_avg_sq_{i}"], rtol=1e-5, atol=1e-12)


def test_fused_optimizer_resume_and_gradless_parameters(cuda_dev):
    """(1) state_dict() -> fresh B200AdamW.load_state_dict() resumes bit-identically: the EMA warm-up alpha
    min(1 - 1/(global_step+1), ema_alpha) follows the persisted loop step (training/utils.py:156) instead of restarting at 0.
    (2) A checkpoint of torch's AdamW (what the reference writes) loads, global_step derived from the per-parameter steps.
    (3) A parameter without a gradient is skipped by AdamW like in torch but its EMA copy still moves
    (update_ema_variables updates every parameter), checked against oracle/train_glue_ref on the other rows."""
    from oracle.train_glue_ref import clip_adamw_ema_step
    from rsuper_b200.optim import B200AdamW
    hyper = dict(lr=6e-4, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.05)

    def make():
        ps = [torch.nn.Parameter(t) for t in _param_set(cuda_dev, 0)]
        ema = [p.detach().clone() + 0.25 for p in ps]
        return ps, ema, B200AdamW(ps, max_norm=1.0, ema_params=ema, ema_alpha=0.99, **hyper)

    def feed(ps, step, skip=()):
        for i, (p, g) in enumerate(zip(ps, _param_set(cuda_dev, 30 + step))):
            p.grad = None if i in skip else (g.clone() if p.data_ptr() % 16 == 0 else
                                             torch.cat([g.new_zeros(1), g.reshape(-1)])[1:].view(g.shape))

    # straight run of 4 steps
    ps_a, ema_a, opt_a = make()
    for step in range(4):
        feed(ps_a, step); opt_a.step()
    # 2 steps, checkpoint, fresh objects, 2 more steps
    ps_b, ema_b, opt_b = make()
    for step in range(2):
        feed(ps_b, step); opt_b.step()
    ckpt = opt_b.state_dict()
    assert ckpt["b200_global_step"] == 2
    ps_c = [torch.nn.Parameter(p.detach().clone()) for p in ps_b]
    ema_c = [e.clone() for e in ema_b]
    opt_c = B200AdamW(ps_c, max_norm=1.0, ema_params=ema_c, ema_alpha=0.99, **hyper)
    opt_c.load_state_dict(ckpt)
    assert opt_c.global_step == 2
    for step in range(2, 4):
        feed(ps_c, step); opt_c.step()
    for a, c in zip(ps_a + ema_a, ps_c + ema_c):
        assert torch.equal(a.detach(), c.detach())
    # torch AdamW checkpoint (no b200_global_step key): the loop step is recovered from the parameter steps
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps_b]
    opt_t = torch.optim.AdamW(ref, foreach=False, fused=False, **hyper)
    for step in range(3):
        feed(ref, step); opt_t.step()
    opt_d = B200AdamW([torch.nn.Parameter(p.detach().clone()) for p in ref], max_norm=1.0, **hyper)
    opt_d.load_state_dict(opt_t.state_dict())
    assert opt_d.global_step == 3
    # explicit loop step like the reference's update_ema_variables(net, ema_net, alpha, step)
    ps_e, ema_e, opt_e = make()
    feed(ps_e, 0); opt_e.step(global_step=500)
    assert opt_e.global_step == 501

    # gradient-less parameter: rows 1 and 4 have grad None at step 1
    ps_g, ema_g, opt_g = make()
    feed(ps_g, 0); opt_g.step()
    before_p = [p.detach().clone() for p in ps_g]
    before_e = [e.clone() for e in ema_g]
    feed(ps_g, 1, skip=(1, 4)); opt_g.step()
    alpha = min(1 - 1 / 2, 0.99)
    for i in (1, 4):
        assert torch.equal(ps_g[i].detach(), before_p[i])                       # AdamW left it alone
        torch.testing.assert_close(ema_g[i], alpha * before_e[i] + (1 - alpha) * before_p[i], rtol=1e-6, atol=1e-7)
        assert float(opt_g.state[ps_g[i]]["step"]) == 1
    # the rows that did get a gradient follow the oracle (norm over those rows only, like clip_grad_norm_)
    ps_o, ema_o, opt_o = make()
    feed(ps_o, 0); opt_o.step()
    keep = [0, 2, 3, 5]
    P = [ps_o[i].detach().clone().contiguous() for i in keep]
    G = [g.clone() for i, g in enumerate(_param_set(cuda_dev, 31)) if i in keep]
    M = [opt_o.state[ps_o[i]]["exp_avg"].clone() for i in keep]
    V = [opt_o.state[ps_o[i]]["exp_avg_sq"].clone() for i in keep]
    E = [ema_o[i].clone().contiguous() for i in keep]
    clip_adamw_ema_step(P, G, M, V, E, step=2, global_step=1, max_norm=1.0, ema_alpha=0.99, **hyper)
    for j, i in enumerate(keep):
        torch.testing.assert_close(ps_g[i].detach(), P[j], rtol=2e-6, atol=2e-7)
        torch.testing.assert_close(ema_g[i], E[j], rtol=2e-6, atol=2e-7)


def test_fused_optimizer_rejects_cpu_parameters():
    from rsuper_b200.optim import B200AdamW
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        B200AdamW([p], max_norm=1.0).step()


# ------------------------------------------------------------------------------------------------------------------
# N2: bit-packed masks -> device uint8 masks
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,shape", [(2, (16, 16, 16)), (3, (8, 12, 20)), (8, (4, 6, 10)), (9, (5, 7, 9)), (42, (6, 5, 7))])
def test_unpack_masks_bit_exact(cuda_dev, C, shape):
    from oracle import synth
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(C)
    masks = (torch.rand((2, C) + shape, generator=g) < 0.3).to(torch.uint8)
    packed = np.stack([synth.pack_masks(masks[b]) for b in range(2)])          # the reference's storage format
    assert packed.shape == (2, (C + 7) // 8) + shape
    want = torch.stack([synth.unpack_masks(packed[b], C) for b in range(2)])
    assert torch.equal(want, masks)
    dev_packed = torch.from_numpy(packed).to(cuda_dev)
    assert torch.equal(ops.unpack_masks(dev_packed, C).cpu(), want)
    assert torch.equal(ops.unpack_masks(dev_packed, C, invert=True).cpu(), 1 - want)
    with pytest.raises(ValueError):
        ops.unpack_masks(dev_packed, C + 8)


def test_assemble_batch_feeds_calculate_loss(cuda_dev):
    """Batch assembled from the packed on-disk format == the batch uploaded as plain tensors, down to the loss values."""
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import batch as B
    from rsuper_b200 import losses
    classes = ["organ", "pancreatic_lesion", "veins"]
    shape = (32, 32, 48)
    kinds = ["mask", "report"]
    ref = synth.make_batch(kinds, classes, shape, seed=5)
    got = B.assemble_batch(
        images=[ref["image"][b, 0].numpy() for b in range(2)],
        labels_packed=[synth.pack_masks(ref["label"][b]) for b in range(2)], num_classes=len(classes), device=cuda_dev,
        unk_packed=[None, synth.pack_masks(ref["unk_channels"][1])],
        chosen_packed=[None, synth.pack_masks(ref["mask"][1])],
        volumes=[None, ref["volumes"][1].numpy()], diameters=[None, ref["diameters"][1].numpy()])
    for k in ("image", "label", "unk_channels", "mask", "volumes", "diameters"):
        assert got[k].device.type == cuda_dev.type and tuple(got[k].shape) == tuple(ref[k].shape), k
        assert torch.equal(got[k].cpu().float(), ref[k].float()), k
    assert got["label"].dtype == torch.uint8
    logits = synth.synthetic_logits(2, len(classes), shape, seed=2, device=cuda_dev)
    args = LR.default_args()
    res = []
    for bt in (got, {k: v.to(cuda_dev) for k, v in ref.items()}):
        lg = logits.clone().requires_grad_(True)
        out = losses.calculate_loss({"segmentation": lg}, bt["label"], bt["unk_channels"], args, None, bt["mask"], bt["volumes"],
                                    bt["diameters"], classes, input_tensor=bt["image"])
        res.append({k: v.item() for k, v in out.items()})
    assert res[0].keys() == res[1].keys()
    for k in res[0]:
        assert abs(res[0][k] - res[1][k]) <= 1e-6 * max(1.0, abs(res[1][k])), (k, res[0][k], res[1][k])


# ------------------------------------------------------------------------------------------------------------------
# N3: sliding-window inference, organ gating, connected components
# ------------------------------------------------------------------------------------------------------------------
def _sliding_net(dev):
    """The tiny deterministic net tests/golden/make_golden.py ran through the REAL inference_sliding_window."""
    w = torch.frac(torch.sin(torch.arange(3 * 27, dtype=torch.float64) * 12.9898) * 43758.5453).float().reshape(3, 1, 3, 3, 3) - 0.5
    net = torch.nn.Conv3d(1, 3, 3, padding=1, bias=True)
    with torch.no_grad():
        net.weight.copy_(w)
        net.bias.copy_(torch.tensor([0.1, -0.2, 0.05]))
    return net.to(dev).eval()


def test_sliding_window_blend_matches_reference_golden(cuda_dev, golden):
    """Window placement, padding of small volumes, gated windows, sigmoid + mean blend — against outputs recorded from the
    REAL inference3d.inference_sliding_window (same net, same volumes), and against the oracle on the full volume."""
    from types import SimpleNamespace
    from oracle.inference_ref import inference_sliding_window as oracle_sw
    from oracle.unet_ref import synthetic_image
    from rsuper_b200.inference import inference_sliding_window
    net = _sliding_net(cuda_dev)
    gate = torch.zeros(1, 1, 40, 48, 56)
    gate[:, :, 4:20, 8:30, 10:20] = 1
    cases = [("ragged", (40, 48, 56), (16, 16, 32), None), ("small", (12, 20, 16), (16, 16, 32), None),
             ("exact", (32, 32, 32), (16, 16, 16), None), ("gated", (40, 48, 56), (16, 16, 32), gate)]
    for tag, shp, win, g in cases:
        vol = synthetic_image(1, *shp, seed=9)
        a = SimpleNamespace(window_size=list(win), classes=3)
        out = inference_sliding_window(net, vol.to(cuda_dev), a, pancreas=None if g is None else g.to(cuda_dev))
        assert not out.is_cuda and out.shape == (1, 3) + shp                 # CPU tensor, like the reference returns
        np.testing.assert_allclose(out.numpy()[:, :, 1::3, 1::3, 1::3], golden[f"sliding_{tag}"], rtol=0, atol=2e-6)
        assert abs(out.double().sum().item() - float(golden[f"sliding_{tag}_sum"])) <= 2e-6 * abs(float(golden[f"sliding_{tag}_sum"]))
        want = oracle_sw(_sliding_net("cpu"), vol, win, 3, gate=g)
        assert (out - want).abs().max().item() <= 2e-6
        prob, mask = inference_sliding_window(net, vol.to(cuda_dev), a, pancreas=None if g is None else g.to(cuda_dev),
                                              keep_on_device=True, threshold=0.5)
        assert prob.device.type == cuda_dev.type and mask.dtype == torch.uint8 and torch.equal(mask.bool(), prob > 0.5)
        resolved = (want - 0.5).abs() > 4e-6                                  # threshold decisions the tolerance resolves
        assert torch.equal(mask.cpu().bool()[resolved], (want > 0.5)[resolved])


def test_sliding_window_with_b200_unet(cuda_dev):
    """The windows go through B200UNet (eval, no grad): blended probabilities vs the oracle UNet blended by the oracle."""
    from types import SimpleNamespace
    from oracle.inference_ref import inference_sliding_window as oracle_sw
    from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
    from rsuper_b200.inference import inference_sliding_window
    from rsuper_b200.unet import B200UNet
    net = B200UNet(1, 8, num_classes=2, precision="fp32").to(cuda_dev)
    sd = synthetic_state_dict(8, 2, device=cuda_dev)
    net.load_state_dict(sd)
    vol = synthetic_image(1, 64, 96, 64, seed=4, device=cuda_dev)
    a = SimpleNamespace(window_size=[64, 64, 64], classes=2)
    got = inference_sliding_window(net, vol, a, keep_on_device=True)
    want = oracle_sw(lambda x: unet_forward(x.to(cuda_dev), sd), vol.cpu(), (64, 64, 64), 2)
    err = (got.cpu() - want).abs().max().item()
    print(f"[sliding/unet] max abs prob error {err:.3e}")
    assert err <= 2e-3                                                        # logits within 1e-3 relative -> probabilities


def _np_cc(mask_np):
    from oracle.inference_ref import connected_components
    return connected_components(mask_np)


@pytest.mark.parametrize("case", ["hand", "random_sparse", "random_dense", "blobs", "empty", "full", "snake"])
def test_connected_components_bit_exact(cuda_dev, case):
    """Component COUNT, the partition itself (raster-order numbering) and keep_largest_component — against the oracle
    (scipy face connectivity == SimpleITK's default)."""
    from oracle import synth
    from oracle.inference_ref import keep_largest_component as oracle_largest
    from rsuper_b200 import ops
    g = torch.Generator().manual_seed(7)
    if case == "hand":
        m = np.zeros((6, 6, 6), dtype=np.uint8)
        m[0, 0, 0] = 1; m[1, 1, 1] = 1; m[3:5, 3:5, 3:5] = 1; m[3, 3, 5] = 1
    elif case == "random_sparse":
        m = (torch.rand((24, 20, 28), generator=g) < 0.12).numpy().astype(np.uint8)
    elif case == "random_dense":
        m = (torch.rand((17, 23, 19), generator=g) < 0.45).numpy().astype(np.uint8)     # near the percolation threshold: long merges
    elif case == "blobs":
        m = synth.make_batch(["mask"], ["organ", "a_lesion", "b_lesion"], (48, 40, 56), seed=21)["label"][0].amax(0).numpy().astype(np.uint8)
    elif case == "empty":
        m = np.zeros((5, 6, 7), dtype=np.uint8)
    elif case == "full":
        m = np.ones((9, 8, 7), dtype=np.uint8)
    else:  # one long serpentine component: the worst case for pointer chains
        m = np.zeros((1, 32, 33), dtype=np.uint8)
        m[0, ::2, :] = 1
        m[0, 1::4, -1] = 1
        m[0, 3::4, 0] = 1
    want_labels, want_n = _np_cc(m)
    labels, n, largest = ops.cc_label(torch.from_numpy(m).to(cuda_dev), keep_largest=True)
    assert int(n.item()) == want_n
    lab = labels.cpu().numpy()
    assert np.array_equal(lab >= 0, m > 0)
    roots = np.unique(lab[lab >= 0])                       # ascending root index == raster order of first voxels
    assert len(roots) == want_n
    dense = np.zeros_like(lab)
    dense[lab >= 0] = np.searchsorted(roots, lab[lab >= 0]) + 1
    assert np.array_equal(dense, want_labels)              # same numbering as scipy.ndimage.label / SimpleITK
    assert np.array_equal(largest.cpu().numpy().astype(bool), oracle_largest(m))
    labels2, n2, none = ops.cc_label(torch.from_numpy(m).to(cuda_dev))
    assert none is None and int(n2.item()) == want_n and torch.equal(labels2, labels)


def test_organ_gating_matches_oracle(cuda_dev):
    from types import SimpleNamespace
    from oracle.inference_ref import gate_lesion_by_organ
    from rsuper_b200.inference import postprocess_npz
    g = torch.Generator().manual_seed(3)
    classes = ["pancreas", "pancreatic_lesion", "kidney_right", "kidney_left", "kidney_lesion"]
    prob = torch.rand((1, 5, 12, 14, 16), generator=g)
    prob[0, 0] *= (torch.rand((12, 14, 16), generator=g) < 0.05)        # sparse organs: the dilation matters
    prob[0, 2] *= (torch.rand((12, 14, 16), generator=g) < 0.03)
    prob[0, 3] *= (torch.rand((12, 14, 16), generator=g) < 0.03)
    out = postprocess_npz(prob.to(cuda_dev), classes, SimpleNamespace(organ_mask_on_lesion=True))
    p = prob[0].numpy()
    assert np.array_equal(out["pancreatic_lesion"].cpu().numpy(), gate_lesion_by_organ(p[1], p[0]))
    assert np.array_equal(out["kidney_lesion"].cpu().numpy(), gate_lesion_by_organ(p[4], p[2] + p[3]))
    assert np.array_equal(out["pancreas"].cpu().numpy(), p[0])
    plain = postprocess_npz(prob.to(cuda_dev), classes, SimpleNamespace(organ_mask_on_lesion=False))
    assert np.array_equal(plain["kidney_lesion"].cpu().numpy(), p[4])


# ------------------------------------------------------------------------------------------------------------------
# Full size (BASELINE.json configs[1]: 2 x 128^3, base 32): size-independent properties of the hot path
# ------------------------------------------------------------------------------------------------------------------
def _full_net(cuda_dev, precision):
    from oracle.unet_ref import synthetic_state_dict
    from rsuper_b200.unet import B200UNet
    net = B200UNet(1, 32, num_classes=2, precision=precision).to(cuda_dev)
    sd = synthetic_state_dict(32, 2, device=cuda_dev)
    net.load_state_dict(sd)
    return net, sd


def test_full_size_samples_are_independent_and_runs_repeat(cuda_dev):
    """InstanceNorm is per sample and nothing else couples the batch: a sample's logits in a batch of two equal
    its logits alone; a second run of the same batch reproduces the first (only the order of the fp32 statistics atomics
    differs).  fp32 parity mode, 128^3."""
    from oracle.unet_ref import synthetic_image
    net, _ = _full_net(cuda_dev, "fp32")
    x = synthetic_image(2, 128, 128, 128, seed=5, device=cuda_dev)
    with torch.no_grad():
        both = net(x)["segmentation"].clone()
        again = net(x)["segmentation"].clone()
        solo = net(x[1:2].contiguous())["segmentation"].clone()
    assert torch.isfinite(both).all()
    e_rep, e_solo = rel(again, both), rel(solo, both[1:2])
    print(f"[full-size] repeat rel {e_rep:.2e}, batch-of-2 vs alone rel {e_solo:.2e}")
    assert e_rep <= 1e-3 and e_solo <= 1e-3


def test_full_size_logits_mask_and_component_count_vs_oracle(cuda_dev):
    """North-star bars at full size, against the oracle run in fp32 on the same GPU (cuDNN, TF32 off): logits within 1e-3
    relative; the argmax mask identical wherever the oracle's own margin exceeds that tolerance; and — when the masks are
    identical — the face-connected component count of the lesion mask identical (it is a function of the mask)."""
    from oracle.inference_ref import connected_components
    from oracle.unet_ref import synthetic_image, unet_forward
    net, sd = _full_net(cuda_dev, "fp32")
    for side in (32, 64):                      # growth of the error with the depth of reduction, on record
        xs = synthetic_image(1, side, side, side, seed=6, device=cuda_dev)
        with torch.no_grad():
            print(f"[full-size] parity mode at {side}^3: logits rel-to-max {rel(net(xs)['segmentation'], unet_forward(xs, sd)):.3e}")
    x = synthetic_image(1, 128, 128, 128, seed=6, device=cuda_dev)
    with torch.no_grad():
        got = net(x)["segmentation"]
        want = unet_forward(x, sd)
    e = rel(got, want)
    tol = 1e-3 * want.abs().max().item()
    margin = (want[:, 0] - want[:, 1]).abs()
    differ = got.argmax(1) != want.argmax(1)
    print(f"[full-size] logits rel-to-max {e:.3e}; argmax differs at {int(differ.sum())} of {differ.numel()} voxels, "
          f"{int((differ & (margin > 2 * tol)).sum())} of them outside the tolerance band")
    assert e <= 1e-3
    assert int((differ & (margin > 2 * tol)).sum()) == 0
    if int(differ.sum()) == 0:
        a = connected_components(got.argmax(1)[0].cpu().numpy())[1]
        b = connected_components(want.argmax(1)[0].cpu().numpy())[1]
        assert a == b


def test_bf16_mode_error_growth_vs_emulating_oracle(cuda_dev):
    """The benchmarked mode (bf16 operands + bf16 activation storage) at 32^3 / 64^3 / 128^3 and at the benchmark shape
    2 x 128^3, base 32: its logits error against the fp32 oracle (cuDNN, TF32 off) may not exceed twice the error of the
    oracle evaluated with bf16 rounding at the same points (`emulate=True, storage='bf16'`) — the kernels add no error
    of their own at any depth of reduction — and the argmax mask agrees with the fp32 oracle's at least as often."""
    from oracle.unet_ref import synthetic_image, unet_forward
    net, sd = _full_net(cuda_dev, "bf16")
    for n, side in ((1, 32), (1, 64), (1, 128), (2, 128)):
        x = synthetic_image(n, side, side, side, seed=6, device=cuda_dev)
        with torch.no_grad():
            got = net(x)["segmentation"]
            want = unet_forward(x, sd)
            emul = unet_forward(x, sd, emulate=True, storage="bf16")
        e, e_emul = rel(got, want), rel(emul, want)
        agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
        agree_emul = (emul.argmax(1) == want.argmax(1)).float().mean().item()
        print(f"[bf16 growth] {n} x {side}^3: kernels vs fp32 oracle {e:.3e} (argmax {agree:.6f}) | bf16-emulating oracle vs fp32 oracle "
              f"{e_emul:.3e} (argmax {agree_emul:.6f}) | kernels vs emulating oracle {rel(got, emul):.3e}")
        assert torch.isfinite(got).all()
        if side >= 64:
            # (32^3 is on record only: its bottom level normalises over 2^3 voxels, where the order of the fp32 statistics
            #  atomics alone moves the logits by 5e-2 .. 3e-1 from run to run — the oracle's own bf16 error there is 1e-1)
            assert e <= 2.0 * e_emul + 1e-3
            assert agree >= agree_emul - 2e-3
        del got, want, emul


def test_full_size_seg_loss_vs_oracle_at_identical_logits(cuda_dev):
    """Loss within 1e-5 and d(logits) within 1e-4 of the oracle at 2 x 2 x 128^3 (the reductions run over 4.2 M voxels)."""
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import losses
    classes = ["organ", "pancreatic_lesion"]
    lab = synth.make_batch(["mask", "mask"], classes, (128, 128, 128), seed=8, device=cuda_dev)["label"]
    logits = synth.synthetic_logits(2, 2, (128, 128, 128), seed=1, device=cuda_dev)
    a = logits.clone().requires_grad_(True)
    b = logits.clone().requires_grad_(True)
    args = LR.default_args(report_volume_loss_basic=0.0)
    ours = losses.calculate_loss({"segmentation": a}, lab, None, args, None, None, None, None, classes)["overall"]
    ref = LR.seg_loss(b, lab.float(), torch.ones_like(b))
    ours.backward(); ref.backward()
    dl = abs(ours.item() - ref.item()) / abs(ref.item())
    dg = rel(a.grad, b.grad)
    print(f"[full-size] seg loss {ours.item():.7f} vs oracle {ref.item():.7f} (rel {dl:.2e}); dlogits rel-to-max {dg:.2e}")
    assert dl <= 1e-5 and dg <= 1e-4


def test_full_size_train_steps_with_fused_optimizer(cuda_dev):
    """Three full-size bf16 train steps (forward, loss, backward, B200AdamW = clip + AdamW + EMA) against the same steps
    with the stock torch glue (clip_grad_norm_, fused AdamW, foreach EMA) from identical initial state: identical losses
    at step 0, parameters and EMA within optimizer rounding after the steps, finite everywhere, loss not increasing."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image
    from rsuper_b200 import losses
    from rsuper_b200.optim import B200AdamW
    classes = ["organ", "pancreatic_lesion"]
    x = synthetic_image(2, 128, 128, 128, seed=5, device=cuda_dev)
    lab = synth.make_batch(["mask", "mask"], classes, (128, 128, 128), seed=8, device=cuda_dev)["label"]
    args = LR.default_args(report_volume_loss_basic=0.0)
    hyper = dict(lr=6e-4, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.05)
    runs = []
    for fused in (True, False):
        net, _ = _full_net(cuda_dev, "bf16")
        params = list(net.parameters())
        ema = [p.detach().clone() for p in params]
        opt = B200AdamW(params, max_norm=1.0, ema_params=ema, ema_alpha=0.99, **hyper) if fused else \
            torch.optim.AdamW(params, fused=True, **hyper)
        ls = []
        for step in range(3):
            opt.zero_grad(set_to_none=True)
            loss = losses.calculate_loss(net(x), lab, None, args, None, None, None, None, classes)["overall"]
            loss.backward()
            if not fused:
                torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            if not fused:
                alpha = min(1 - 1 / (step + 1), 0.99)
                with torch.no_grad():
                    torch._foreach_mul_(ema, alpha)
                    torch._foreach_add_(ema, [p.detach() for p in params], alpha=1 - alpha)
            ls.append(loss.item())
        runs.append((ls, [p.detach().clone() for p in params], ema))
        del net, opt
    (l_f, p_f, e_f), (l_s, p_s, e_s) = runs
    print(f"[full-size] losses fused {l_f} stock {l_s}")
    assert all(np.isfinite(l_f)) and abs(l_f[0] - l_s[0]) <= 2e-3 * abs(l_s[0])      # same weights: bf16 / atomics-order noise only
    assert all(abs(a - b) <= 2e-2 * abs(b) for a, b in zip(l_f, l_s))            # the two runs track each other
    # one AdamW step moves a weight by about lr at most; the two runs see gradients that differ by bf16 / atomics noise, so
    # compare at the scale of the update, not at fp32 resolution: worst case (opposite signs every step) 2 x 3 x lr
    lr = hyper["lr"]
    tot = cnt = 0.0
    for a, b in zip(p_f + e_f, p_s + e_s):
        assert torch.isfinite(a).all()
        d = (a - b).abs()
        assert d.max().item() <= 2 * 3 * lr * 1.05
        tot += d.sum().item(); cnt += d.numel()
    print(f"[full-size] mean |fused - stock| over params + EMA = {tot / cnt:.3e} (lr {lr})")
    assert tot / cnt <= 0.25 * lr
    fresh = list(_full_net(cuda_dev, "bf16")[0].parameters())
    moved = max((a - b.detach()).abs().max().item() for a, b in zip(p_f, fresh))
    assert moved > 1e-4                                   # the steps really updated the weights


# BASELINE.json configs[2..4] at their FULL patch sizes (the oracle-sized versions are tests/test_unet_gpu.py::CONFIG_CASES)
FULL_CONFIGS = [
    ("cfg3", ["organ", "pancreatic_lesion"], (128, 128, 128), ("report", "mask")),
    ("cfg4", ["pancreas", "pancreatic_lesion", "veins"], (96, 192, 192), ("mask", "report")),
    ("cfg5", ["organ"] + sorted(f"{o}_lesion" for o in ("adrenal", "bladder", "colon", "esophagus", "kidney", "liver", "spleen")),
     (160, 160, 160), ("report", "mask")),
]


@pytest.mark.parametrize("name,classes,shape,kinds", FULL_CONFIGS)
def test_full_size_config_train_step_properties(cuda_dev, name, classes, shape, kinds):
    """The whole step (base-32 UNet bf16 -> calculate_loss with Volume + Ball terms on the mixed batch -> backward) at the
    full patch size of each BASELINE.json config.  No oracle finishes these sizes in seconds, so the checks are the
    size-independent ones: the reference's dict keys, finite non-negative terms, 'overall' = sum of the terms, every
    parameter receives a finite gradient (the DDP find_unused_parameters=False contract), and the
    mask-supervised part equals the oracle's segmentation loss evaluated on the SAME logits (torch ops on the GPU)."""
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import losses
    from rsuper_b200.unet import B200UNet
    from oracle.unet_ref import synthetic_state_dict
    C = len(classes)
    net = B200UNet(1, 32, num_classes=C, precision="bf16").to(cuda_dev)
    net.load_state_dict(synthetic_state_dict(32, C, device=cuda_dev))
    batch = synth.make_batch(list(kinds), classes, shape, seed=17, device=cuda_dev)
    args = LR.default_args()
    out = net(batch["image"])
    assert out["segmentation"].shape == (len(kinds), C) + tuple(shape)
    res = losses.calculate_loss(out, batch["label"], batch["unk_channels"], args, None, batch["mask"], batch["volumes"],
                                batch["diameters"], classes, input_tensor=batch["image"])
    vals = {k: v.item() for k, v in res.items()}
    print(f"[full-size] {name} {shape}: {({k: round(v, 6) for k, v in vals.items()})}")
    assert {"segmentation", "overall"} <= set(vals) and all(np.isfinite(v) and v >= 0 for v in vals.values())
    assert abs(vals["overall"] - sum(v for k, v in vals.items() if k != "overall")) <= 1e-5 * max(1.0, vals["overall"])
    res["overall"].backward()
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0, (name, k)
    # mask-only batch at the same size: our loss == the oracle's formula on identical logits
    lab = batch["label"]
    lg = out["segmentation"].detach()
    mine = losses.calculate_loss({"segmentation": lg.clone().requires_grad_(True)}, lab, None,
                                 LR.default_args(report_volume_loss_basic=0.0), None, None, None, None, classes)["overall"].item()
    ref = LR.seg_loss(lg, lab.float(), torch.ones_like(lg)).item()
    assert abs(mine - ref) <= 1e-5 * abs(ref), (name, mine, ref)


# ------------------------------------------------------------------------------------------------------------------
# N2 (second half): online intensity augmentations
# ------------------------------------------------------------------------------------------------------------------
def test_intensity_augmentations_match_reference_golden(cuda_dev):
    """rsuper_b200.augment with the draws the REAL training/augmentation.py made (tests/golden/reference_augment.npz):
    multiply / additive / noise / contrast are the same fp32 roundings (1e-6), gamma (powf) and the separable blur 1e-5;
    then the loader's whole gate block under the same np.random / torch seeds."""
    import os
    from oracle import augment_ref as AR
    from oracle.unet_ref import synthetic_image
    from rsuper_b200 import augment as A
    from test_oracle_golden import AUG_SEEDS, AUG_SHAPE, aug_single_draw
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_augment.npz"))
    x = synthetic_image(1, *AUG_SHAPE, seed=21)
    xd = x.to(cuda_dev)
    sub = lambda t: t.cpu().numpy()[0, 0, ::2, ::3, ::2]
    for name in AUG_SEEDS:
        d = aug_single_draw(name, AUG_SHAPE)
        y = {"multiply": lambda: A.brightness_multiply(xd, factor=d), "additive": lambda: A.brightness_additive(xd, 0.1, offset=d),
             "gamma": lambda: A.gamma(xd, gamma=d), "contrast": lambda: A.contrast(xd, factor=d),
             "blur": lambda: A.gaussian_blur(xd, sigma=d), "noise": lambda: A.gaussian_noise(xd, 0.13, noise=d.to(cuda_dev))}[name]()
        assert y.shape == xd.shape and y.device.type == cuda_dev.type
        tol = 1e-5 if name in ("gamma", "blur") else 1e-6
        np.testing.assert_allclose(sub(y), gold[name], rtol=tol, atol=tol, err_msg=name)
        assert abs(y.double().sum().item() - float(gold[f"{name}_sum"])) <= 1e-5 * max(1.0, abs(float(gold[f"{name}_sum"]))), name
    # the statistics kernel itself
    st = __import__("rsuper_b200.ops", fromlist=["ops"]).aug_stats(xd).cpu()
    want = torch.stack([x.min(), x.max(), x.mean(), x.std()])
    torch.testing.assert_close(st, want, rtol=1e-6, atol=1e-6)
    # the loader's gate block: same seeds as the reference run -> same gates, same draws
    np.random.seed(5)
    torch.manual_seed(6)
    y = A.online_intensity_augmentation(xd)
    np.testing.assert_allclose(sub(y), gold["seq_gated"], rtol=1e-5, atol=2e-6)
    np.random.seed(5)
    torch.manual_seed(6)
    draws = AR.draws_like_reference(x.shape, gates=[True] * 6)
    y = xd
    y = A.brightness_multiply(y, factor=draws["multiply"])
    y = A.brightness_additive(y, 0.1, offset=draws["additive"])
    y = A.gamma(y, gamma=draws["gamma"])
    y = A.contrast(y, factor=draws["contrast"])
    y = A.gaussian_blur(y, sigma=draws["blur"])
    y = A.gaussian_noise(y, draws["noise"][0], noise=draws["noise"][1].to(cuda_dev))
    np.testing.assert_allclose(sub(y), gold["seq_all"], rtol=2e-5, atol=5e-6)
    # device-side noise: same distribution (mean / std of the increment), different stream
    z = A.gaussian_noise(xd, 0.2, device_noise=True) - xd
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 0.2) < 5e-3


def test_public_dice_loss_multiclass_vs_oracle(cuda_dev):
    """The reference's public copy-out function DiceLossMultiClass (README.md:119-128) on the seg-loss kernels: value and
    gradient (alpha carries gradient) against the oracle restatement, with unknown voxels, class weights and the 3-D / 4-D
    input forms the reference accepts."""
    from oracle import losses_ref as LR
    from oracle import synth
    from rsuper_b200 import losses
    shape = (24, 20, 28)
    lab = synth.make_batch(["mask", "report"], ["organ", "a_lesion", "veins"], shape, seed=4, device=cuda_dev)
    logits = synth.synthetic_logits(2, 3, shape, seed=6, device=cuda_dev)
    known = 1 - lab["unk_channels"].float()
    cw = torch.tensor([[1.0, 2.0, 0.5], [1.0, 0.0, 0.5]], device=cuda_dev).view(2, 3, 1, 1, 1).expand(2, 3, *shape)
    for weights in (None, cw):
        a = logits.clone().requires_grad_(True)
        b = logits.clone().requires_grad_(True)
        ours = losses.DiceLossMultiClass(a, lab["label"], known, class_weights=weights)
        ref = LR.dice_loss_multiclass(b, lab["label"].float(), known, class_weights=weights)
        assert abs(ours.item() - ref.item()) <= 1e-5 * abs(ref.item())
        (3.0 * ours).backward(); (3.0 * ref).backward()
        assert rel(a.grad, b.grad) <= 1e-4
    one = losses.DiceLossMultiClass(logits[0, 1], lab["label"][0, 1], known[0, 1])            # [D, H, W]
    ref1 = LR.dice_loss_multiclass(logits[0:1, 1:2], lab["label"][0:1, 1:2].float(), known[0:1, 1:2])
    assert abs(one.item() - ref1.item()) <= 1e-5 * abs(ref1.item())
    four = losses.DiceLossMultiClass(logits[1], lab["label"][1], known[1])                   # [C, D, H, W]
    ref4 = LR.dice_loss_multiclass(logits[1:2], lab["label"][1:2].float(), known[1:2])
    assert abs(four.item() - ref4.item()) <= 1e-5 * abs(ref4.item())
    with pytest.raises(NotImplementedError):
        losses.DiceLossMultiClass(logits, lab["label"], known, reduce=False)


def test_widened_edge_cases(cuda_dev):
    """Empty / degenerate / ragged inputs of the widened rows: parameters without gradients and zero-size parameters are left
    alone by the fused optimizer; batch > 1 and a 1-voxel-thick volume through the sliding-window blend; single-voxel and
    line-shaped volumes through the connected components; one class (one bit used per packed byte)."""
    from types import SimpleNamespace
    from oracle import synth
    from oracle.inference_ref import connected_components as oracle_cc
    from oracle.inference_ref import inference_sliding_window as oracle_sw
    from rsuper_b200 import ops
    from rsuper_b200.inference import connected_components, inference_sliding_window, keep_largest_component
    from rsuper_b200.optim import B200AdamW
    # optimizer: frozen parameter (grad None), empty parameter
    a, frozen, empty = (torch.nn.Parameter(torch.full((10,), 0.5, device=cuda_dev)), torch.nn.Parameter(torch.ones(7, device=cuda_dev)),
                        torch.nn.Parameter(torch.zeros(0, device=cuda_dev)))
    opt = B200AdamW([a, frozen, empty], lr=0.1, weight_decay=0.0, max_norm=None)
    a.grad = torch.ones_like(a)
    empty.grad = torch.zeros_like(empty)
    opt.step()
    assert torch.equal(frozen.detach().cpu(), torch.ones(7)) and frozen not in opt.state
    torch.testing.assert_close(a.detach().cpu(), torch.full((10,), 0.4), rtol=1e-5, atol=1e-6)     # first Adam step = -lr * sign(g)
    B200AdamW([torch.nn.Parameter(torch.ones(3, device=cuda_dev))], max_norm=1.0).step()         # nothing has a gradient: no launch, no error
    # sliding window: batch of two, window = whole (thin) volume
    net = _sliding_net(cuda_dev)
    vol = torch.stack([synth.synthetic_logits(1, 1, (16, 24, 20), seed=s)[0] for s in (1, 2)])
    got = inference_sliding_window(net, vol.to(cuda_dev), SimpleNamespace(window_size=[16, 16, 16], classes=3))
    want = oracle_sw(_sliding_net("cpu"), vol, (16, 16, 16), 3)
    assert got.shape == (2, 3, 16, 24, 20) and (got - want).abs().max().item() <= 2e-6
    # connected components: single voxel, a line, a 1x1x1 volume
    for m in (np.ones((1, 1, 1), np.uint8), np.zeros((1, 1, 1), np.uint8), np.array([1, 1, 0, 1, 0, 0, 1, 1, 1], np.uint8).reshape(1, 1, 9),
              np.array([1, 0, 1, 1], np.uint8).reshape(4, 1, 1)):
        labels, n = connected_components(torch.from_numpy(m).to(cuda_dev))
        assert n == oracle_cc(m)[1]
        big = keep_largest_component(torch.from_numpy(m).to(cuda_dev)).cpu().numpy()
        assert big.sum() == (np.bincount(oracle_cc(m)[0].reshape(-1))[1:].max() if n else 0)
    # the raster-first component wins a tie in size (strict '>' in the reference loop)
    tie = np.array([1, 1, 0, 1, 1], np.uint8).reshape(1, 1, 5)
    assert keep_largest_component(torch.from_numpy(tie).to(cuda_dev)).cpu().numpy().reshape(-1).tolist() == [1, 1, 0, 0, 0]
    # one class: bit 7 of every packed byte
    one = (torch.rand(1, 1, 3, 5, 7) < 0.5).to(torch.uint8)
    packed = torch.from_numpy(np.stack([synth.pack_masks(one[0])])).to(cuda_dev)
    assert torch.equal(ops.unpack_masks(packed, 1).cpu(), one)


def test_capturable_optimizer_equals_eager_optimizer(cuda_dev):
    """B200AdamW(capturable=True) — step-dependent scalars read from device memory, refreshed by prepare_step() — produces
    bit-identical parameters / state / EMA to the by-value mode over steps with a changing learning rate."""
    from rsuper_b200.optim import B200AdamW
    sets = []
    for capturable in (False, True):
        ps = [torch.nn.Parameter(t) for t in _param_set(cuda_dev, 0)]
        ema = [p.detach().clone() for p in ps]
        opt = B200AdamW(ps, lr=6e-4, weight_decay=0.05, max_norm=1.0, ema_params=ema, capturable=capturable)
        for step in range(4):
            opt.param_groups[0]["lr"] = 6e-4 * (0.5 + 0.25 * step)            # an LR schedule writes group['lr']
            for p, g in zip(ps, _param_set(cuda_dev, 20 + step)):
                p.grad = g.clone() if p.data_ptr() % 16 == 0 else torch.cat([g.new_zeros(1), g.reshape(-1)])[1:].view(g.shape)
            if capturable:
                opt.prepare_step()
            opt.step()
        sets.append(([p.detach().clone() for p in ps], ema, [opt.state[p]["exp_avg_sq"] for p in ps], opt.global_step))
        if capturable:
            with pytest.raises(RuntimeError, match="prepare_step"):
                opt.step()
    (p0, e0, v0, s0), (p1, e1, v1, s1) = sets
    assert s0 == s1 == 4
    for a, b in zip(p0 + e0 + v0, p1 + e1 + v1):
        assert torch.equal(a, b)


def test_graphed_train_step_matches_eager(cuda_dev):
    """The whole step captured in a CUDA graph (GraphedTrainStep) against the same steps issued eagerly from identical
    initial state: same losses (bf16 / atomics noise), parameters within optimizer rounding."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict
    from rsuper_b200 import losses
    from rsuper_b200.graph_step import GraphedTrainStep
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.unet import B200UNet
    classes = ["organ", "pancreatic_lesion"]
    x = synthetic_image(2, 64, 64, 64, seed=5, device=cuda_dev)
    lab = synth.make_batch(["mask", "mask"], classes, (64, 64, 64), seed=8, device=cuda_dev)["label"]
    args = LR.default_args(report_volume_loss_basic=0.0)
    args.nan_check = False                                                    # host sync: moved to GraphedTrainStep.check
    loss_fn = lambda out, lb: losses.calculate_loss(out, lb, None, args, None, None, None, None, classes)["overall"]
    runs = []
    for graphed in (False, True):
        net = B200UNet(1, 16, num_classes=2, precision="bf16").to(cuda_dev)
        net.load_state_dict(synthetic_state_dict(16, 2, device=cuda_dev))
        params = list(net.parameters())
        ema = [p.detach().clone() for p in params]
        opt = B200AdamW(params, lr=6e-4, weight_decay=0.05, max_norm=1.0, ema_params=ema, capturable=True)
        ls = []
        if graphed:
            step = GraphedTrainStep(net, loss_fn, opt, x, lab, warmup=1)      # one eager step inside (it is a real update)
            assert step.warmup_steps == 1 and opt.global_step == 1 and step.launches_per_step > 100
            ls.append(step.loss.item())
            for _ in range(3):
                ls.append(GraphedTrainStep.check(step(x, lab).item()))
        else:
            for _ in range(4):
                opt.zero_grad(set_to_none=True)
                opt.prepare_step()
                loss = loss_fn(net(x), lab)
                loss.backward()
                opt.step()
                ls.append(loss.item())
        runs.append((ls, [p.detach().clone() for p in params], ema, opt.global_step))
    (l0, p0, e0, s0), (l1, p1, e1, s1) = runs
    print(f"[graph] eager losses {l0}, graphed losses {l1}")
    assert s0 == s1 == 4
    assert all(abs(a - b) <= 2e-2 * abs(a) for a, b in zip(l0, l1)) and abs(l0[0] - l1[0]) <= 2e-3 * abs(l0[0])
    tot = cnt = 0.0
    for a, b in zip(p0 + e0, p1 + e1):
        d = (a - b).abs()
        assert torch.isfinite(b).all() and d.max().item() <= 2 * 4 * 6e-4 * 1.05
        tot += d.sum().item(); cnt += d.numel()
    assert tot / cnt <= 0.25 * 6e-4


def test_train_step_prefetch_feeds_the_same_batches(cuda_dev):
    """B200TrainStep.prefetch (next batch uploaded from pinned host memory on a copy stream while the current step runs) against
    passing the same host batches to step(...) directly: two alternating batches, identical initial state — the same losses step
    by step (the right batch reaches the right step) and the same parameters."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict
    from rsuper_b200 import losses
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.train_step import B200TrainStep
    from rsuper_b200.unet import B200UNet
    classes = ["organ", "pancreatic_lesion"]
    batches = []
    for seed in (5, 6):
        x = synthetic_image(2, 64, 64, 64, seed=seed)
        lab = synth.make_batch(["mask", "mask"], classes, (64, 64, 64), seed=seed + 3, device="cpu")["label"].contiguous()
        if seed == 6:
            lab = torch.ones_like(lab)                           # an all-foreground batch: its loss is far from the other batch's
        batches.append((x.pin_memory(), lab.pin_memory()))
    args = LR.default_args(report_volume_loss_basic=0.0)
    args.nan_check = False
    loss_fn = lambda out, lb: losses.calculate_loss(out, lb, None, args, None, None, None, None, classes)["overall"]
    order = [0, 1, 1, 0, 1]
    runs = []
    for prefetch in (False, True):
        net = B200UNet(1, 16, num_classes=2, precision="bf16").to(cuda_dev)
        net.load_state_dict(synthetic_state_dict(16, 2, device=cuda_dev))
        params = list(net.parameters())
        opt = B200AdamW(params, lr=6e-4, weight_decay=0.05, max_norm=1.0, capturable=True)
        step = B200TrainStep(net, loss_fn, opt, [t.to(cuda_dev) for t in batches[0]], schedule="graph", warmup=1)
        ls = []
        if prefetch:
            with pytest.raises(RuntimeError, match="prefetch"):
                step()
            step.prefetch(*batches[order[0]])
            pending = None
            for i in range(len(order)):
                step()
                h = step.loss_async()                              # 4-byte pinned copy + event: read one step late
                if i + 1 < len(order):
                    step.prefetch(*batches[order[i + 1]])          # overlaps with the step that was just enqueued
                if pending is not None:
                    ls.append(pending.get())
                pending = h
            ls.append(pending.get())
        else:
            for b in order:
                ls.append(step(*batches[b]).item())
        runs.append((ls, [p.detach().clone() for p in params]))
    (l0, p0), (l1, p1) = runs
    print(f"[prefetch] direct losses {l0}, prefetched losses {l1}")
    # graph replays are not bit-reproducible (atomics in the statistics / weight-gradient reductions): the bound of
    # test_graphed_train_step_matches_eager; a batch reaching the wrong step would move the loss by far more
    assert all(abs(a - b) <= 2e-2 * abs(a) for a, b in zip(l0, l1)) and abs(l0[0] - l1[0]) <= 2e-3 * abs(l0[0])
    # order 0,1,1,0,1: every switch between the two batches moves the loss by far more than the two runs differ
    assert min(abs(l0[i] - l0[i + 1]) for i in (0, 2, 3)) > 3 * max(abs(a - b) for a, b in zip(l0, l1))
    for a, b in zip(p0, p1):
        assert (a - b).abs().max().item() <= 2 * len(order) * 6e-4 * 1.05


def test_split_schedule_matches_eager_on_report_batches(cuda_dev):
    """B200TrainStep(schedule='split') — UNet forward and backward + optimizer as two CUDA graphs, calculate_loss with the
    Volume / Ball report losses (host-controlled tumour loop) launched eagerly in between — against schedule='eager' from
    identical initial state on a mixed mask / report batch: same loss dictionaries' total at every step, parameters and EMA
    within optimizer rounding."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_state_dict
    from rsuper_b200 import losses
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.train_step import B200TrainStep
    from rsuper_b200.unet import B200UNet
    classes = ["organ", "pancreatic_lesion"]
    shape = (48, 64, 48)
    bt = synth.make_batch(["report", "mask"], classes, shape, seed=17, device=cuda_dev)
    args = LR.default_args()
    args.nan_check = False
    assert not losses.capturable(args, True) and losses.capturable(args, False)

    def loss_fn(out, lab, unk, msk, vol, dia):
        return losses.calculate_loss(out, lab, unk, args, None, msk, vol, dia, classes)["overall"]

    keys = ["image", "label", "unk_channels", "mask", "volumes", "diameters"]
    runs = []
    for schedule in ("eager", "split"):
        net = B200UNet(1, 16, num_classes=2, precision="bf16").to(cuda_dev)
        net.load_state_dict(synthetic_state_dict(16, 2, device=cuda_dev))
        params = list(net.parameters())
        ema = [p.detach().clone() for p in params]
        opt = B200AdamW(params, lr=6e-4, weight_decay=0.05, max_norm=1.0, ema_params=ema, capturable=True)
        step = B200TrainStep(net, loss_fn, opt, [bt[k] for k in keys], schedule=schedule, warmup=1)
        # the split constructor already took its one warm-up step on this batch (and left the loss of the NEXT forward,
        # evaluated while seeding the second capture, in step.loss): its three calls are steps 2-4 of the eager run
        ls = []
        for _ in range(3 if schedule == "split" else 4):
            ls.append(B200TrainStep.check(step(*[bt[k] for k in keys]).item()))
        assert step.launches_per_step > 150
        runs.append((ls, [p.detach().clone() for p in params], ema, opt.global_step))
    (l0, p0, e0, s0), (l1, p1, e1, s1) = runs
    print(f"[split] eager losses {l0}, split-graph losses {l1}")
    assert s0 == s1 == 4
    assert len(l0) == 4 and len(l1) == 3
    assert all(abs(a - b) <= 2e-2 * abs(a) for a, b in zip(l0[1:], l1)) and abs(l0[1] - l1[0]) <= 2e-3 * abs(l0[1])
    tot = cnt = 0.0
    for a, b in zip(p0 + e0, p1 + e1):
        d = (a - b).abs()
        assert torch.isfinite(b).all() and d.max().item() <= 2 * 4 * 6e-4 * 1.05
        tot += d.sum().item(); cnt += d.numel()
    assert tot / cnt <= 0.25 * 6e-4
