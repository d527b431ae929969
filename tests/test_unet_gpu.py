"""GPU parity of the full B200UNet forward / backward (through the reference-facing module API)
against the oracle, and against the REAL reference's recorded logits (tests/golden)."""
import copy
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BASE, C, S = 8, 2, 32


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def _make(cuda_dev, precision, slope=0.0, base=BASE, classes=C):
    from oracle.unet_ref import synthetic_state_dict
    from rsuper_b200.unet import B200UNet
    net = B200UNet(1, base, num_classes=classes, scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5,
                   block="BasicBlock", precision=precision, negative_slope=slope).to(cuda_dev)
    sd = synthetic_state_dict(base, classes, device=cuda_dev)
    assert list(sd.keys()) == [k for k, _ in net.named_parameters()]  # SURVEY.md §8b state-dict contract
    net.load_state_dict(sd, strict=True)
    return net, sd


def _truth(x, sd, slope=0.0):
    """fp64 evaluation of the oracle: the yardstick both bf16-operand computations are measured against."""
    from oracle.unet_ref import unet_forward
    sd64 = {k: v.double() for k, v in sd.items()}
    return unet_forward(x.double(), sd64, slope=slope)


def test_logits_vs_reference_golden(cuda_dev, golden):
    """Against the REAL reference's recorded fp32 logits (cfg1: base 8, 32^3).  North-star bar: logits within 1e-3
    relative fp32 and the argmax mask bit-exact — met by precision='fp32' (split 3xbf16 tensor-core products).
    The bf16 perf mode is bounded by the bf16-operand noise of this net (2^3 voxels at the bottom level, where one
    operand rounding is a visible perturbation); the fp64-yardstick tests below show it equals the oracle's own
    noise when the oracle rounds at the same points."""
    from oracle.unet_ref import synthetic_image, unet_forward
    x = synthetic_image(1, S, S, S, seed=3, device=cuda_dev)
    ref = torch.from_numpy(golden["unet_logits"]).to(cuda_dev)
    for precision in ("fp32", "bf16"):
        net, sd = _make(cuda_dev, precision)
        with torch.no_grad():
            out = net(x)["segmentation"]
        e = rel(out, ref)
        agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
        print(f"[golden] {precision}: rel_to_max={e:.3e} argmax agreement={agree:.5f}")
        if precision == "fp32":
            assert e <= 1e-3 and agree == 1.0
        else:
            # the perf mode is held to the error the ORACLE itself makes against the real reference's logits when it rounds
            # at the same points (bf16 conv operands, bf16 activation storage): at most twice that, argmax no worse
            with torch.no_grad():
                emul = unet_forward(x, sd, emulate=True, storage="bf16")
            e_emul = rel(emul, ref)
            agree_emul = (emul.argmax(1) == ref.argmax(1)).float().mean().item()
            print(f"[golden] bf16-emulating oracle: rel_to_max={e_emul:.3e} argmax agreement={agree_emul:.5f}")
            assert e <= 2.0 * e_emul + 1e-3 and agree >= agree_emul - 5e-3


@pytest.mark.parametrize("precision,slope,side", [("fp32", 0.0, 64), ("bf16", 0.0, 64), ("fp32", 0.01, 32), ("bf16", 0.0, 32)])
def test_forward_accuracy_vs_fp64_truth(cuda_dev, precision, slope, side):
    """fp32 mode: within the north-star 1e-3 of the fp64 truth, argmax bit-exact.  bf16 mode: error against the
    fp64 truth must not exceed that of the oracle evaluated with the same rounding points (bf16 operands + bf16
    storage): the kernels add no error of their own."""
    from oracle.unet_ref import synthetic_image, unet_forward
    net, sd = _make(cuda_dev, precision, slope)
    x = synthetic_image(2 if side == 32 else 1, side, side, side, seed=4, device=cuda_dev)
    with torch.no_grad():
        out = net(x)["segmentation"]
        truth = _truth(x, sd, slope)
        emul = unet_forward(x, sd, slope=slope, emulate=True, storage="bf16")
        pure = unet_forward(x, sd, slope=slope)
    e_mine, e_emul, e_pure = rel(out.double(), truth), rel(emul.double(), truth), rel(pure.double(), truth)
    agree = (out.argmax(1) == truth.argmax(1)).float().mean().item()
    agree_emul = (emul.argmax(1) == truth.argmax(1)).float().mean().item()
    print(f"[fwd] {precision} slope={slope} {side}^3: err vs fp64 truth: kernels {e_mine:.3e} | bf16-emulating oracle {e_emul:.3e} | "
          f"fp32 oracle {e_pure:.3e}; argmax agreement kernels {agree:.5f} emul {agree_emul:.5f}")
    if precision == "fp32":
        # reference configuration (ReLU): the north-star 1e-3.  The LeakyReLU variant on the 32^3 toy (InstanceNorm over
        # 2^3 voxels at the bottom, nothing clamped to zero) amplifies the 2^-17 split-product error by up to ~400x
        # depending on the run's atomic-add order in the statistics (observed 5e-4 .. 3.4e-3); it gets a 1e-2 bound.
        assert e_mine <= (1e-3 if slope == 0.0 else 1e-2) and agree == 1.0
    else:
        assert e_mine <= 2.0 * e_emul + 1e-3
        assert agree >= agree_emul - 5e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_backward_accuracy_vs_fp64_truth(cuda_dev, precision):
    S = 64 if precision == "fp32" else 32   # fp32 bars are checked on a patch whose bottom level is 4^3, not 2^3, voxels
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, unet_forward
    from rsuper_b200 import losses
    net, sd = _make(cuda_dev, precision)
    classes = ["organ", "pancreatic_lesion"]
    x = synthetic_image(2, S, S, S, seed=5, device=cuda_dev)
    batch = synth.make_batch(["mask", "mask"], classes, (S, S, S), seed=7, device=cuda_dev)
    args = LR.default_args(report_volume_loss_basic=0.0)
    out = net(x)
    loss = losses.calculate_loss(out, batch["label"], None, args, None, None, None, None, classes)
    assert sorted(loss.keys()) == ["overall", "report", "segmentation"]
    loss["overall"].backward()

    def oracle_grads(dtype, emulate):
        sdr = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
        lg = unet_forward(x.to(dtype), sdr, emulate=emulate, storage=precision)
        l = LR.calculate_loss({"segmentation": lg}, batch["label"].long(), None, args, None, None, None, None, classes)
        l["overall"].backward()
        return l["overall"].item(), {k: v.grad for k, v in sdr.items()}

    l64, g64 = oracle_grads(torch.float64, False)
    l32, g32 = oracle_grads(torch.float32, False)
    lem, gem = oracle_grads(torch.float32, True)
    print(f"[bwd] {precision}: loss kernels {loss['overall'].item():.6f} | fp32 oracle {l32:.6f} | bf16-emulating oracle {lem:.6f} | fp64 truth {l64:.6f}")
    if precision == "fp32":
        assert abs(loss["overall"].item() - l64) <= 1e-5 * abs(l64)   # north star: loss within 1e-5
    else:
        assert abs(loss["overall"].item() - l64) <= 2.0 * abs(lem - l64) + 2e-3 * abs(l64)
    worst = (0.0, 0.0, 0.0, "")
    for k, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, k   # DDP: every parameter gets a grad
        e_mine, e_emul, e_32 = rel(p.grad.double(), g64[k]), rel(gem[k].double(), g64[k]), rel(g32[k].double(), g64[k])
        if e_mine > worst[0]:
            worst = (e_mine, e_emul, e_32, k)
        if precision == "fp32":
            # split products carry ~2^-17 per operand; this 2^3-bottom toy net amplifies any perturbation of the
            # forward activations in its gradients (the fp32 oracle's own error is printed for scale)
            assert e_mine <= 5e-2, (k, e_mine, e_32)
        else:
            # the emulating oracle rounds forward operands only (autograd backward stays fp32) while the kernels
            # also round dy for dgrad & wgrad: allow that extra 2^-8-level term
            assert e_mine <= 3.0 * e_emul + 2e-2, (k, e_mine, e_emul)
    print(f"[bwd] {precision}: worst grad err vs fp64 truth {worst[0]:.3e} (fp32 oracle {worst[2]:.3e}, bf16-emulating oracle {worst[1]:.3e}) at {worst[3]}")


def test_module_contract(cuda_dev):
    """What train_ddp.py does to a model: deepcopy for EMA, pickle for checkpoints,
    train()/eval(), parameters()/buffers() iteration, strict=False loading, dict output."""
    net, sd = _make(cuda_dev, "bf16")
    assert len(list(net.buffers())) == 0 and len(list(net.parameters())) == 45
    ema = copy.deepcopy(net)
    blob = pickle.dumps(net.cpu())
    net2 = pickle.loads(blob).to(cuda_dev)
    net2.load_state_dict(sd, strict=False)
    net2.eval()
    x = torch.zeros(1, 1, 32, 32, 32, device=cuda_dev)
    with torch.no_grad():
        o = net2(x)
    assert isinstance(o, dict) and o["segmentation"].shape == (1, C, 32, 32, 32)
    for pe, p in zip(ema.parameters(), net2.parameters()):
        pe.data.mul_(0.99).add_(p.data.to(pe.device), alpha=0.01)  # update_ema_variables, training/utils.py:154-161


def test_input_validation(cuda_dev):
    from rsuper_b200.unet import B200UNet
    net, _ = _make(cuda_dev, "bf16")
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 24, 32, 32, device=cuda_dev))   # not divisible by 16
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 32, 32, 32))                      # CPU tensor: no fallback path
    with pytest.raises(ValueError):
        B200UNet(1, 8, num_classes=2, block="Bottleneck")       # inner convolutions run at base / 2 channels: base % 16
    with pytest.raises(NotImplementedError):
        B200UNet(1, 8, num_classes=2, norm="bn")
    with pytest.raises(ValueError):
        B200UNet(1, 8, num_classes=2, up_mode="nearest")


# BASELINE.json configs[2..4], scaled down to sizes the CPU oracle finishes in seconds: the same class lists, mixed
# mask / report batches, a non-cubic PanTS-shaped patch and the 7-tumor 8-class head.
CONFIG_CASES = [
    ("cfg3", ["organ", "pancreatic_lesion"], (48, 64, 48), ("report", "mask")),   # bottom level 3x4x3 (a 2^3 bottom amplifies split-product noise past 1e-3)
    ("cfg4", ["pancreas", "pancreatic_lesion", "veins"], (32, 64, 48), ("mask", "report")),
    ("cfg5", ["organ"] + sorted(f"{o}_lesion" for o in ("adrenal", "bladder", "colon", "esophagus", "kidney", "liver", "spleen")),
     (32, 32, 48), ("report", "mask")),
]


@pytest.mark.parametrize("name,classes,shape,kinds", CONFIG_CASES)
def test_train_step_configs_vs_oracle(cuda_dev, name, classes, shape, kinds):
    """UNet forward -> calculate_loss (BCE + Dice + Volume + Ball on report samples) -> backward through the
    reference-facing API in the parity mode, against the fp32 oracle on the same seeded batch.  The pseudo-mask
    construction of the Ball loss is discrete (argmax / top-k on the logits), so the loss terms are compared at the
    oracle's own logits too: that isolates the loss kernels (1e-5) from the network's 1e-4-level logit noise."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import unet_forward
    from rsuper_b200 import losses
    C = len(classes)
    net, sd = _make(cuda_dev, "fp32", classes=C)
    batch = synth.make_batch(list(kinds), classes, shape, seed=17, device=cuda_dev)
    args = LR.default_args()
    x = batch["image"]
    out = net(x)
    with torch.no_grad():
        ref_logits = unet_forward(x, sd)
    assert out["segmentation"].shape == (len(kinds), C) + tuple(shape)
    # (the strict 1e-3 / bit-exact-argmax bars are asserted on 64^3 patches in the accuracy tests above; these small
    # patches have a 2x..4x smaller bottom level, where InstanceNorm amplifies the 2^-17 split-product noise)
    e_logits = rel(out["segmentation"], ref_logits)
    agree = (out["segmentation"].argmax(1) == ref_logits.argmax(1)).float().mean().item()
    print(f"[cfg] {name}: logits rel err {e_logits:.3e}, argmax agreement {agree:.6f}")
    assert e_logits <= 2e-3 and agree >= 0.9999

    def call(mod, logits, on_cpu):
        b = {k: (v.cpu() if on_cpu else v) for k, v in batch.items()}
        lab = b["label"].long() if on_cpu else b["label"]
        return mod.calculate_loss({"segmentation": logits}, lab, b["unk_channels"].float(), args, None, b["mask"].float(),
                                  b["volumes"], b["diameters"], classes, input_tensor=b["image"])

    # (1) loss kernels at identical logits: every term within 1e-5 of the oracle (CPU fp32, like tests/test_report_losses_gpu.py)
    lg = ref_logits.clone().requires_grad_(True)
    lr = ref_logits.cpu().clone().requires_grad_(True)
    mine, ref = call(losses, lg, False), call(LR, lr, True)
    assert sorted(mine.keys()) == sorted(ref.keys())
    for k in ref:
        assert abs(mine[k].item() - ref[k].item()) <= 1e-5 * max(1.0, abs(ref[k].item())), (name, k, mine[k].item(), ref[k].item())
    mine["overall"].backward()
    ref["overall"].backward()
    e_grad = rel(lg.grad.cpu(), lr.grad)
    print(f"[cfg] {name}: losses {({k: round(v.item(), 6) for k, v in mine.items()})} dlogits rel err {e_grad:.3e}")
    assert e_grad <= 1e-4, (name, e_grad)
    # (2) the whole step through the module: finite, every parameter receives a gradient, loss close to the oracle's
    full = call(losses, out["segmentation"], False)
    full["overall"].backward()
    assert abs(full["segmentation"].item() - ref["segmentation"].item()) <= 1e-4 * abs(ref["segmentation"].item())
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), (name, k)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_singleconv_unet_vs_reference_and_oracle(cuda_dev, golden, precision):
    """block='SingleConv' (post-activation conv -> IN -> ReLU, conv_layers.py:56-68): same state-dict
    names as the reference, logits against the REAL reference's recorded output (32^3) and against the fp64 oracle
    (64^3: north-star 1e-3 / bit-exact argmax in the parity mode), loss and every weight gradient against autograd."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
    from rsuper_b200 import losses
    from rsuper_b200.unet import B200UNet
    net = B200UNet(1, BASE, num_classes=C, scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5, block="SingleConv",
                   precision=precision).to(cuda_dev)
    sd = synthetic_state_dict(BASE, C, device=cuda_dev, block="SingleConv")
    assert list(sd.keys()) == [k for k, _ in net.named_parameters()] and len(sd) == 20
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out32 = net(synthetic_image(1, 32, 32, 32, seed=3, device=cuda_dev))["segmentation"]
    ref32 = torch.from_numpy(golden["unet_single_logits"]).to(cuda_dev)
    e32 = rel(out32, ref32)
    agree32 = (out32.argmax(1) == ref32.argmax(1)).float().mean().item()
    import os
    S = int(os.environ.get("RSB_TEST_SINGLE_S", "64"))
    x = synthetic_image(2, S, S, S, seed=6, device=cuda_dev)
    classes = ["organ", "pancreatic_lesion"]
    batch = synth.make_batch(["mask", "mask"], classes, (S, S, S), seed=8, device=cuda_dev)
    args = LR.default_args(report_volume_loss_basic=0.0)
    out = net(x)
    loss = losses.calculate_loss(out, batch["label"], None, args, None, None, None, None, classes)
    loss["overall"].backward()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    truth = unet_forward(x.double(), sd64)
    l64 = LR.calculate_loss({"segmentation": truth}, batch["label"].long(), None, args, None, None, None, None, classes)
    l64["overall"].backward()
    with torch.no_grad():
        emul = unet_forward(x, sd, emulate=True, storage="bf16")
    e_mine, e_emul = rel(out["segmentation"].double(), truth), rel(emul.double(), truth)
    agree = (out["segmentation"].argmax(1) == truth.argmax(1)).float().mean().item()
    errs = {k: rel(p.grad.double(), sd64[k].grad) for k, p in net.named_parameters()}
    worst = max(errs.values())
    print("[single] grad errs " + " ".join(f"{k.replace('.conv.conv.weight', '')}={v:.1e}" for k, v in errs.items()))
    print(f"[single] {precision}: golden 32^3 rel {e32:.3e} argmax {agree32:.5f} | 64^3 vs fp64 truth {e_mine:.3e} (bf16-emulating oracle "
          f"{e_emul:.3e}) argmax {agree:.5f} | loss {loss['overall'].item():.6f} vs {l64['overall'].item():.6f} | worst grad err {worst:.3e}")
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    if precision == "fp32":
        assert e32 <= 3e-3 and agree32 >= 0.9999          # 2^3-voxel bottom level: see test_logits_vs_reference_golden
        assert e_mine <= 1e-3 and agree == 1.0
        assert abs(loss["overall"].item() - l64["overall"].item()) <= 1e-5 * abs(l64["overall"].item())
        # Without residual paths a ReLU whose pre-activation sits within the 2^-17 split-product noise of zero flips against
        # the fp64 truth, and at the 4^3-voxel bottom level ONE flip moves a gradient by ~1/64 of its scale: the error is
        # quantised and varies run to run with the atomic-add order of the statistics (observed worst 1.4e-3 .. 8e-2 on the
        # same inputs).  The typical parameter is held tight, the worst one loosely.
        assert float(np.median(list(errs.values()))) <= 5e-3 and worst <= 0.25
    else:
        assert e_mine <= 2.0 * e_emul + 1e-3
        assert abs(loss["overall"].item() - l64["overall"].item()) <= 2e-2 * abs(l64["overall"].item())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_bottleneck_unet_vs_reference_and_oracle(cuda_dev, precision):
    """block='Bottleneck' (1x1x1 -> 3x3x3 -> 1x1x1 pre-activation residual block, conv_layers.py:97-123): same
    state-dict names as the reference, logits / loss / gradient norms against values recorded from the REAL reference module
    (tests/golden/make_golden_bneck.py, base 16, 32^3), and logits + every weight gradient against the fp64 oracle on a patch
    whose bottom level is not 2^3 voxels.  The 1x1x1 convolutions run through the tensor-core kernel's centre-tap mode."""
    import os
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
    from rsuper_b200 import losses
    from rsuper_b200.unet import B200UNet
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_bottleneck.npz"))
    base, classes = 16, ["organ", "pancreatic_lesion"]
    net = B200UNet(1, base, num_classes=2, scale=[[2, 2, 2]] * 4, norm="in", kernel_size=[[3, 3, 3]] * 5, block="Bottleneck",
                   precision=precision).to(cuda_dev)
    sd = synthetic_state_dict(base, 2, device=cuda_dev, block="Bottleneck")
    assert list(sd.keys()) == [k for k, _ in net.named_parameters()] == list(gold["names"]) and len(sd) == 62
    net.load_state_dict(sd, strict=True)
    args = LR.default_args(report_volume_loss_basic=0.0)
    # the real reference's recorded run (1 x 32^3)
    x32 = synthetic_image(1, 32, 32, 32, seed=3, device=cuda_dev)
    b32 = synth.make_batch(["mask"], classes, (32, 32, 32), seed=5, device=cuda_dev)
    out32 = net(x32)
    loss32 = losses.calculate_loss(out32, b32["label"], None, args, None, None, None, None, classes)["overall"]
    loss32.backward()
    ref32 = torch.from_numpy(gold["logits"]).to(cuda_dev)
    e32 = rel(out32["segmentation"][:, :, ::2, ::2, ::2], ref32)
    gn = np.array([p.grad.double().norm().item() for p in net.parameters()])
    gn_err = np.abs(gn - gold["grad_norms"]) / (gold["grad_norms"] + 1e-12)
    print(f"[bottleneck] {precision}: 32^3 logits vs real reference rel {e32:.3e}; loss {loss32.item():.6f} vs {float(gold['loss']):.6f}; "
          f"gradient-norm rel err median {np.median(gn_err):.2e} max {gn_err.max():.2e}")
    if precision == "fp32":
        assert e32 <= 3e-3 and abs(loss32.item() - float(gold["loss"])) <= 1e-4 * float(gold["loss"])
        assert np.median(gn_err) <= 2e-3 and gn_err.max() <= 5e-2
    # fp64 oracle on a better conditioned patch
    S = int(os.environ.get("RSB_TEST_SINGLE_S", "64"))
    if S < 64:
        return          # CPU emulation run (tests/test_emulated_kernels.py): the recorded run of the real reference is the check
    for p in net.parameters():
        p.grad = None
    x = synthetic_image(1, S, S, S, seed=6, device=cuda_dev)
    batch = synth.make_batch(["mask"], classes, (S, S, S), seed=8, device=cuda_dev)
    out = net(x)
    loss = losses.calculate_loss(out, batch["label"], None, args, None, None, None, None, classes)["overall"]
    loss.backward()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    truth = unet_forward(x.double(), sd64)
    l64 = LR.calculate_loss({"segmentation": truth}, batch["label"].long(), None, args, None, None, None, None, classes)["overall"]
    l64.backward()
    with torch.no_grad():
        emul = unet_forward(x, sd, emulate=True, storage="bf16")
    e_mine, e_emul = rel(out["segmentation"].double(), truth), rel(emul.double(), truth)
    agree = (out["segmentation"].argmax(1) == truth.argmax(1)).float().mean().item()
    errs = {k: rel(p.grad.double(), sd64[k].grad) for k, p in net.named_parameters()}
    worst = max(errs, key=errs.get)
    print(f"[bottleneck] {precision}: {S}^3 vs fp64 oracle {e_mine:.3e} (bf16-emulating oracle {e_emul:.3e}) argmax {agree:.5f} | loss "
          f"{loss.item():.6f} vs {l64.item():.6f} | worst grad err {errs[worst]:.3e} at {worst}, median {np.median(list(errs.values())):.2e}")
    for k, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all(), k
    if precision == "fp32":
        assert e_mine <= 1e-3 and agree == 1.0 and abs(loss.item() - l64.item()) <= 1e-5 * abs(l64.item())
        # the worst single tensor moves between 5e-2 and 8e-2 from run to run (reduction order of the statistics atomics seen
        # through one ill-conditioned 1x1x1 layer); the median is the stable figure
        assert errs[worst] <= 1.5e-1 and np.median(list(errs.values())) <= 2e-3
    else:
        assert e_mine <= 2.0 * e_emul + 1e-3
        assert abs(loss.item() - l64.item()) <= 2e-2 * abs(l64.item())


@pytest.mark.parametrize("block,precision", [("BasicBlock", "fp32"), ("BasicBlock", "bf16"), ("Bottleneck", "fp32")])
def test_transposed_conv_upsampling_variant(cuda_dev, block, precision):
    """up_mode='transposed': ConvTranspose3d(C, C, kernel_size=2, stride=2) + bias in place of the trilinear interpolation of
    up_block (the north star's transposed-conv variant; vnet.py:108 semantics) — run as a 1x1x1 tensor-core conv to 8 C
    channels + a depth-to-space pass.  Checked against the torch composition of the same primitives (the oracle with
    F.conv_transpose3d) in fp64: logits, loss and every gradient incl. the transposed-conv weights and biases."""
    import os
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
    from rsuper_b200 import losses
    from rsuper_b200.unet import B200UNet
    base, classes = 16, ["organ", "pancreatic_lesion"]
    S = int(os.environ.get("RSB_TEST_SINGLE_S", "64"))
    net = B200UNet(1, base, num_classes=2, block=block, precision=precision, up_mode="transposed").to(cuda_dev)
    sd = synthetic_state_dict(base, 2, device=cuda_dev, block=block, up_mode="transposed")
    assert list(sd.keys()) == [k for k, _ in net.named_parameters()]
    assert tuple(sd["up1.up.weight"].shape) == (10 * base, 10 * base, 2, 2, 2) and tuple(sd["up4.up.bias"].shape) == (2 * base,)
    net.load_state_dict(sd, strict=True)
    args = LR.default_args(report_volume_loss_basic=0.0)
    x = synthetic_image(1, S, S, S, seed=6, device=cuda_dev)
    batch = synth.make_batch(["mask"], classes, (S, S, S), seed=8, device=cuda_dev)
    out = net(x)
    loss = losses.calculate_loss(out, batch["label"], None, args, None, None, None, None, classes)["overall"]
    loss.backward()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    truth = unet_forward(x.double(), sd64)
    l64 = LR.calculate_loss({"segmentation": truth}, batch["label"].long(), None, args, None, None, None, None, classes)["overall"]
    l64.backward()
    with torch.no_grad():
        emul = unet_forward(x, sd, emulate=True, storage="bf16")
    e_mine, e_emul = rel(out["segmentation"].double(), truth), rel(emul.double(), truth)
    agree = (out["segmentation"].argmax(1) == truth.argmax(1)).float().mean().item()
    errs = {k: rel(p.grad.double(), sd64[k].grad) for k, p in net.named_parameters()}
    # the transposed-conv bias feeds an InstanceNorm: its true gradient is zero (fp64: ~1e-15), so it is compared on the scale
    # of the matching weight gradient instead of relative to itself
    for j in range(1, 5):
        wmax = sd64[f"up{j}.up.weight"].grad.abs().max().item()
        bp = dict(net.named_parameters())[f"up{j}.up.bias"]
        errs[f"up{j}.up.bias"] = (bp.grad.double() - sd64[f"up{j}.up.bias"].grad).abs().max().item() / wmax
    worst = max(errs, key=errs.get)
    up_errs = {k: v for k, v in errs.items() if ".up." in k}
    print(f"[transposed] {block} {precision} {S}^3: logits vs fp64 oracle {e_mine:.3e} (bf16-emulating oracle {e_emul:.3e}) argmax {agree:.5f} | "
          f"loss {loss.item():.6f} vs {l64.item():.6f} | worst grad err {errs[worst]:.3e} at {worst}; transposed-conv params "
          + " ".join(f"{k}={v:.1e}" for k, v in up_errs.items()))
    for k, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all(), k
    if precision == "fp32":
        if S >= 64:
            assert e_mine <= 1e-3 and agree == 1.0 and abs(loss.item() - l64.item()) <= 1e-5 * abs(l64.item())
            # (single tensors of these ill-conditioned nets move by a few per cent even between the fp32 and the fp64 oracle:
            #  tools/probe_grad_error.py; the median is the tight bound)
            assert errs[worst] <= 1e-1 and np.median(list(errs.values())) <= 5e-3 and max(up_errs.values()) <= 1e-1
        else:
            assert e_mine <= 1e-2 and agree >= 0.9999 and abs(loss.item() - l64.item()) <= 1e-4 * abs(l64.item())
            # 32^3 (CPU emulation run): up1 / up2 sit on 2^3 / 4^3-voxel instance norms, where the fp32-vs-fp64 gradient of this
            # synthetic net is noise (0.3 - 0.5 for Bottleneck blocks); the two well-conditioned levels carry the check
            assert max(v for k, v in up_errs.items() if k.startswith(("up3.", "up4."))) <= 1e-1
    else:
        assert e_mine <= 2.0 * e_emul + 1e-3 and abs(loss.item() - l64.item()) <= 2e-2 * abs(l64.item())
