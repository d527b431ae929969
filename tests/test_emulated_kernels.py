"""CPU: the staged HBM-bound kernels (csrc/train_glue.cu, csrc/infer.cu, csrc/augment.cu) compiled for the HOST against the
execution-model shim tests/emul/cuda_emul.h and driven through the SAME host code and the SAME parity test bodies as on the GPU
(tests/test_widen_gpu.py), on CPU tensors.  This checks indexing, reductions, the union-find and the optimizer arithmetic of
the sources as they lie in csrc/ before their first run on a B200; it is test infrastructure (the product never loads the
emulated library) and says nothing about the tensor-core kernels, which cannot be emulated this way."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))

CPU = torch.device("cpu")


@pytest.fixture(scope="module")
def emu_handle():
    import build_emul
    from rsuper_b200 import _lib
    h = ctypes.CDLL(build_emul.build())
    names = []
    for name, (res, args) in _lib.SIGNATURES.items():
        if hasattr(h, name):
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
            names.append(name)
    assert len(names) >= 32, names
    return h, set(names)


@pytest.fixture
def emulated(monkeypatch, emu_handle):
    """rsuper_b200 host code -> emulated kernels on CPU tensors."""
    from rsuper_b200 import _lib, ops
    h, names = emu_handle
    real = _lib.lib()

    class Lib:
        def __getattr__(self, name):
            if name in names:
                return getattr(h, name)
            if name in ("rsb_conv3_n_tile", "rsb_conv3_packed_weight_bytes", "rsb_conv3_pack_plan", "rsb_conv3_wgrad_workspace_bytes",
                        "rsb_ball_workspace_bytes", "rsb_version"):
                return getattr(real, name)          # host-only planning functions of the real library
            raise AttributeError(f"{name} is not emulated (tensor-core / TMA kernels need the GPU)")

    lib = Lib()
    monkeypatch.setattr(ops, "lib", lambda: lib)
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(_lib, "check", lambda rc, what: (_ for _ in ()).throw(
        RuntimeError(f"rsuper_b200: {what} failed (rc={rc}): {h.rsb_last_error().decode()}")) if rc != 0 else None)
    monkeypatch.setattr(ops, "check", _lib.check)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_on_device", lambda t: True)
    return lib


import test_widen_gpu as W  # noqa: E402  (the GPU test bodies: plain functions of (device, ...))


_slow = pytest.mark.skipif(os.environ.get("RSB_EMUL_FULL") != "1", reason="slow under emulation; run with RSB_EMUL_FULL=1")


@pytest.mark.parametrize("with_ema,max_norm", [(True, 1.0), pytest.param(False, 1.0, marks=_slow), pytest.param(True, None, marks=_slow)])
def test_emulated_fused_clip_adamw_ema_matches_torch(emulated, with_ema, max_norm):
    W.test_fused_clip_adamw_ema_matches_torch(CPU, with_ema, max_norm)


def test_emulated_fused_optimizer_resume_and_gradless_parameters(emulated):
    W.test_fused_optimizer_resume_and_gradless_parameters(CPU)


def test_emulated_fused_clip_adamw_ema_matches_reference_golden(emulated):
    W.test_fused_clip_adamw_ema_matches_reference_golden(CPU)


@pytest.mark.parametrize("C,shape", [(2, (16, 16, 16)), (3, (8, 12, 20)), (9, (5, 7, 9)), (42, (6, 5, 7))])
def test_emulated_unpack_masks_bit_exact(emulated, C, shape):
    W.test_unpack_masks_bit_exact(CPU, C, shape)


def test_emulated_sliding_window_blend_matches_reference_golden(emulated, golden):
    W.test_sliding_window_blend_matches_reference_golden(CPU, golden)


@pytest.mark.parametrize("case", ["hand", "random_sparse", "random_dense", "empty", "full", "snake",
                                  pytest.param("blobs", marks=pytest.mark.skipif(os.environ.get("RSB_EMUL_FULL") != "1", reason="slow under emulation"))])
def test_emulated_connected_components_bit_exact(emulated, case):
    W.test_connected_components_bit_exact(CPU, case)


def test_emulated_organ_gating_matches_oracle(emulated):
    W.test_organ_gating_matches_oracle(CPU)


@_slow
def test_emulated_intensity_augmentations_match_reference_golden(emulated):
    W.test_intensity_augmentations_match_reference_golden(CPU)


def test_emulated_public_dice_loss_multiclass_vs_oracle(emulated):
    W.test_public_dice_loss_multiclass_vs_oracle(CPU)


# ---- validation of the shim itself: kernels that are ALREADY green on the B200 (profiles/r01*_gpu_tests.log) must also be
# green when their source runs under the shim, against the same oracle and the same golden vectors of the real reference ----
def test_shim_reproduces_gpu_verified_seg_loss(emulated, golden):
    import test_kernels_gpu as K
    K.test_seg_loss_forward_backward_vs_oracle(CPU)
    K.test_seg_loss_matches_golden(CPU, golden)


def test_shim_reproduces_gpu_verified_dilation(emulated, golden):
    import test_kernels_gpu as K
    K.test_dilation_bit_exact(CPU, golden)


def test_emulated_widened_edge_cases(emulated):
    W.test_widened_edge_cases(CPU)


# ---- the HBM-bound satellites of the train step (csrc/elementwise.cu: GPU-verified) under the shim: CPU regression net for
# edits to these kernels between GPU visits ----
def test_shim_runs_gpu_verified_elementwise_kernels(emulated):
    import test_kernels_gpu as K
    K.test_layout_and_stats(CPU)
    K.test_norm_act_split_and_slices(CPU)
    K.test_instnorm_backward_apply(CPU)
    K.test_maxpool_forward_backward_with_ties(CPU)
    for dims in (((4, 4, 4), (8, 8, 8)), ((2, 2, 2), (4, 4, 4)), ((8, 6, 4), (16, 12, 8))):
        K.test_upsample_trilinear(CPU, dims)


@pytest.mark.skipif(os.environ.get("RSB_EMUL_FULL") != "1", reason="slow under emulation; run with RSB_EMUL_FULL=1")
def test_shim_runs_gpu_verified_stem_and_head(emulated):
    import test_kernels_gpu as K
    K.test_stem_and_head(CPU)


full = pytest.mark.skipif(os.environ.get("RSB_EMUL_FULL") != "1",
                          reason="~2 min of emulated report-loss kernels: developer regression net, run with RSB_EMUL_FULL=1")


@full
def test_shim_runs_gpu_verified_report_losses(emulated, golden):
    """Volume loss, isolate_tumor (bit-exact pseudo masks), GWRP weights, Ball loss and the calculate_loss dicts — the whole
    report-supervised loss path (csrc/report_loss.cu + seg_loss.cu + morph.cu) on the CPU."""
    import test_report_losses_gpu as R
    R.test_volume_loss(CPU, golden)
    R.test_isolate_tumor_bit_exact(CPU, golden)
    R.test_gwrp_weights(CPU, golden)
    R.test_ball_loss(CPU, golden, dict())
    R.test_calculate_loss_dicts(CPU, golden)


@pytest.mark.parametrize("tag", ["merged", pytest.param("single", marks=_slow)])
def test_emulated_lesion_group_max_merge(emulated, tag):
    import test_report_losses_gpu as R
    R.test_lesion_group_max_merge(CPU, tag)


@full
def test_emulated_assemble_batch_feeds_calculate_loss(emulated):
    W.test_assemble_batch_feeds_calculate_loss(CPU)


def test_emulated_capturable_optimizer_equals_eager_optimizer(emulated):
    W.test_capturable_optimizer_equals_eager_optimizer(CPU)


# ---- the WHOLE train step on the CPU: real host code + real sources of every HBM-bound kernel + tests/emul/conv3_double.cpp
# standing in for the tensor-core kernels (naive loops behind the same C-ABI) ----
def test_emulated_unet_logits_vs_reference_golden(emulated, golden):
    """tests/test_unet_gpu.py::test_logits_vs_reference_golden on the CPU: the engine's forward orchestration (packing plan,
    concat slices, statistics plumbing, split-precision operands) against the REAL reference's recorded logits."""
    import test_unet_gpu as U
    U.test_logits_vs_reference_golden(CPU, golden)


def test_emulated_train_step_vs_oracle(emulated):
    """Forward + loss + backward of the B200UNet engine (parity mode) + B200AdamW on CPU tensors vs the oracle's autograd:
    logits, loss, every parameter gradient, and one optimizer step."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict, unet_forward
    from rsuper_b200 import losses
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.unet import B200UNet
    classes = ["organ", "pancreatic_lesion"]
    sd = synthetic_state_dict(8, 2)
    x = synthetic_image(1, 32, 32, 32, seed=3)
    lab = synth.make_batch(["mask"], classes, (32, 32, 32), seed=2)["label"]
    args = LR.default_args(report_volume_loss_basic=0.0)
    net = B200UNet(1, 8, num_classes=2, precision="fp32")
    net.load_state_dict(sd)
    out = net(x)
    loss = losses.calculate_loss(out, lab, None, args, None, None, None, None, classes)["overall"]
    loss.backward()
    ref_p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_logits = unet_forward(x, ref_p)
    ref_loss = LR.seg_loss(ref_logits, lab.float(), torch.ones_like(ref_logits))
    ref_loss.backward()
    e = ((out["segmentation"] - ref_logits).abs().max() / ref_logits.abs().max()).item()
    assert e <= 1e-3 and abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item()), (e, loss.item(), ref_loss.item())
    worst = 0.0
    for k, p in net.named_parameters():
        g, r = p.grad, ref_p[k].grad
        assert g is not None and torch.isfinite(g).all(), k
        worst = max(worst, ((g - r).norm() / (r.norm() + 1e-12)).item())
    print(f"[emulated step] logits rel {e:.2e}, worst relative gradient error {worst:.2e}")
    assert worst <= 5e-2          # a 32^3 patch normalises over 2^3 voxels at the bottom: the GPU tests use the same loose bar there
    before = [p.detach().clone() for p in net.parameters()]
    opt = B200AdamW(net.parameters(), lr=6e-4, weight_decay=0.05, max_norm=1.0)
    opt.step()
    moved = max((a - b).abs().max().item() for a, b in zip(before, [p.detach() for p in net.parameters()]))
    assert 1e-5 < moved <= 6e-4 * 1.1


@full
def test_emulated_singleconv_unet_vs_reference_and_oracle(emulated, golden):
    """block='SingleConv' engine (post-activation path, act_backward_stats) end to end on the CPU."""
    import test_unet_gpu as U
    U.test_singleconv_unet_vs_reference_and_oracle(CPU, golden, "fp32")


def test_emulated_bottleneck_unet_vs_reference_and_oracle(emulated, monkeypatch):
    """block='Bottleneck' engine (centre-tap 1x1x1 convolutions, summed data gradients ahead of the activation mask) end to end
    on the CPU against the real reference's recorded run and the fp64 oracle."""
    import test_unet_gpu as U
    monkeypatch.setenv("RSB_TEST_SINGLE_S", "32")
    U.test_bottleneck_unet_vs_reference_and_oracle(CPU, "fp32")


@pytest.mark.parametrize("block", [pytest.param("BasicBlock", marks=_slow), pytest.param("Bottleneck", marks=_slow)])
def test_emulated_transposed_conv_upsampling_variant(emulated, monkeypatch, block):
    """up_mode='transposed' (1x1x1 conv to 8 C channels + depth-to-space, space-to-depth + summed-bias backward) on the CPU."""
    import test_unet_gpu as U
    monkeypatch.setenv("RSB_TEST_SINGLE_S", "32")
    U.test_transposed_conv_upsampling_variant(CPU, block, "fp32")


@full
def test_emulated_static_gradient_steps_equal_fresh_gradient_steps(emulated):
    """Precondition of GraphedTrainStep (rsuper_b200/graph_step.py): the captured body starts with
    zero_grad(set_to_none=False) so that gradients keep their addresses; autograd then ACCUMULATES the engine's gradients into
    the zeroed buffers.  Gradients obtained that way (buffers pre-filled with garbage) must equal fresh gradients — compared
    directly, because Adam's update is invariant to a gradient scale — and the optimizer's device table must not be rebuilt
    between such steps."""
    from oracle import losses_ref as LR
    from oracle import synth
    from oracle.unet_ref import synthetic_image, synthetic_state_dict
    from rsuper_b200 import losses
    from rsuper_b200.optim import B200AdamW
    from rsuper_b200.unet import B200UNet
    classes = ["organ", "pancreatic_lesion"]
    x = synthetic_image(1, 32, 32, 32, seed=3)
    lab = synth.make_batch(["mask"], classes, (32, 32, 32), seed=2)["label"]
    args = LR.default_args(report_volume_loss_basic=0.0)
    args.nan_check = False
    grads = []
    for static in (False, True):
        net = B200UNet(1, 8, num_classes=2, precision="bf16")
        net.load_state_dict(synthetic_state_dict(8, 2))
        params = list(net.parameters())
        opt = B200AdamW(params, lr=6e-4, weight_decay=0.05, max_norm=1.0, capturable=True)
        if static:
            for p in params:
                p.grad = torch.full_like(p, 7.0)         # stale content that zero_grad(set_to_none=False) must clear
        tables = []
        for step in range(2):
            opt.zero_grad(set_to_none=not static)
            opt.prepare_step()
            loss = losses.calculate_loss(net(x), lab, None, args, None, None, None, None, classes)["overall"]
            loss.backward()
            if step == 0:
                grads.append([p.grad.detach().clone() for p in params])
            opt.step()
            tables.append(opt._tables[0][1].data_ptr())
        if static:
            assert tables[0] == tables[1]            # same gradient addresses -> the device table was not rebuilt
        assert opt.global_step == 2
    for i, (a, b) in enumerate(zip(*grads)):
        err = ((a - b).norm() / (a.norm() + 1e-20)).item()
        assert err <= 0.15, (i, err)                 # run-to-run bf16 + atomics-order noise on this 2^3-bottom toy net is ~2 %; a stale or doubled gradient is >= 100 % off


# ---- MedFormer voxel-side kernels (csrc/medformer.cu) and the module built on them -----------------------------------------
def test_emulated_medformer_dwconv_se_and_softmax_pool(emulated):
    """Depthwise 3x3x3 conv (forward, data gradient, weight gradient; channel counts spanning several chunks), SEBlock scale /
    dot, SemanticMapGeneration's softmax-over-voxels pooling forward and backward — against torch in fp64."""
    import test_medformer_gpu as MF
    MF.test_medformer_dwconv_kernels_vs_torch(CPU)
    MF.test_medformer_se_scale_and_dot_vs_torch(CPU)
    MF.test_medformer_softmax_pool_vs_torch(CPU)


def test_emulated_medformer_biattention(emulated):
    """BidirectionAttention core: both softmaxes of one logit matrix, both einsums and all four gradients (warp-per-voxel
    kernels run as real threads: shuffles, shared-memory staging, block and global atomics)."""
    import test_medformer_gpu as MF
    MF.test_medformer_biattention_vs_torch(CPU)


@_slow
def test_emulated_medformer_blocks_and_whole_model(emulated):
    """Every composite block forward + backward against the fp64 oracle, then the whole B200MedFormer against the real
    reference's recorded logits / deep-supervision head / loss and the fp64 oracle's gradients (~5 min under emulation)."""
    import test_medformer_gpu as MF
    MF.test_medformer_blocks_forward_backward_vs_oracle(CPU, "fp32")
    MF.test_medformer_vs_reference_golden_and_oracle(CPU, "fp32")
