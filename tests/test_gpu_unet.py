"""Whole-network parity: unet3D_baseline (B200 kernels) + fused partial-label loss vs the reference's own outputs
(tests/golden/unet_*.npz, written by oracle/make_golden.py from the unmodified reference) and vs the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _run(tag, golden_dir, dtype, algo):
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(dtype)
    mm.set_conv_algo(algo)
    try:
        g = np.load(os.path.join(golden_dir, f"unet_{tag}.npz"))
        shape, seed = tuple(int(v) for v in g["shape"]), int(g["seed"])
        model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
        model.load_state_dict(O.synth_state_dict(32, 16, seed))
        model.train()
        x = O.synth_patch(shape, 1000 + seed, "ct" if seed == 0 else "mri")
        lab = O.synth_labels((shape[0],) + shape[2:], 2000 + seed, 16, 32)
        w16 = g["w16"].tolist()
        cmask = O.remap_unsupervised(lab, w16)
        logits, a, b = model(x.cuda(), cmask.cuda())
        assert a == [] and b == [] and logits.dtype == torch.float32
        crit = EDiceLoss_partial(16)
        L = crit(logits, cmask.squeeze(1).cuda(), mask=[torch.tensor(w16)] * shape[0], soft_max=True)
        L.backward()
        grads = {k: p.grad for k, p in model.named_parameters()}
        model.eval()
        with torch.no_grad():
            ev = model(x.cuda())
        assert torch.equal(ev, logits.detach())
        return g, logits.detach(), L.item(), grads
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("tag", ["b1", "b2"])
def test_unet_fp32_exact_path(golden_dir, tag):
    """fp32 path: logits rel-L2 <= 1e-5, loss abs <= 1e-5, identical argmax outside near-ties (top-2 gap of the
    reference < 1e-4), gradient norms within 5e-3 and gradient rel-L2 <= 1e-2.

    Why 1e-2 and not 1e-4 for whole-network gradients: a single ReLU gate whose pre-activation lies within fp32
    rounding of zero may resolve differently in two fp32 implementations; ONE flipped gate out of the 524 288
    activations of a full-resolution layer changes the rel-L2 of everything upstream by 2e-3 (measured against an
    fp64 run of the oracle, tools/debug_block.py).  Kernel-level gradients, where no gate is involved or none flips,
    are held to 1e-4 in test_gpu_kernels.py."""
    g, logits, loss, grads = _run(tag, golden_dir, torch.float32, "direct")
    ref = torch.from_numpy(g["logits"])
    assert rel(logits, ref) < 1e-5
    assert abs(loss - float(g["loss"])) < 1e-5
    am, amr = logits.cpu().argmax(1), ref.argmax(1)
    top2 = ref.topk(2, dim=1).values
    gap = top2[:, 0] - top2[:, 1]
    assert ((am != amr) & (gap >= 1e-4)).sum().item() == 0
    for k, gr in grads.items():
        s = g["grad:" + k]
        n = gr.double().norm().item()
        assert abs(n - s[0]) <= 5e-3 * max(s[0], 1e-6) + 1e-7, (k, n, s[0])
    for k in ["conv1.weight", "layer0.0.conv1.weight", "layer1.0.downsample.2.weight", "precls_conv.2.weight",
              "layer0.0.gn1.weight"]:
        assert rel(grads[k], torch.from_numpy(g["gradfull:" + k])) < 1e-2, k


@pytest.mark.parametrize("algo", ["direct", "auto"])
@pytest.mark.parametrize("tag", ["b1", "b2"])
def test_unet_bf16_path(golden_dir, tag, algo):
    """bf16 path (tcgen05 convs when algo='auto'): logits rel-L2 <= 2e-2, loss rel <= 1e-2, gradient norms within
    20 %; gradient direction: rel-L2 <= 5e-2 at the classifier, cosine >= 0.9 for the deepest tensors.

    bf16 activation storage (relative rounding 4e-3) flips about 0.3 % of the ReLU gates per GroupNorm layer w.r.t.
    the fp32 reference, so the per-tensor gradient rel-L2 grows from ~3e-2 next to the loss to ~0.3 at the stem of
    this randomly initialised 16x32x32 test network; the CUDA-core bf16 path ('direct') and the tcgen05 path show the
    same figures, i.e. this is the storage format, not the kernels (kernel-level bf16 checks: test_gpu_kernels.py)."""
    g, logits, loss, grads = _run(tag, golden_dir, torch.bfloat16, algo)
    ref = torch.from_numpy(g["logits"])
    assert rel(logits, ref) < 2e-2
    assert abs(loss - float(g["loss"])) < 1e-2 * float(g["loss"])
    for k in ["conv1.weight", "layer0.0.conv1.weight", "layer1.0.downsample.2.weight", "precls_conv.2.weight",
              "layer0.0.gn1.weight"]:
        r = torch.from_numpy(g["gradfull:" + k]).double().flatten()
        x = grads[k].double().cpu().flatten()
        if k == "precls_conv.2.weight":
            assert rel(grads[k], torch.from_numpy(g["gradfull:" + k])) < 5e-2, k
        assert (x @ r / (x.norm() * r.norm())).item() > 0.9, k
    bad = []
    for k, gr in grads.items():
        s = g["grad:" + k]
        n = gr.double().norm().item()
        if abs(n - s[0]) > 2e-1 * max(s[0], 1e-6) + 1e-6:
            bad.append((k, n, s[0]))
    assert not bad, bad
