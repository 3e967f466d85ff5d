"""More §8 rows on the device: the wide (base 64) backbone of BASELINE configs[4], the full configs[0] extent
(1x64x128x128) against the CPU oracle, the remaining loss classes, and the poly LR + fused SGD interplay."""
import numpy as np
import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _model(base, dtype, algo, seed=0):
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(dtype)
    mm.set_conv_algo(algo)
    sd = O.synth_state_dict(base, 16, seed)
    m = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True, base=base).cuda()
    m.load_state_dict(sd)
    return m, sd


def _reset():
    import multimodal_pl_b200 as mm

    mm.set_conv_algo("auto")
    mm.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("dtype,algo,tol", [(torch.float32, "direct", 1e-5), (torch.bfloat16, "auto", 2e-2)])
def test_wide_backbone_base64(dtype, algo, tol):
    """configs[4] recipe (widths 64..512) at a small extent: logits / loss vs the oracle built from the same recipe."""
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial

    try:
        model, sd = _model(64, dtype, algo)
        model.train()
        x = O.synth_patch((1, 1, 16, 32, 32), 7, "mri")
        lab = O.synth_labels((1, 16, 32, 32), 8, 16, 32)
        w = [1.0] + [0.0] * 7 + [1.0] + [0.0] * 7
        ref = O.unet3d_forward(sd, x, base=64)
        ref_loss = O.partial_label_loss(ref, lab.squeeze(1), w).item()
        logits = model(x.cuda())[0]
        assert rel(logits, ref) < tol
        L = EDiceLoss_partial(16)(logits, lab.squeeze(1).cuda(), mask=[torch.tensor(w)])
        assert abs(L.item() - ref_loss) < max(tol, 1e-5) * max(1.0, ref_loss)
        L.backward()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    finally:
        _reset()


@pytest.mark.parametrize("dtype,algo,tol", [(torch.float32, "direct", 2e-5), (torch.bfloat16, "auto", 3e-2)])
def test_cfg1_full_extent_vs_oracle(dtype, algo, tol):
    """BASELINE configs[0]: batch 1, 1x64x128x128, 16 classes -- forward + loss at the full extent.
    bf16 tolerance: logits rel-L2 <= 3e-2 (31 layers of bf16 storage, relative rounding 2^-9 each, including the bf16
    27-tap expansion of the input image that feeds the tensor-core stem; measured 2.2e-2), loss rel <= 3e-2."""
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial

    try:
        model, sd = _model(32, dtype, algo)
        model.eval()
        x = O.synth_patch((1, 1, 64, 128, 128), 21, "ct")
        lab = torch.nn.functional.interpolate(O.synth_labels((1, 16, 32, 32), 22, 16, 32), size=(64, 128, 128),
                                              mode="nearest")
        w = [1.0, 0, 0, 0, 1.0] + [0.0] * 11
        with torch.no_grad():
            ref = O.unet3d_forward(sd, x)
            ref_loss = O.partial_label_loss(ref, lab.squeeze(1), w).item()
            logits = model(x.cuda())
            L = EDiceLoss_partial(16)(logits, lab.squeeze(1).cuda(), mask=[torch.tensor(w)]).item()
        assert rel(logits, ref) < tol
        assert abs(L - ref_loss) < max(tol, 1e-5) * max(1.0, ref_loss)
        if dtype == torch.float32:
            top2 = ref.topk(2, dim=1).values
            near_tie = (top2[:, 0] - top2[:, 1]) < 1e-4
            assert ((logits.cpu().argmax(1) != ref.argmax(1)) & ~near_tie).sum().item() == 0
    finally:
        _reset()


def test_dice_and_gated_losses_match_oracle():
    from multimodal_pl_b200.loss_functions.loss_partial import DiceLoss, EDiceLoss_full2, EDiceLoss_partial

    g = torch.Generator().manual_seed(4)
    z = torch.randn((2, 5, 4, 6, 6), generator=g)
    p = torch.softmax(z, 1)
    tgt = torch.randint(0, 5, (2, 4, 6, 6), generator=g).float()
    w = [0.5, 1.0, 0.0, 2.0, 1.0]
    gate = (torch.rand((2, 5, 4, 6, 6), generator=g) > 0.3)
    d = DiceLoss(5)
    assert abs(d(p.cuda(), tgt.cuda(), weight=w, softmax=False).item() - O.dice_loss_class(p, tgt, w).item()) < 1e-6
    assert abs(d(p.cuda(), tgt.cuda(), weight=w, softmax=False, mask=gate.cuda()).item() -
               O.dice_loss_class(p, tgt, w, gate).item()) < 1e-6
    # sigmoid variant of EDiceLoss_partial (soft_max=False)
    ps = torch.sigmoid(z)
    ref = O.dice_loss_class(ps, tgt, w)
    for c in range(5):
        ref = ref + torch.nn.functional.binary_cross_entropy(ps[:, c], (tgt == c).float()) * w[c]
    got = EDiceLoss_partial(5)(z.cuda(), tgt.cuda(), mask=[torch.tensor(w)] * 2, soft_max=False)
    assert abs(got.item() - ref.item()) < 1e-5
    # binary gated Dice (EDiceLoss_full2), the form get_loss uses: inputs [1,1,D,H,W], target [1,D,H,W], mask [1,1,D,H,W]
    x = torch.randn((1, 1, 4, 6, 6), generator=g)
    t = torch.rand((1, 4, 6, 6), generator=g)
    m = (torch.rand((1, 1, 4, 6, 6), generator=g) > 0.4).float()
    f2 = EDiceLoss_full2(2)
    for uce, sig in [(False, True), (False, False), (True, True)]:
        xin = x if sig else torch.sigmoid(x)
        a = f2(xin.cuda(), t.cuda(), uce=uce, mask=m.cuda(), sigmoid=sig).item()
        b = O.binary_gated_dice(xin, t, m, uce=uce, sigmoid=sig).item()
        assert abs(a - b) < 1e-6, (uce, sig)


@pytest.mark.parametrize("uce,sig", [(False, True), (False, False), (True, True), (True, False)])
def test_gated_dice_fused_kernel_gradients(uce, sig):
    """mmpl_masked_dice_{fwd,bwd}: value and gradients w.r.t. the score AND the soft target vs the oracle's autograd
    (EDiceLoss_full2 / DiceLoss._dice_loss, loss_partial.py:24-36, :150-170), ragged size, partial gate."""
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_full2

    g = torch.Generator().manual_seed(9)
    x = torch.randn((1, 1, 5, 7, 9), generator=g)
    if not sig:
        x = torch.sigmoid(x)
    t = torch.rand((1, 5, 7, 9), generator=g)
    m = (torch.rand((1, 1, 5, 7, 9), generator=g) > 0.4).float()
    xr, tr = x.clone().requires_grad_(True), t.clone().requires_grad_(True)
    ref = O.binary_gated_dice(xr, tr, m, uce=uce, sigmoid=sig)
    ref.backward()
    xd, td = x.cuda().requires_grad_(True), t.cuda().requires_grad_(True)
    got = EDiceLoss_full2(2)(xd, td, uce=uce, mask=m.cuda(), sigmoid=sig)
    got.backward()
    assert abs(got.item() - ref.item()) < 1e-6
    assert (xd.grad.cpu() - xr.grad).abs().max().item() < 1e-6 * max(1.0, xr.grad.abs().max().item()) + 1e-8
    assert (td.grad.cpu() - tr.grad).abs().max().item() < 1e-6 * max(1.0, tr.grad.abs().max().item()) + 1e-8
    # no gate == all voxels
    a = EDiceLoss_full2(2)(x.cuda(), t.cuda(), uce=uce, mask=None, sigmoid=sig).item()
    b = O.binary_gated_dice(x, t, None, uce=uce, sigmoid=sig).item()
    assert abs(a - b) < 1e-6


@pytest.mark.parametrize("tag", ["mixed", "none_supervised"])
def test_get_loss_refiner_branch_matches_reference_fixture(golden_dir, tag):
    """get_loss with a refiner output (reference losses.py:131-178): value and gradients w.r.t. the logits, the three
    attention maps and the refiner output vs tests/golden/get_loss_refine.npz, written by oracle/make_golden_get_loss.py
    from the UNMODIFIED reference function."""
    import importlib.util
    import os

    from multimodal_pl_b200.loss_functions.losses import get_loss

    spec = importlib.util.spec_from_file_location(
        "make_golden_get_loss", os.path.join(os.path.dirname(golden_dir), "..", "oracle", "make_golden_get_loss.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)                      # only its seeded input generator is used (no reference import)
    g = np.load(os.path.join(golden_dir, "get_loss_refine.npz"))
    output, target, attns, refine, deep = gen.inputs()
    leaves = [output.cuda().requires_grad_(True)] + [a.cuda().requires_grad_(True) for a in attns] + \
             [refine.cuda().requires_grad_(True)]
    label_t = [bool(v) for v in g[tag + ":label_t"]]
    loss, confi = get_loss(leaves[0], 0, [d.cuda() for d in deep], target.cuda(),
                           mask=[torch.from_numpy(g[tag + ":wmask"])], attns=leaves[1:4], refine_output=leaves[4],
                           label_t=label_t, aux_weight=0.7, weight_feature=0.3)
    loss.backward()
    assert abs(loss.item() - float(g[tag + ":loss"])) < 2e-6 and confi == float(g[tag + ":confi"])
    for name, t in zip(["output", "attn0", "attn1", "attn2", "refine"], leaves):
        ref = torch.from_numpy(g[tag + ":grad:" + name])
        got = t.grad.cpu()
        assert (got - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item()) + 1e-8, name
        assert ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item() < 1e-4, name


def test_unet3d_with_feam3_matches_reference_fixture(golden_dir):
    """unet3D_with_feam3 (reference unet3D.py:938-1190) on the fp32 exact path vs tests/golden/feam3.npz, written by
    oracle/make_golden_feam3.py from the UNMODIFIED reference model: the four train-mode outputs, gradients under a fixed
    objective (including that parameters the forward never uses get no gradient), eval-mode logits and the EMA class
    tokens after renew_token()."""
    import importlib.util
    import os

    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.unet3D import unet3D_with_feam3

    spec = importlib.util.spec_from_file_location(
        "make_golden_feam3", os.path.join(os.path.dirname(golden_dir), "..", "oracle", "make_golden_feam3.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)                       # seeded objective / constants only (no reference import)
    g = np.load(os.path.join(golden_dir, "feam3.npz"))
    mm.set_compute_dtype(torch.float32)
    mm.set_conv_algo("direct")
    try:
        sd, tokens = O.synth_feam3_state_dict(gen.CLASSES, gen.SEED)
        model = unet3D_with_feam3([1, 2, 2, 2, 2], num_classes=gen.CLASSES, weight_std=True).cuda()
        model.load_state_dict(sd)
        model.class_token1, model.class_token2, model.class_token3 = [t.clone() for t in tokens]
        model.train()
        x = O.synth_patch(gen.SHAPE, 1000 + gen.SEED, "ct").cuda()
        lab = O.synth_labels((gen.SHAPE[0],) + gen.SHAPE[2:], 2000 + gen.SEED, gen.CLASSES, 32).cuda()
        logits, attn, deep, feats = model(x, lab)
        assert len(attn) == len(deep) == len(feats) == 3

        def rel(a, b):
            a, b = a.detach().double().cpu(), torch.from_numpy(b).double()
            return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

        assert rel(logits, g["logits"]) < 1e-5
        for i in range(3):
            assert rel(attn[i], g[f"attn{i}"]) < 1e-4, i
            assert rel(deep[i], g[f"deep{i}"]) < 1e-5, i
            assert rel(feats[i], g[f"feat{i}"]) < 1e-5, i
        dys = gen.seeded_dys([tuple(logits.shape)] + [tuple(a.shape) for a in attn] + [tuple(d.shape) for d in deep])
        gen.objective(logits, attn, deep, dys).backward()
        params = dict(model.named_parameters())
        # whole-network gradients: 1e-2, the bound test_gpu_unet.py documents (a ReLU gate within fp32 rounding of zero
        # may resolve differently in two fp32 implementations; one flip moves upstream rel-L2 by ~2e-3)
        errs = {k: rel(params[k].grad, g["grad:" + k]) for k in gen.GRAD_KEYS}
        assert all(v < 1e-2 for v in errs.values()), errs
        assert (params["eam84.proj.weight"].grad is None) == bool(g["unused_grad_is_none"][0])
        model.renew_token(feats, lab)
        for i, t in enumerate([model.class_token1, model.class_token2, model.class_token3]):
            assert rel(t, g[f"token{i}"]) < 1e-5, i
        model.eval()
        with torch.no_grad():
            ev = model(x)
        assert torch.equal(ev, logits.detach()) == bool(g["eval_equals_train_logits"][0])
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


def test_unet3d_with_feam3_bf16_train_step_runs():
    """bf16 / tcgen05 path of the full train-loop model: finite outputs and gradients, logits close to the fixture."""
    import os

    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.loss_functions.losses import get_loss
    from multimodal_pl_b200.unet3D import unet3D_with_feam3

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feam3.npz"))
    mm.set_compute_dtype(torch.bfloat16)
    sd, tokens = O.synth_feam3_state_dict(16, 3)
    model = unet3D_with_feam3([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
    model.load_state_dict(sd)
    model.class_token1, model.class_token2, model.class_token3 = [t.clone() for t in tokens]
    model.train()
    x = O.synth_patch((1, 1, 16, 32, 32), 1003, "ct").cuda()
    lab = O.synth_labels((1, 16, 32, 32), 2003, 16, 32).cuda()
    logits, attn, deep, feats = model(x, lab)
    ref = torch.from_numpy(g["logits"])
    assert ((logits.float().cpu() - ref).norm() / ref.norm()).item() < 2e-2
    # attention maps (LayerNorm rows + folded 15-row classifier), deep-supervision heads and stored features on the bf16
    # path vs the fp32 reference fixture.  On this 16x32x32 input the 1/8-resolution level has 32 voxels per channel, so
    # its GroupNorm statistics amplify the bf16 storage noise of the 20 layers above it: the bound is per level (the strict
    # check of these heads is the fp32 fixture test above; block-level bf16 parity: test_gpu_parity_strict.py)
    errs = {}
    for i, tol in enumerate((0.35, 0.15, 0.08)):
        for name, got in (("attn", attn[i]), ("deep", deep[i]), ("feat", feats[i])):
            r = torch.from_numpy(g[f"{name}{i}"])
            errs[(name, i)] = ((got.float().cpu() - r).norm() / r.norm()).item()
            assert errs[(name, i)] < tol, errs
    loss, _ = get_loss(logits, 0, deep, lab, [torch.ones(16)])
    (loss + sum(a.float().mean() for a in attn) + sum(d.float().mean() for d in deep)).backward()
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), k
    assert model.eam84.kv.weight.grad is not None and model.deepout2[2].weight.grad is not None
    model.renew_token(feats, lab)
    assert torch.isfinite(model.class_token3).all()


def test_poly_lr_drives_fused_sgd():
    from multimodal_pl_b200.engine import FusedSGD
    from multimodal_pl_b200.utils import adjust_learning_rate, lr_poly

    p = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    opt = FusedSGD([p], lr=0.1, momentum=0.0, weight_decay=0.0)
    ref = torch.ones(1000)
    for epoch in range(3):
        lr = adjust_learning_rate(opt, epoch, 0.1, 10, 0.9)
        assert abs(lr - O.lr_poly(0.1, epoch, 10)) < 1e-12 and abs(lr - lr_poly(0.1, epoch, 10, 0.9)) < 1e-12
        opt.zero_grad()
        (p * 2.0).sum().backward()
        opt.step()
        ref = ref - lr * 2.0
    assert torch.allclose(p.detach().cpu(), ref, atol=1e-6)
