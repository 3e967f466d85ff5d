"""Full-size checks (BASELINE.json configs[1] / configs[3] extents) through size-independent properties and sampled
voxels: tile edges, last partial tiles and 64-bit indexing only show up at these sizes."""
import numpy as np
import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


def test_partial_loss_full_cfg2_size_vs_oracle():
    """[2,16,64,192,192] logits (75 M values): fused loss == CPU oracle to 2e-6 rel; zero weights -> exactly 0 with a
    zero gradient; the loss is linear in the class-weight vector."""
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial

    g = torch.Generator().manual_seed(0)
    z = (2.0 * torch.randn((2, 16, 64, 192, 192), generator=g))
    lab = O.synth_labels((2, 16, 48, 48), 3, 16, 32)
    lab = torch.nn.functional.interpolate(lab, size=(64, 192, 192), mode="nearest").squeeze(1).contiguous()
    w = [1.0, 0, 0, 0, 1.0] + [0.0] * 11
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = O.partial_label_loss(z, lab, w).item()
    zc, lc = z.cuda().requires_grad_(True), lab.cuda()
    crit = EDiceLoss_partial(16)
    L = crit(zc, lc, mask=[torch.tensor(w)] * 2)
    assert abs(L.item() - ref) <= 2e-6 * max(1.0, abs(ref)), (L.item(), ref)
    L.backward()
    assert torch.isfinite(zc.grad).all() and zc.grad.abs().max().item() > 0
    # softmax gradients sum to zero over the class axis
    assert zc.grad.sum(dim=1).abs().max().item() < 1e-9
    L0 = crit(zc.detach().requires_grad_(True), lc, mask=[torch.zeros(16)] * 2)
    assert L0.item() == 0.0
    w2 = [0.0, 1.0, 0.5] + [0.0] * 13
    La = crit(zc.detach(), lc, mask=[torch.tensor(w2)] * 2).item()
    Lab = crit(zc.detach(), lc, mask=[torch.tensor([a + b for a, b in zip(w, w2)])] * 2).item()
    assert abs(Lab - (L.item() + La)) < 5e-6 * max(1.0, abs(Lab))


@pytest.mark.parametrize("cin,cout,stride,dims", [(32, 32, 1, (64, 192, 192)), (32, 64, 2, (64, 192, 192)),
                                                   (64, 64, 1, (32, 96, 96)), (256, 256, 1, (4, 12, 12))])
def test_tcgen05_conv_sampled_voxels_full_size(cin, cout, stride, dims):
    """tcgen05 conv at the real layer extents of configs[1]: 200 sampled output voxels (all 8 corners, faces, tile
    seams at multiples of 16/8/TD, random interior) against an fp64 dot product of the same bf16 operands."""
    import multimodal_pl_b200 as mm

    mm.set_compute_dtype(torch.bfloat16)
    mm.set_conv_algo("tcgen05")
    try:
        D, H, W = dims
        g = torch.Generator().manual_seed(1)
        x = torch.randn((2, D, H, W, cin), generator=g).bfloat16()
        w = torch.randn((cout, cin, 3, 3, 3), generator=g)
        xd = x.cuda().permute(0, 4, 1, 2, 3)
        y = mm.ops.ws_conv3d(xd, w.cuda(), stride, True, None)
        Do, Ho, Wo = y.shape[2:]
        what = O.ws_weight(w).bfloat16().double()                       # the kernel consumes bf16 standardised weights
        rng = np.random.RandomState(0)
        pts = [(n, d, h, w_) for n in (0, 1) for d in (0, Do - 1) for h in (0, Ho - 1) for w_ in (0, Wo - 1)]
        for _ in range(60):
            pts.append((rng.randint(2), rng.randint(Do), rng.randint(Ho), rng.randint(Wo)))
        for _ in range(60):   # tile seams
            pts.append((rng.randint(2), min(Do - 1, 4 * rng.randint(1 + Do // 4)), min(Ho - 1, 16 * rng.randint(1 + Ho // 16)),
                        min(Wo - 1, 8 * rng.randint(1 + Wo // 8))))
        for _ in range(60):
            pts.append((rng.randint(2), max(0, min(Do - 1, 4 * rng.randint(1 + Do // 4) - 1)),
                        max(0, min(Ho - 1, 16 * rng.randint(1 + Ho // 16) - 1)), max(0, min(Wo - 1, 8 * rng.randint(1 + Wo // 8) - 1))))
        yc = y.float().cpu()
        xf = x.double()
        worst = 0.0
        for n, d, h, w_ in pts:
            acc = torch.zeros(cout, dtype=torch.float64)
            for kd in range(3):
                for kh in range(3):
                    for kw in range(3):
                        zd, zh, zw = d * stride + kd - 1, h * stride + kh - 1, w_ * stride + kw - 1
                        if 0 <= zd < D and 0 <= zh < H and 0 <= zw < W:
                            acc += what[:, :, kd, kh, kw] @ xf[n, zd, zh, zw]
            got = yc[n, :, d, h, w_].double()
            err = ((got - acc).abs() / (acc.abs().mean() + 1e-3)).max().item()
            worst = max(worst, err)
        assert worst < 4e-2, worst      # bf16 output rounding (2^-8) relative to the typical magnitude
    finally:
        mm.set_conv_algo("auto")


def test_sliding_window_cfg4_partition_of_unity():
    """configs[3]: 300x512x512 volume, 64x192x192 tiles -> 96 tiles.  With a pointwise network f(x) the Gaussian
    blend must return f(x) itself (sum g f / sum g), so the fused argmax/Dice equals the direct computation exactly."""
    from multimodal_pl_b200.evaluate import predict_sliding_dice, tile_origins

    D, H, W, C = 300, 512, 512, 16
    g = torch.Generator().manual_seed(5)
    vol = torch.randn((1, 1, D, H, W), generator=g)
    a = torch.randn(C, generator=g).cuda().view(1, C, 1, 1, 1)
    b = torch.randn(C, generator=g).cuda().view(1, C, 1, 1, 1)
    calls = []

    def net(img, task_id):
        calls.append(tuple(img.shape))
        return img * a + b

    lab = torch.randint(0, C, (1, 1, D, H, W), generator=g).float()
    assert len(tile_origins((1, 1, D, H, W), (64, 192, 192))) == 96
    dices, senc, spec, am = predict_sliding_dice(None, [net], vol.numpy(), (64, 192, 192), C, None, label=lab,
                                                 num_class=15)
    assert len(calls) == 96 and all(c == (1, 1, 64, 192, 192) for c in calls)
    direct = (vol.cuda() * a + b)
    top2 = direct.double().topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-6
    ref_am = direct.argmax(1)
    assert ((am.long() != ref_am) & safe).sum().item() == 0
    labc = lab.cuda().squeeze(1)
    for l in (1, 7, 15):
        p, t = (ref_am == l), (labc == l)
        ref_d = 2.0 * (p & t).sum().double() / (p.sum().double() + t.sum().double() + 1)
        assert abs(float(dices[l - 1]) - ref_d.item()) < 1e-6


@pytest.mark.parametrize("shape", [(1, 1, 48, 176, 144), (3, 1, 32, 80, 112)])
def test_unet_ragged_patch_bf16_vs_fp32_paths(shape):
    """Patch sizes that are not multiples of the 4x16x8 conv tiles (only of the network's 16x down-sampling) and an odd
    batch: the bf16 tcgen05 path agrees with the fp32 exact CUDA-core path (itself pinned to the reference at 1e-5 in
    test_gpu_unet.py) within the bf16 tolerance, forward and loss; all parameter gradients finite."""
    import mmpl_oracle as O
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    sd = O.synth_state_dict(32, 16, 2)
    x = O.synth_patch(shape, 11, "ct").cuda()
    lab = O.synth_labels((shape[0],) + shape[2:], 12, 16, 32).cuda()
    w = [torch.ones(16)] * shape[0]
    outs = {}
    try:
        for dt, algo in ((torch.float32, "direct"), (torch.bfloat16, "auto")):
            mm.set_compute_dtype(dt)
            mm.set_conv_algo(algo)
            model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
            model.load_state_dict(sd)
            model.train()
            logits = model(x, lab)[0]
            loss = EDiceLoss_partial(16)(logits, lab.squeeze(1), mask=w)
            loss.backward()
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
            outs[dt] = (logits.detach().float(), loss.item())
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)
    ref, got = outs[torch.float32], outs[torch.bfloat16]
    assert ((got[0] - ref[0]).norm() / ref[0].norm()).item() < 2e-2
    assert abs(got[1] - ref[1]) < 1e-2 * abs(ref[1])
