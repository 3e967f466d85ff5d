"""Sliding-window inference + Dice on the device vs the reference's own outputs (tests/golden/sliding_window.npz) and
vs the CPU oracle with the full unet3D (fp32 exact path): identical argmax / Dice, blended logits to 1e-12 (fp64)."""
import os

import numpy as np
import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


def test_blend_and_dice_match_reference_fixture(golden_dir):
    from multimodal_pl_b200.evaluate import get_dice, predict_sliding, predict_sliding_dice

    g = np.load(os.path.join(golden_dir, "sliding_window.npz"))
    net = torch.nn.Conv3d(1, 5, 3, padding=1)
    net.weight.data.copy_(torch.from_numpy(g["sw_w"]))
    net.bias.data.copy_(torch.from_numpy(g["sw_b"]))
    net = net.cuda().eval()
    vol = O.synth_patch(tuple(int(v) for v in g["sw_vol_shape"]), 5, "ct").numpy()
    full = predict_sliding(None, [lambda im, tid: net(im)], vol, (8, 16, 16), 5, None)
    assert full.dtype == torch.float64 and tuple(full.shape) == tuple(g["sw_out"].shape)
    ref = torch.from_numpy(g["sw_out"])
    # same fp32 products, same float64 accumulation order per voxel as the reference; the conv itself runs on cuDNN
    # here vs oneDNN in the fixture, hence 1e-5 rather than bit-exact
    assert (full.cpu() - ref).abs().max().item() < 1e-5
    lab = torch.from_numpy(g["dice_labels"])
    dices, senc, spec, am = get_dice(full, lab, None, num_class=4)
    top2 = ref.topk(2, dim=1).values
    near_tie = (top2[:, 0] - top2[:, 1]) < 1e-4
    mism = (am.cpu().numpy().astype(np.uint8) != g["argmax"]) & ~near_tie.numpy()
    assert mism.sum() == 0
    assert np.allclose([float(d) for d in dices], g["dice"], atol=2e-4)
    assert np.allclose([float(d) for d in senc], g["senc"], atol=2e-4)
    assert np.allclose([float(d) for d in spec], g["spec"], atol=2e-4)
    # fused variant (never materialises the normalised volume) agrees with the two-step API exactly
    d2, s2, p2, am2 = predict_sliding_dice(None, [lambda im, tid: net(im)], vol, (8, 16, 16), 5, None, label=lab,
                                           num_class=4)
    assert torch.equal(am2.long(), am)
    assert np.allclose([float(d) for d in d2], [float(d) for d in dices], atol=0)


def test_get_dice_exact_counts():
    """Integer work is bit-exact: counts -> Dice equals the oracle's formula on random logits."""
    from multimodal_pl_b200.evaluate import get_dice

    g = torch.Generator().manual_seed(3)
    logits = torch.randn((2, 16, 7, 9, 11), generator=g)
    lab = torch.randint(0, 16, (2, 1, 7, 9, 11), generator=g).float()
    dices, senc, spec, am = get_dice(logits.cuda(), lab.cuda(), None, num_class=13)
    rd, rs, rp, ram = O.get_dice(logits, lab, num_class=13)
    assert torch.equal(am.cpu(), ram)
    assert np.allclose([float(d) for d in dices], [float(d) for d in rd], atol=1e-12)
    assert np.allclose([float(d) for d in senc], [float(d) for d in rs], atol=1e-12)
    assert np.allclose([float(d) for d in spec], [float(d) for d in rp], atol=1e-12)


def test_get_dice_with_atlas_prior():
    """get_dice(atlas=...) (reference evaluate_amos.py:142-151): cpred_l = softmax(x)[:, l+1] + 0.15 > 1 - atlas[:, l]."""
    from multimodal_pl_b200.evaluate import get_dice

    g = torch.Generator().manual_seed(5)
    logits = torch.randn((1, 6, 5, 7, 9), generator=g)
    lab = torch.randint(0, 6, (1, 5, 7, 9), generator=g).float()
    atlas = torch.rand((1, 5, 5, 7, 9), generator=g)
    dices, senc, spec, am = get_dice(logits.cuda(), lab.cuda(), None, atlas=atlas.cuda(), num_class=5)
    pr = torch.softmax(logits, 1)
    assert torch.equal(am.cpu(), pr.argmax(1))
    for l in range(5):
        cp = ((pr[:, l + 1] + 0.15) > (1 - atlas[:, l])).double().view(1, -1)
        t = (lab == (l + 1)).double().view(1, -1)
        inter = (cp * t).sum(1)
        assert abs(float(dices[l]) - float((2 * inter / (cp.sum(1) + t.sum(1) + 1)).mean())) < 1e-9
        assert abs(float(senc[l]) - float((inter / (t.sum(1) + 1)).mean())) < 1e-9
        assert abs(float(spec[l]) - float((inter / (cp.sum(1) + 1)).mean())) < 1e-9


def test_unet_sliding_window_fp32_argmax_identical():
    """Full path on a small volume (2x2x2 tiles of 16x32x32): fp32 exact kernels vs the CPU oracle; argmax identical
    outside near-ties (reference top-2 gap < 1e-4), Dice within 1e-3."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.evaluate import predict_sliding_dice
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.float32)
    mm.set_conv_algo("direct")
    try:
        sd = O.synth_state_dict(32, 16, 0)
        model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda().eval()
        model.load_state_dict(sd)
        vol = O.synth_patch((1, 1, 24, 40, 48), 11, "ct")
        lab = O.synth_labels((1, 24, 40, 48), 12, 16, 32)
        with torch.no_grad():
            ref = O.predict_sliding(lambda im: O.unet3d_forward(sd, im), vol.numpy(), (16, 32, 32), 16)
        rd, _, _, ram = O.get_dice(ref, lab, num_class=15)
        dices, _, _, am = predict_sliding_dice(None, [lambda im, tid: model(im)], vol.numpy(), (16, 32, 32), 16, None,
                                               label=lab, num_class=15)
        top2 = ref.topk(2, dim=1).values
        near_tie = (top2[:, 0] - top2[:, 1]) < 1e-4
        assert ((am.cpu().long() != ram) & ~near_tie).sum().item() == 0
        assert np.allclose([float(d) for d in dices], [float(d) for d in rd], atol=1e-3)
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


def test_tta_eight_flip_average_matches_oracle_and_graphed_equals_eager():
    """predict_sliding(tta=True) (reference evaluate_amos.py:247-255): mean over the identity and the seven axis-flip
    combinations.  fp32 exact kernels vs the CPU oracle (rel-L2 <= 1e-5), and the CUDA-graph network
    (engine.GraphedInference hands out ONE static output buffer that every replay overwrites) must give the same volume
    as the eager network -- the un-flipped prediction has to be copied before the next replay."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import GraphedInference
    from multimodal_pl_b200.evaluate import predict_sliding
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.float32)
    mm.set_conv_algo("direct")
    try:
        sd = O.synth_state_dict(32, 16, 3)
        model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda().eval()
        model.load_state_dict(sd)
        vol = O.synth_patch((1, 1, 24, 40, 32), 31, "mri")
        tile = (16, 32, 32)
        with torch.no_grad():
            ref = O.predict_sliding(lambda im: O.unet3d_forward(sd, im), vol.numpy(), tile, 16, tta=True)
            plain = O.predict_sliding(lambda im: O.unet3d_forward(sd, im), vol.numpy(), tile, 16, tta=False)
        eager = predict_sliding(None, [lambda im, tid: model(im)], vol.numpy(), tile, 16, None, tta=True)
        err = ((eager.cpu() - ref).norm() / ref.norm()).item()
        assert err < 1e-5, err
        assert ((plain - ref).norm() / ref.norm()).item() > 1e-3      # the network is not flip-equivariant: TTA matters
        net = GraphedInference(model, vol[:, :, :16, :32, :32].cuda().contiguous())
        graphed = predict_sliding(None, [net], vol.numpy(), tile, 16, None, tta=True)
        # identical kernels; only the order of the fp64 statistics atomics may differ between a replay and an eager run
        assert ((graphed - eager).norm() / eager.norm()).item() < 1e-6
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


def test_fused_classifier_blend_path_equals_generic_path():
    """predict_sliding_dice(acc_dtype=float32) with a bf16 unet3D_baseline takes the production path: classifier + Gaussian
    accumulation in ONE kernel into a depth-major accumulator (mmpl_cls_blend), vectorised argmax/Dice (mmpl_sw_finalize).
    Same MMA sequence and same tile order as classifier -> fp32 logits tile -> mmpl_sw_blend, so the argmax volume and
    the Dice counts must be identical; the CUDA-graph engine (one graph for all tiles, origin in device memory) likewise.
    Also: uint8 labels == float labels."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import GraphedSlidingWindow
    from multimodal_pl_b200.evaluate import predict_sliding_dice
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.bfloat16)
    sd = O.synth_state_dict(32, 16, 2)
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda().eval()
    model.load_state_dict(sd)
    vol = O.synth_patch((1, 1, 40, 72, 88), 41, "ct")
    lab = O.synth_labels((1, 40, 72, 88), 42, 16, 32)
    tile = (16, 32, 32)
    generic = predict_sliding_dice(None, [lambda im, tid: model(im)], vol, tile, 16, None, label=lab,
                                   acc_dtype=torch.float32, num_class=15)
    fused = predict_sliding_dice(None, [model], vol.numpy(), tile, 16, None, label=lab, acc_dtype=torch.float32,
                                 num_class=15)
    assert fused[3].shape == generic[3].shape and fused[3].dtype == torch.uint8
    assert torch.equal(fused[3], generic[3])
    assert [float(a) for a in fused[0]] == [float(a) for a in generic[0]]
    # tile batches: 36 tiles as single replays, as 7 x 5 + 1, as 4 x 8 + one remainder replay of 4 on one stream, and
    # (default) as 6 x 6 alternating between two streams -- every tile is still accumulated by its own launch in the
    # reference's order (the accumulation of a batch waits for the previous batch's), so nothing may change
    for tb, lanes in ((1, 2), (5, 2), (8, 1), (None, None)):
        eng = GraphedSlidingWindow(model, (40, 72, 88), tile, 16, tile_batch=tb, lanes=lanes)
        assert eng.tile_batch == (tb or 6) and eng.lanes == (lanes or 2) and eng.run == 36
        for _ in range(2):          # second volume: the accumulator is reset between volumes
            r0 = eng.tiles_replayed
            graphed = predict_sliding_dice(None, [eng], vol, tile, 16, None, label=lab.to(torch.uint8),
                                           acc_dtype=torch.float32, num_class=15)
            assert eng.tiles_replayed - r0 == 36
            assert torch.equal(graphed[3], generic[3])
            assert [float(a) for a in graphed[0]] == [float(a) for a in generic[0]]
        eng.close()
    # against the fp32 CPU oracle: bf16 logits flip only near-ties of the blended volume
    with torch.no_grad():
        ref = O.predict_sliding(lambda im: O.unet3d_forward(sd, im), vol.numpy(), tile, 16)
    top2 = ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 5e-2 * top2[:, 0].abs().clamp_min(1.0)
    assert (fused[3].cpu().long() != ref.argmax(1))[decided].float().mean().item() < 2e-3
