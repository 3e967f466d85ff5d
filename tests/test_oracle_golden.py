"""Pins oracle/mmpl_oracle.py against fixtures produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import mmpl_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_ws_conv_matches_reference(golden_dir):
    g = _load(golden_dir, "ws_conv.npz")
    for name, k in [("stem", 3), ("c3", 3), ("c1", 1)]:
        y = O.ws_conv3d(torch.from_numpy(g[name + "_x"]), torch.from_numpy(g[name + "_w"]), 1, k // 2)
        assert torch.equal(y, torch.from_numpy(g[name + "_y"])), name   # same ATen ops, same order -> bit-exact


def test_ws_analytic():
    w = torch.randn(8, 4, 3, 3, 3)
    s = O.ws_weight(w).view(8, -1)
    assert torch.allclose(s.mean(1), torch.zeros(8), atol=1e-6)
    assert torch.allclose(s.var(1), torch.ones(8), atol=1e-5)
    # constant filter: centred weight is exactly 0 -> 0 / sqrt(1e-12) = 0
    assert torch.equal(O.ws_weight(torch.full((2, 3, 3, 3, 3), 0.7)), torch.zeros(2, 3, 3, 3, 3))


def test_partial_loss_matches_reference(golden_dir):
    g = _load(golden_dir, "partial_loss.npz")
    for name in ["ct_one_organ", "mri_bg_only", "all_ones", "all_zero", "c4_frac"]:
        for uce in (True, False):
            z = torch.from_numpy(g[name + "_z"]).requires_grad_(True)
            L = O.partial_label_loss(z, torch.from_numpy(g[name + "_t"]), g[name + "_w"].tolist(), uce=uce)
            tag = f"{name}_uce{int(uce)}"
            assert abs(L.item() - float(g[tag + "_loss"])) <= 1e-6 * max(1.0, abs(float(g[tag + "_loss"]))), tag
            if L.requires_grad:
                gz, = torch.autograd.grad(L, z)
                assert torch.allclose(gz, torch.from_numpy(g[tag + "_grad"]), rtol=1e-5, atol=1e-9), tag


def test_partial_loss_closed_form_sums(golden_dir):
    """The four per-class sums reassemble to the reference loss."""
    g = _load(golden_dir, "partial_loss.npz")
    for name in ["ct_one_organ", "all_ones", "c4_frac"]:
        z, t, w = g[name + "_z"], g[name + "_t"], g[name + "_w"].astype(np.float64)
        s = O.partial_label_loss_sums(z, t)
        C, nv = z.shape[1], t.size
        dice = (w * (1 - (2 * s["I"] + 1e-5) / (s["Z"] + s["Y"] + 1e-5))).sum() / C
        ce = (w * s["E"]).sum() / nv
        assert abs(dice + ce - float(g[name + "_uce1_loss"])) < 2e-6 * max(1.0, dice + ce), name   # fp32 reference vs f64
        assert abs(dice - float(g[name + "_uce0_loss"])) < 2e-6, name


def test_partial_loss_saturated_clamp(golden_dir):
    g = _load(golden_dir, "partial_loss.npz")
    z = torch.zeros((1, 4, 2, 2, 2))
    z[:, 0] = 200.0
    L = O.partial_label_loss(z, torch.ones((1, 2, 2, 2)), [1, 1, 1, 1])
    assert abs(L.item() - float(g["saturated_loss"])) < 1e-5


def test_unet_forward_backward_matches_reference(golden_dir):
    for tag in ["b1", "b2"]:
        g = _load(golden_dir, f"unet_{tag}.npz")
        shape, seed = tuple(int(v) for v in g["shape"]), int(g["seed"])
        sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(32, 16, seed).items()}
        x = O.synth_patch(shape, 1000 + seed, "ct" if seed == 0 else "mri")
        lab = O.synth_labels((shape[0],) + shape[2:], 2000 + seed, 16, 32)
        w16 = g["w16"].tolist()
        cmask = O.remap_unsupervised(lab, w16)
        logits = O.unet3d_forward(sd, x)
        ref = torch.from_numpy(g["logits"])
        assert torch.allclose(logits, ref, rtol=1e-4, atol=1e-4), (logits - ref).abs().max()
        L = O.partial_label_loss(logits, cmask.squeeze(1), w16)
        assert abs(L.item() - float(g["loss"])) < 1e-5
        L.backward()
        for k, p in sd.items():
            s = g["grad:" + k]
            got = p.grad.double().flatten()
            assert abs(got.norm().item() - s[0]) <= 1e-3 * max(s[0], 1e-6) + 1e-7, k
        for k in ["conv1.weight", "layer0.0.conv1.weight", "layer1.0.downsample.2.weight", "precls_conv.2.weight",
                  "layer0.0.gn1.weight"]:
            ref = torch.from_numpy(g["gradfull:" + k])
            rel = (sd[k].grad - ref).norm() / ref.norm()
            assert rel < 1e-3, (k, rel)


def test_sliding_window_matches_reference(golden_dir):
    g = _load(golden_dir, "sliding_window.npz")
    assert np.array_equal(O.gaussian_importance((8, 16, 16)), g["gauss_8_16_16"])
    big = O.gaussian_importance((64, 192, 192))
    st = g["gauss_big_stats"]
    assert big.max() == st[0] and big.min() == st[1] and (big == 0).sum() == st[3]
    assert np.array_equal(big[32, 96, :], g["gauss_big_line"])
    net = torch.nn.Conv3d(1, 5, 3, padding=1)
    net.weight.data.copy_(torch.from_numpy(g["sw_w"]))
    net.bias.data.copy_(torch.from_numpy(g["sw_b"]))
    vol = O.synth_patch(tuple(int(v) for v in g["sw_vol_shape"]), 5, "ct").numpy()
    with torch.no_grad():
        full = O.predict_sliding(lambda im: net(im), vol, (8, 16, 16), 5)
    assert np.allclose(full.numpy(), g["sw_out"], rtol=0, atol=1e-12)
    dices, senc, spec, am = O.get_dice(full, torch.from_numpy(g["dice_labels"]), num_class=4)
    assert np.array_equal(am.numpy().astype(np.uint8), g["argmax"])
    assert np.allclose([float(d) for d in dices], g["dice"], atol=1e-7)
    assert np.allclose([float(d) for d in senc], g["senc"], atol=1e-7)
    assert np.allclose([float(d) for d in spec], g["spec"], atol=1e-7)


def test_tile_grid_cfg4(golden_dir):
    g = _load(golden_dir, "sliding_window.npz")
    grid = O.tile_grid((300, 512, 512), (64, 192, 192))
    assert len(grid) == int(np.prod(g["cfg4_tiles"])) == 96
    assert sorted({d for d, _, _ in grid}) == [0, 48, 96, 144, 192, 236]
    assert sorted({y for _, y, _ in grid}) == [0, 144, 288, 320]
    assert O.tile_starts(40, 32, 24) == [0, 8]


def test_supervise_mask_adapter(tmp_path):
    p = tmp_path / "m.csv"
    p.write_text('name,mask\namos_0001.nii.gz,"[0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]"\n'
                 'amos_0507.nii.gz,"[0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]"\n')
    t = O.read_supervise_mask(str(p))
    assert t["amos_0001"] == [1.0, 0, 0, 0, 1.0] + [0.0] * 11 and len(t["amos_0507"]) == 16
    lab = torch.tensor([0., 4., 5., 4., 1.])
    assert O.remap_unsupervised(lab, t["amos_0001"]).tolist() == [0, 4, 0, 4, 0]


def test_sgd_matches_torch():
    p = torch.randn(100, requires_grad=True)
    opt = torch.optim.SGD([p], lr=0.01, momentum=0.9, weight_decay=1e-4)
    q, buf = p.detach().clone(), None
    for _ in range(3):
        g = torch.randn(100)
        p.grad = g.clone()
        opt.step()
        q, buf = O.sgd_step(q, g, buf, 0.01)
    assert torch.allclose(p.detach(), q, atol=1e-7)
    assert abs(O.lr_poly(0.01, 250, 500) - 0.01 * 0.5 ** 0.9) < 1e-12


def test_feam3_oracle_matches_reference_fixture(golden_dir):
    """O.unet3d_feam3_forward (restatement of unet3D_with_feam3.forward, unet3D.py:1095-1190) vs tests/golden/feam3.npz
    written from the unmodified reference model: logits, attention maps, deep-supervision maps, stored features."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "make_golden_feam3", os.path.join(os.path.dirname(golden_dir), "..", "oracle", "make_golden_feam3.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(golden_dir, "feam3.npz"))
    sd, tokens = O.synth_feam3_state_dict(gen.CLASSES, gen.SEED)
    x = O.synth_patch(gen.SHAPE, 1000 + gen.SEED, "ct")
    with torch.no_grad():
        logits, attn, deep, feats = O.unet3d_feam3_forward(sd, tokens, x)

    def rel(a, b):
        b = torch.from_numpy(b).double()
        return ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()

    assert rel(logits, g["logits"]) < 1e-5
    for i in range(3):
        assert rel(attn[i], g[f"attn{i}"]) < 1e-5 and rel(deep[i], g[f"deep{i}"]) < 1e-5 and rel(feats[i], g[f"feat{i}"]) < 1e-5


@pytest.mark.parametrize("tag", ["mixed", "none_supervised"])
def test_get_loss_refine_oracle_matches_reference_fixture(golden_dir, tag):
    """O.get_loss_refine (restatement of the refiner branch of get_loss, losses.py:107-178) vs the value and the gradients
    the unmodified reference function produced (tests/golden/get_loss_refine.npz)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "make_golden_get_loss", os.path.join(os.path.dirname(golden_dir), "..", "oracle", "make_golden_get_loss.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(golden_dir, "get_loss_refine.npz"))
    output, target, attns, refine, deep = gen.inputs()
    leaves = [output.requires_grad_(True)] + [a.requires_grad_(True) for a in attns] + [refine.requires_grad_(True)]
    loss = O.get_loss_refine(leaves[0], deep, target, g[tag + ":wmask"].tolist(), leaves[1:4], leaves[4],
                             [bool(v) for v in g[tag + ":label_t"]], aux_weight=0.7, weight_feature=0.3)
    loss.backward()
    assert abs(loss.item() - float(g[tag + ":loss"])) < 1e-6
    for name, t in zip(["output", "attn0", "attn1", "attn2", "refine"], leaves):
        ref = torch.from_numpy(g[tag + ":grad:" + name])
        assert (t.grad - ref).abs().max().item() < 1e-6 * max(1.0, ref.abs().max().item()) + 1e-9, name


def test_bf16_storage_emulation_and_its_noise_floor(golden_dir):
    """``store_dtype=torch.bfloat16`` (the storage precision of the B200 production path, emulated on the CPU):
    (1) with ``store_dtype=None`` the oracle still reproduces the reference fixture exactly;
    (2) the bf16-storage logits stay within the 2e-2 bf16 tolerance of the reference's fp32 logits;
    (3) the noise floor the GPU parity tests are calibrated on: the SAME bf16-storage oracle with its convolutions
        accumulated in fp64 instead of fp32 (a different but equally valid summation order) already differs by more than
        1e-3 in the logits and by more than 5e-2 in the deepest gradient tensors, while the tensors next to the loss agree
        to 1e-3 -- i.e. whole-network per-tensor gradient agreement at 5e-2 is not a property ANY two bf16-storage
        implementations have at random initialisation; tests/test_gpu_parity_strict.py therefore demands 5e-2 / cosine
        0.999 per residual block (teacher-forced) and "3 x this floor" for the whole network."""
    import torch.nn.functional as F

    g = _load(golden_dir, "unet_b1.npz")
    shape, seed = tuple(int(v) for v in g["shape"]), int(g["seed"])
    x = O.synth_patch(shape, 1000 + seed, "ct")
    lab = O.synth_labels((shape[0],) + shape[2:], 2000 + seed, 16, 32)
    w16 = g["w16"].tolist()
    cm = O.remap_unsupervised(lab, w16).squeeze(1)

    def run(store, fp64_convs=False):
        sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(32, 16, seed).items()}
        orig = F.conv3d
        if fp64_convs:
            F.conv3d = lambda i, w, b=None, *a, **k: orig(i.double(), w.double(), None if b is None else b.double(), *a,
                                                         **k).to(i.dtype)
        try:
            lg = O.unet3d_forward(sd, x, store_dtype=store)
            O.partial_label_loss(lg, cm, w16).backward()
        finally:
            F.conv3d = orig
        return lg.detach(), {k: v.grad for k, v in sd.items()}

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    ref = torch.from_numpy(g["logits"])
    l32, _ = run(None)
    assert torch.equal(l32, ref)
    lb, gb = run(torch.bfloat16)
    assert 1e-3 < rel(lb, ref) < 2e-2
    lt, gt = run(torch.bfloat16, fp64_convs=True)
    assert rel(lt, lb) > 1e-3
    floors = {k: rel(gt[k], gb[k]) for k in gb}
    assert max(floors.values()) > 5e-2, max(floors.values())
    assert floors["precls_conv.2.weight"] < 1e-3 and floors["precls_conv.2.bias"] < 1e-3
