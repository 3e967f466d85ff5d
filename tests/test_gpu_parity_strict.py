"""Network- and block-level GRADIENT parity of the CUDA path against the CPU oracle, with stated tolerances.

Why two granularities.  A randomly initialised 31-layer ReLU network is chaotic in its stored precision: two *correct*
bf16-storage implementations that differ only in the order of their fp32 summations already drift apart, because a
summation-order difference of 1e-6 relative occasionally lands a value on the other side of a bf16 rounding boundary,
that 1-ulp difference perturbs 27*Cout outputs of the next convolution, and after ~4 layers every stored element is
rounded independently (``tests/test_oracle_golden.py::test_bf16_storage_noise_floor`` measures it on the CPU: logits
8e-3, deepest gradients 0.2-0.3 rel-L2 between the oracle and the same oracle with fp64-accumulated convolutions).
Therefore:

  * BLOCK level (teacher forcing: identical stored inputs and identical upstream gradients on both sides, one residual
    block deep) the tolerance of SURVEY.md 8(c) is demanded literally:
      bf16 path  output rel-L2 <= 1e-2, every gradient tensor rel-L2 <= 5e-2 and cosine >= 0.999
      fp32 path  output rel-L2 <= 1e-5, every gradient tensor rel-L2 <= 1e-4
    against ``oracle.no_bottleneck(store_dtype=...)``.  Every block flavour of the network and every tcgen05 kernel
    mode (fused GroupNorm-backward epilogue included) is on this list.
  * NETWORK level (all 107 gradient tensors) the CUDA path must be as close to the storage-emulating oracle as that
    oracle is to its own re-ordered self: per tensor  err <= 3 * floor + 2e-2  (bf16), and the median over tensors of
    err / floor must stay below 2 (measured on B200: 0.9-1.6).
    fp32 path: floor = fp32 oracle vs fp64 oracle (~6e-5: pure rounding, no ReLU gate resolves differently).  Any two
    independent fp32 implementations differ by ~2e-7 in a pre-activation, which over the 1.6e7 gated activations of this
    network flips about ONE gate, and one flipped gate moves every upstream gradient tensor by ~2e-3 rel-L2 (measured:
    the CUDA path shows exactly that signature -- 80 tensors at 1.5e-3..2.3e-3, the rest at the floor).  The network-level
    fp32 bound is therefore  err <= 3 * floor + 5e-3  (two gates' worth); the literal 1e-4 is demanded where no gate can
    hide a kernel error: per block, above.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def _reset():
    import multimodal_pl_b200 as mm

    mm.set_conv_algo("auto")
    mm.set_compute_dtype(torch.bfloat16)


def _rand(shape, seed, scale=1.0):
    return scale * torch.randn(shape, generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------------------------------------------------------
# block level
# ------------------------------------------------------------------------------------------------------------------
BLOCKS = [
    # (inplanes, planes, stride, spatial)         kernels exercised on the bf16 path
    (32, 32, 1, (6, 20, 12)),     # conv_tc<32,32,4,S1K3,WRES> fprop / fprop+res / dgrad+GN, identity residual via alias
    (32, 64, 2, (8, 20, 12)),     # parity split, S2F, S2D dgrad+GN (8 classes), 1x1x1 stride-2 fprop/dgrad, dual GN
    (64, 64, 1, (5, 18, 9)),      # streamed-weight 64-channel kernel
    (64, 128, 2, (6, 12, 10)),
    (128, 128, 1, (3, 9, 8)),
    (256, 256, 1, (2, 5, 7)),     # small-problem tiling (NA = 1, single-plane tiles)
    (256, 128, 1, (4, 6, 8)),     # decoder block: width-halving 3x3x3 + 1x1x1 stride-1 downsample, dual GN
    (64, 32, 1, (6, 16, 16)),     # decoder block feeding the full-resolution stage
]


def _block_state(cin, cout, stride, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {"gn1.weight": 1 + 0.2 * torch.randn(cin, generator=g), "gn1.bias": 0.2 * torch.randn(cin, generator=g),
          "conv1.weight": torch.randn((cout, cin, 3, 3, 3), generator=g) / (27 * cin) ** 0.5,
          "gn2.weight": 1 + 0.2 * torch.randn(cout, generator=g), "gn2.bias": 0.2 * torch.randn(cout, generator=g),
          "conv2.weight": torch.randn((cout, cout, 3, 3, 3), generator=g) / (27 * cout) ** 0.5}
    if stride != 1 or cin != cout:
        sd["downsample.0.weight"] = 1 + 0.2 * torch.randn(cin, generator=g)
        sd["downsample.0.bias"] = 0.2 * torch.randn(cin, generator=g)
        sd["downsample.2.weight"] = torch.randn((cout, cin, 1, 1, 1), generator=g) / cin ** 0.5
    return sd


def _run_block(cin, cout, stride, sp, dtype, algo):
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.unet3D import GNReLUConv, NoBottleneck, conv3x3x3

    st = torch.bfloat16 if dtype == torch.bfloat16 else None
    sd = _block_state(cin, cout, stride, 17 * cin + cout + stride)
    x = _rand((2, cin) + sp, 3) + 0.25
    if st is not None:
        x = x.bfloat16().float()                      # a stored tensor: identical on both sides
    # ---- oracle
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    yr = O.no_bottleneck(xr, sdr, "", stride, store_dtype=st)
    dy = _rand(tuple(yr.shape), 4)
    if st is not None:
        dy = dy.bfloat16().float()
    yr.backward(dy)
    # ---- device: the product's own module, built the way unet3D_baseline._make_layer builds it
    mm.set_compute_dtype(dtype)
    mm.set_conv_algo(algo)
    try:
        ds = None
        if stride != 1 or cin != cout:       # as unet3D_baseline._make_layer builds it (reference unet3D.py:643-649)
            ds = GNReLUConv(torch.nn.GroupNorm(16, cin), torch.nn.ReLU(inplace=True),
                            conv3x3x3(cin, cout, kernel_size=(1, 1, 1), stride=(stride,) * 3, padding=0, weight_std=True))
        blk = NoBottleneck(cin, cout, (stride,) * 3, downsample=ds, weight_std=True).cuda()
        blk.load_state_dict(sd)
        mm.ops.begin_forward(torch.device("cuda"))
        xd = x.cuda().requires_grad_(True)
        launches = mm._lib.launch_count()
        y = blk(xd)
        y.backward(dy.cuda().to(y.dtype))
        launches = mm._lib.launch_count() - launches
        grads = {k: p.grad for k, p in blk.named_parameters()}
        return (y.float(), yr), (xd.grad.float(), xr.grad), {k: (grads[k], sdr[k].grad) for k in sd}, launches
    finally:
        _reset()


@pytest.mark.parametrize("cin,cout,stride,sp", BLOCKS)
def test_block_bf16_tcgen05_vs_bf16_storage_oracle(cin, cout, stride, sp):
    (y, yr), (dx, dxr), grads, _ = _run_block(cin, cout, stride, sp, torch.bfloat16, "tcgen05")
    assert rel(y, yr) < 1e-2, rel(y, yr)
    assert rel(dx, dxr) < 5e-2 and cosine(dx, dxr) > 0.999, ("dx", rel(dx, dxr), cosine(dx, dxr))
    for k, (g, gr) in grads.items():
        assert rel(g, gr) < 5e-2 and cosine(g, gr) > 0.999, (k, rel(g, gr), cosine(g, gr))


@pytest.mark.parametrize("cin,cout,stride,sp", BLOCKS[:4] + BLOCKS[6:])
def test_block_fp32_exact_path_vs_oracle(cin, cout, stride, sp):
    (y, yr), (dx, dxr), grads, _ = _run_block(cin, cout, stride, sp, torch.float32, "direct")
    assert rel(y, yr) < 1e-5, rel(y, yr)
    assert rel(dx, dxr) < 1e-4, ("dx", rel(dx, dxr))
    for k, (g, gr) in grads.items():
        assert rel(g, gr) < 1e-4, (k, rel(g, gr))


# ------------------------------------------------------------------------------------------------------------------
# network level
# ------------------------------------------------------------------------------------------------------------------
def _oracle_net(sd, x, target, w16, base, store_dtype, reorder=None, dtype=torch.float32):
    """Oracle forward + loss + backward.  ``reorder`` builds the twin used to measure the noise floor -- the same
    arithmetic with a different, equally valid fp32 summation order inside the convolutions: "fp64" accumulates them in
    fp64, "flip" convolves the spatially flipped input with the spatially flipped filter (reversed tap order)."""
    sdr = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    orig = F.conv3d
    if reorder == "fp64":
        def conv64(inp, w, b=None, *a, **k):
            return orig(inp.double(), w.double(), None if b is None else b.double(), *a, **k).to(inp.dtype)
        F.conv3d = conv64
    elif reorder == "flip":
        def convflip(inp, w, b=None, stride=1, padding=0, *a, **k):
            s_ = stride if isinstance(stride, int) else stride[0]
            if s_ != 1:            # a strided window is not flip-symmetric for even extents: leave those four layers
                return orig(inp, w, b, stride, padding, *a, **k)
            return torch.flip(orig(torch.flip(inp, (2, 3, 4)), torch.flip(w, (2, 3, 4)), b, stride, padding, *a, **k),
                              (2, 3, 4))
        F.conv3d = convflip
    try:
        logits = O.unet3d_forward(sdr, x.to(dtype), base, store_dtype=store_dtype)
        loss = O.partial_label_loss(logits, target, w16)
        loss.backward()
    finally:
        F.conv3d = orig
    return logits.detach(), loss.item(), {k: v.grad for k, v in sdr.items()}


def _device_net(sd, x, target, w16, base, dtype, algo):
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(dtype)
    mm.set_conv_algo(algo)
    try:
        model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True, base=base).cuda()
        model.load_state_dict(sd)
        model.train()
        logits = model(x.cuda())[0]
        loss = EDiceLoss_partial(16)(logits, target.cuda(), mask=[torch.tensor(w16)] * x.shape[0], soft_max=True)
        loss.backward()
        torch.cuda.synchronize()
        return logits.detach().cpu(), loss.item(), {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    finally:
        _reset()


def _inputs(shape, seed, base=32):
    sd = O.synth_state_dict(base, 16, seed)
    x = O.synth_patch(shape, 1000 + seed, "ct" if seed % 2 == 0 else "mri")
    lo = tuple(max(4, s // 4) for s in shape[2:])
    lab = F.interpolate(O.synth_labels((shape[0],) + lo, 2000 + seed, 16, 32), size=shape[2:], mode="nearest")
    w16 = [1.0, 0, 0, 0, 1.0, 0, 0, 1.0] + [0.0] * 8
    return sd, x, O.remap_unsupervised(lab, w16).squeeze(1), w16


def _check_against_floor(tag, dev, ref, twin, slack, ratio_cap=3.0, median_cap=2.0):
    """dev / ref / twin = (logits, loss, grads).  floor_k = rel(twin_k, ref_k); demand err_k <= ratio_cap*floor_k + slack."""
    lf = rel(twin[0], ref[0])
    le = rel(dev[0], ref[0])
    assert le <= ratio_cap * lf + slack, (tag, "logits", le, lf)
    ratios, bad = [], []
    for k, gr in ref[2].items():
        floor, err = rel(twin[2][k], gr), rel(dev[2][k], gr)
        ratios.append(err / max(floor, 1e-12))
        if err > ratio_cap * floor + slack:
            bad.append((k, err, floor))
    assert not bad, (tag, bad[:6], len(bad))
    med = float(np.median(ratios))
    assert med < median_cap, (tag, "median err/floor", med)
    return le, lf, med


@pytest.mark.parametrize("shape,seed", [((1, 1, 16, 32, 32), 0), ((2, 1, 16, 32, 32), 1)])
@pytest.mark.parametrize("algo", ["auto", "direct"])
def test_unet_bf16_all_gradients_vs_bf16_storage_oracle(shape, seed, algo):
    sd, x, tgt, w16 = _inputs(shape, seed)
    ref = _oracle_net(sd, x, tgt, w16, 32, torch.bfloat16)
    twin = _oracle_net(sd, x, tgt, w16, 32, torch.bfloat16, reorder="fp64")
    dev = _device_net(sd, x, tgt, w16, 32, torch.bfloat16, algo)
    assert abs(dev[1] - ref[1]) <= 3 * abs(twin[1] - ref[1]) + 1e-3 * abs(ref[1])
    _check_against_floor(f"bf16/{algo}", dev, ref, twin, slack=2e-2)
    # next to the loss nothing is chaotic yet: the classifier gradients meet the literal 8(c) tolerance
    for k in ("precls_conv.2.weight", "precls_conv.2.bias", "precls_conv.0.weight", "precls_conv.0.bias"):
        assert rel(dev[2][k], ref[2][k]) < 5e-2 and cosine(dev[2][k], ref[2][k]) > 0.999, k


@pytest.mark.parametrize("shape,seed", [((1, 1, 16, 32, 32), 0), ((2, 1, 16, 32, 32), 1)])
def test_unet_fp32_all_gradients_vs_fp64_oracle(shape, seed):
    """fp32 exact path against an fp64 run of the oracle: logits 1e-5, loss 1e-6, every gradient tensor within
    3 x (fp32 oracle vs fp64 oracle) + 5e-3 -- see the module docstring for why one or two flipped ReLU gates have to be
    allowed at network level -- and the classifier gradients, which no gate separates from the loss, within 1e-4."""
    sd, x, tgt, w16 = _inputs(shape, seed)
    ref = _oracle_net(sd, x, tgt, w16, 32, None, dtype=torch.float64)
    twin = _oracle_net(sd, x, tgt, w16, 32, None, dtype=torch.float32)
    dev = _device_net(sd, x, tgt, w16, 32, torch.float32, "direct")
    assert rel(dev[0], ref[0]) < 1e-5
    assert abs(dev[1] - ref[1]) < 1e-6 * max(1.0, abs(ref[1])) + 3 * abs(twin[1] - ref[1])
    _check_against_floor("fp32", dev, ref, twin, slack=5e-3, median_cap=100.0)
    # tensors downstream of every gate (next to the loss) see no flip: the literal tolerance holds there
    for k in ("precls_conv.2.weight", "precls_conv.2.bias"):
        assert rel(dev[2][k], ref[2][k]) < 1e-4, (k, rel(dev[2][k], ref[2][k]))


def test_wide_backbone_base64_backward_vs_oracle():
    """BASELINE configs[4] widths (64..512) at 32x64x64, forward AND backward: the 128->128 / 512->512 instantiations and
    the nout > 256 column tiling exist only at these widths."""
    shape = (1, 1, 32, 64, 64)
    sd, x, tgt, w16 = _inputs(shape, 5, base=64)
    ref = _oracle_net(sd, x, tgt, w16, 64, torch.bfloat16)
    twin = _oracle_net(sd, x, tgt, w16, 64, torch.bfloat16, reorder="fp64")
    dev = _device_net(sd, x, tgt, w16, 64, torch.bfloat16, "auto")
    assert rel(dev[0], ref[0]) < 2e-2
    _check_against_floor("bf16/base64", dev, ref, twin, slack=2e-2)
    ref32 = _oracle_net(sd, x, tgt, w16, 64, None)
    dev32 = _device_net(sd, x, tgt, w16, 64, torch.float32, "direct")
    assert rel(dev32[0], ref32[0]) < 1e-5
    for k, gr in ref32[2].items():      # 4x the gated activations of the base-32 test: a handful of flipped gates
        assert rel(dev32[2][k], gr) < 2e-2, (k, rel(dev32[2][k], gr))


# ------------------------------------------------------------------------------------------------------------------
# full extent of the benchmarked configuration
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.slow
def test_cfg2_full_extent_forward_loss_backward_vs_oracle():
    """BASELINE configs[1] as benchmarked: batch 2, 1x64x192x192, 16 classes, bf16 tcgen05 path, forward + partial-label
    loss + backward at the FULL extent against the bf16-storage CPU oracle: logits rel-L2 <= 2e-2, loss rel <= 1e-2, and
    all 107 gradient tensors inside the measured noise floor (see the module docstring).  Every size-dependent kernel
    variant (62 items per CTA, stride-2 TD = 4 tiles, 64-bit offsets in the GroupNorm kernels) is exercised here."""
    shape = (2, 1, 64, 192, 192)
    sd, x, tgt, w16 = _inputs(shape, 2)
    ref = _oracle_net(sd, x, tgt, w16, 32, torch.bfloat16)
    dev = _device_net(sd, x, tgt, w16, 32, torch.bfloat16, "auto")
    assert rel(dev[0], ref[0]) < 2e-2, rel(dev[0], ref[0])
    assert abs(dev[1] - ref[1]) < 1e-2 * abs(ref[1]), (dev[1], ref[1])
    twin = _oracle_net(sd, x, tgt, w16, 32, torch.bfloat16, reorder="flip")
    le, lf, med = _check_against_floor("cfg2/bf16", dev, ref, twin, slack=2e-2)
    print(f"cfg2 full extent: logits err {le:.3e} (floor {lf:.3e}), median grad err/floor {med:.2f}")
    # argmax of the logits (what inference reports) agrees wherever the oracle's decision is not a bf16-level near-tie
    top2 = ref[0].topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 5e-2 * top2[:, 0].abs().clamp_min(1.0)
    assert (dev[0].argmax(1) != ref[0].argmax(1))[decided].float().mean().item() < 1e-3
