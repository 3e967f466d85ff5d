"""The refiner unet3D_g and the discriminator norm_style_discriminator_output on the device kernels vs
tests/golden/aux_nets.npz, written by oracle/make_golden_aux.py from the UNMODIFIED reference modules
(unet3D.py:1507-1623, :1907-1947).  fp32 exact path: outputs 1e-5, gradients 1e-3 (a ReLU gate within fp32 rounding of
zero may resolve differently; one flip moves upstream gradients by ~1e-3); bf16 tensor-core path (zero-padded widths,
space-to-depth rewrite of the 4x4x4 stride-2 convolutions): outputs 3e-2, gradient direction cosine >= 0.95."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


def _gen(golden_dir):
    spec = importlib.util.spec_from_file_location(
        "make_golden_aux", os.path.join(os.path.dirname(golden_dir), "..", "oracle", "make_golden_aux.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)                       # seeded inputs / constants only (main() is not run)
    return gen


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), torch.as_tensor(b).double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), torch.as_tensor(b).double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def _run(model, x, seed, gen):
    model.cuda().train()
    y = model(x.cuda())
    dy = gen.seeded(tuple(y.shape), seed) / float(np.prod(y.shape)) ** 0.5
    (y.float() * dy.cuda()).sum().backward()
    return y, dict(model.named_parameters())


@pytest.mark.parametrize("dtype,algo", [(torch.float32, "direct"), (torch.bfloat16, "auto")])
def test_refiner_unet3d_g_matches_reference_fixture(golden_dir, dtype, algo):
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.aux_nets import unet3D_g

    gen = _gen(golden_dir)
    g = np.load(os.path.join(golden_dir, "aux_nets.npz"))
    mm.set_compute_dtype(dtype)
    mm.set_conv_algo(algo)
    try:
        model = unet3D_g(**gen.REFINER)
        sd = O.synth_named_state({k: tuple(v.shape) for k, v in model.state_dict().items()}, gen.SEED)
        model.load_state_dict(sd)
        y, params = _run(model, gen.refiner_input(), 41, gen)
        assert tuple(y.shape) == gen.REFINER_IN[:1] + (2,) + gen.REFINER_IN[2:] and y.dtype == torch.float32
        exact = dtype == torch.float32
        assert rel(y, g["refiner/out"]) < (1e-5 if exact else 3e-2), rel(y, g["refiner/out"])
        for k in gen.REF_FULL:
            ref = g["refiner/grad:" + k]
            assert tuple(params[k].grad.shape) == ref.shape, k
            if exact:
                assert rel(params[k].grad, ref) < 1e-3, (k, rel(params[k].grad, ref))
            else:
                assert cosine(params[k].grad, ref) > 0.95, (k, cosine(params[k].grad, ref))
        bad = []
        for k, p in params.items():
            n, r = p.grad.double().norm().item(), float(g["refiner/norm:" + k][0])
            if abs(n - r) > (5e-3 if exact else 0.25) * max(r, 1e-9) + 1e-8:
                bad.append((k, n, r))
        assert not bad, bad[:5]
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("dtype,algo", [(torch.float32, "direct"), (torch.bfloat16, "auto")])
def test_discriminator_matches_reference_fixture(golden_dir, dtype, algo):
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.aux_nets import norm_style_discriminator_output

    gen = _gen(golden_dir)
    g = np.load(os.path.join(golden_dir, "aux_nets.npz"))
    mm.set_compute_dtype(dtype)
    mm.set_conv_algo(algo)
    try:
        model = norm_style_discriminator_output(num_classes=2)
        sd = O.synth_named_state({k: tuple(v.shape) for k, v in model.state_dict().items()}, gen.SEED + 1)
        model.load_state_dict(sd)
        y, params = _run(model, gen.disc_input(), 42, gen)
        exact = dtype == torch.float32
        assert tuple(y.shape) == (gen.DISC_IN[0], 2)
        assert rel(y, g["disc/out"]) < (1e-5 if exact else 3e-2), rel(y, g["disc/out"])
        for k in gen.DISC_FULL:
            ref = g["disc/grad:" + k]
            if exact:
                assert rel(params[k].grad, ref) < 1e-3, (k, rel(params[k].grad, ref))
            else:
                assert cosine(params[k].grad, ref) > 0.95, (k, cosine(params[k].grad, ref))
        for k, p in params.items():
            n, r = p.grad.double().norm().item(), float(g["disc/norm:" + k][0])
            assert abs(n - r) <= (5e-3 if exact else 0.25) * max(r, 1e-9) + 1e-8, (k, n, r)
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


def test_aux_kernels_space_to_depth_bias_lrelu_upsample():
    """The three helper kernels against their tensor-op definitions, forward and backward."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200 import ops

    mm.set_compute_dtype(torch.float32)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn((2, 3, 4, 6, 8), generator=g)
        xd = x.cuda().requires_grad_(True)
        y = ops.space_to_depth2(xd, 32)
        n, c, d, h, w = x.shape
        ref = x.view(n, c, d // 2, 2, h // 2, 2, w // 2, 2).permute(0, 3, 5, 7, 1, 2, 4, 6).reshape(n, 8 * c, d // 2, h // 2, w // 2)
        assert torch.equal(y[:, :24].cpu(), ref) and y[:, 24:].abs().max().item() == 0
        dy = torch.randn(tuple(y.shape), generator=g)
        y.backward(dy.cuda())
        dref = dy[:, :24].view(n, 2, 2, 2, c, d // 2, h // 2, w // 2).permute(0, 4, 5, 1, 6, 2, 7, 3).reshape(n, c, d, h, w)
        assert torch.equal(xd.grad.cpu(), dref)
        # bias + LeakyReLU
        a = torch.randn((2, 32, 3, 4, 5), generator=g)
        b = torch.randn(32, generator=g)
        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = torch.nn.functional.leaky_relu(ar + br.view(1, -1, 1, 1, 1), 0.2)
        da = torch.randn(tuple(a.shape), generator=g)
        yr.backward(da)
        ad, bd = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        yd = ops.bias_leaky_relu(ad, bd, 0.2)
        yd.backward(da.cuda())
        assert rel(yd, yr) < 1e-6 and rel(ad.grad, ar.grad) < 1e-6 and rel(bd.grad, br.grad) < 1e-5
        # trilinear x2 of an fp32 NCDHW tensor
        z = torch.randn((2, 2, 3, 5, 4), generator=g)
        zr = z.clone().requires_grad_(True)
        ur = torch.nn.functional.interpolate(zr, scale_factor=2, mode="trilinear")
        du = torch.randn(tuple(ur.shape), generator=g)
        ur.backward(du)
        zd = z.cuda().requires_grad_(True)
        ud = ops.upsample2x_ncdhw(zd)
        ud.backward(du.cuda())
        assert rel(ud, ur) < 1e-6 and rel(zd.grad, zr.grad) < 1e-6
    finally:
        mm.set_compute_dtype(torch.bfloat16)
