"""Device-side input pipeline (multimodal_pl_b200.input_pipeline, csrc/input.cu) vs the numpy restatement of the
reference's dataset code (oracle.prepare_patch_ref / augment_ref; MOTSDataset.py:33-52, :171-186, :269-297, :299-395)."""
import numpy as np
import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,crop,origin", [
    ((40, 70, 33), (32, 48, 16), (3, 11, 7)),        # volume larger than the crop on every axis
    ((20, 70, 12), (32, 48, 16), (2, 0, 4)),         # h and d shorter than crop + 5: zero padding enters the patch
    ((37, 53, 21), (32, 48, 16), (5, 5, 5)),         # crop reaches exactly the padded border (crop + 5)
])
@pytest.mark.parametrize("modality,src", [("ct", "int16"), ("ct", "float32"), ("mri", "int16"), ("mri", "float32")])
def test_prepare_patch_matches_reference_dataset_code(shape, crop, origin, modality, src):
    from multimodal_pl_b200.input_pipeline import padded_shape, prepare_patch

    rng = np.random.RandomState(7)
    if src == "int16":
        image = rng.randint(-1000, 1500, size=shape).astype(np.int16)
    else:
        image = (rng.randn(*shape) * 400).astype(np.float32)
    label = rng.randint(0, 16, size=shape).astype(np.uint8)
    atlas = rng.rand(13, 9, 11, 6).astype(np.float32)
    ps = padded_shape(shape, crop)
    assert all(o + c <= p for o, c, p in zip(origin, crop, ps))
    ri, rl, ra = O.prepare_patch_ref(image, label, atlas, modality == "ct", crop, origin)
    gi, gl, ga = prepare_patch(torch.from_numpy(image).cuda(), torch.from_numpy(label).cuda(), crop, origin, modality,
                               torch.from_numpy(atlas).cuda())
    assert tuple(gi.shape) == ri.shape == (1, crop[2], crop[0], crop[1])
    if src == "int16" and modality == "ct":
        assert np.array_equal(gi.cpu().numpy(), ri)                   # int16 / 325.0 in fp64, rounded once: bit-exact
    else:
        # fp32 sources: numpy's fp32 mean / std use pairwise fp32 sums; MRI int16: fp64 moments summed in another order
        assert np.allclose(gi.cpu().numpy(), ri, rtol=2e-5, atol=2e-6)
    assert np.array_equal(gl.cpu().numpy(), rl)
    assert np.array_equal(ga.cpu().numpy(), ra)
    # uint8 labels for the loss kernel
    _, gl8, _ = prepare_patch(torch.from_numpy(image).cuda(), torch.from_numpy(label).cuda(), crop, origin, modality,
                              label_dtype=torch.uint8)
    assert gl8.dtype == torch.uint8 and np.array_equal(gl8.cpu().numpy().astype(np.float32), rl)


def test_augmentations_match_published_algorithms():
    from multimodal_pl_b200.input_pipeline import augment_patch

    rng = np.random.RandomState(3)
    img = (rng.randn(1, 12, 20, 28) * 0.4).astype(np.float32)
    # brightness (multiplicative, additive) + contrast with preserve_range
    p = {"mult": 1.2, "add": -0.07, "contrast": 1.21}
    got = augment_patch(torch.from_numpy(img).cuda(), p).cpu().numpy()
    assert np.allclose(got, O.augment_ref(img, p), rtol=1e-5, atol=1e-6)
    p = {"contrast": 0.8}
    got = augment_patch(torch.from_numpy(img).cuda(), p).cpu().numpy()
    assert np.allclose(got, O.augment_ref(img, p), rtol=1e-5, atol=1e-6)
    # Gaussian blur == scipy.ndimage.gaussian_filter (reflect boundary, truncate 4)
    for sigma in (0.5, 0.83, 1.0):
        got = augment_patch(torch.from_numpy(img).cuda(), {"blur_sigma": sigma}).cpu().numpy()
        assert np.allclose(got, O.augment_ref(img, {"blur_sigma": sigma}), rtol=1e-4, atol=1e-5), sigma
    # Gaussian noise: counter-based generator -- reproducible from the seed, zero mean, the requested std, white
    big = torch.zeros((1, 32, 64, 64), device="cuda")
    a = augment_patch(big.clone(), {"noise_std": 0.05, "seed": 11})
    b = augment_patch(big.clone(), {"noise_std": 0.05, "seed": 11})
    c = augment_patch(big.clone(), {"noise_std": 0.05, "seed": 12})
    assert torch.equal(a, b) and not torch.equal(a, c)
    v = a.flatten().double()
    assert abs(v.mean().item()) < 5e-4 and abs(v.std().item() - 0.05) < 5e-4
    assert abs((v[1:] * v[:-1]).mean().item()) < 2e-5                                   # lag-1 autocorrelation
    k = ((v / 0.05) ** 4).mean().item()
    assert abs(k - 3.0) < 0.1                                                           # Gaussian kurtosis


def test_pipeline_object_feeds_the_train_step():
    """PatchPipeline output plugs into the model + loss unchanged (shapes, dtypes, uint8 labels)."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.input_pipeline import PatchPipeline
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.bfloat16)
    rng = np.random.RandomState(0)
    vol = rng.randint(-1000, 1500, size=(48, 56, 30)).astype(np.int16)
    lab = rng.randint(0, 16, size=(48, 56, 30)).astype(np.uint8)
    pipe = PatchPipeline((32, 32, 16), atlas=torch.rand(15, 8, 8, 8), seed=1, label_dtype=torch.uint8)
    img, l8, cat = pipe(vol, lab, "ct")
    assert tuple(img.shape) == (1, 16, 32, 32) and l8.dtype == torch.uint8 and tuple(cat.shape) == (15, 16, 32, 32)
    assert img.abs().max().item() <= 1.0 + 0.6          # CT window +-1 before the (rare) augmentations
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda().train()
    logits = model(img.unsqueeze(0))[0]
    loss = EDiceLoss_partial(16)(logits, l8, mask=[torch.ones(16)])
    loss.backward()
    assert torch.isfinite(loss)
