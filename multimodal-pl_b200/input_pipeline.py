"""Device-side input pipeline for the train loop.

What the reference does per sample on the CPU, in numpy, inside a DataLoader worker
(``AMOSDataSet_newatlas.__getitem__``, MOTSDataset.py:299-395): nearest-neighbour resize of the organ atlas to the image
shape (:357), zero-padding of image / label / atlas up to crop + 5 (:370-372), intensity scaling -- CT clipped to +-325 HU and
divided by 325, MRI z-scored over the padded volume (:374, ``truncate`` :171-186) --, a random crop (:377-383), the
transpose to [1, D, H, W] (:389-391) and, through batchgenerators, the intensity augmentations of ``get_train_transform``
(:33-52: Gaussian noise, Gaussian blur, multiplicative / additive brightness, contrast).  Here the raw volume is uploaded
once (int16 / fp32 as stored) and every step of that list is a kernel of libmmpl_b200.so (csrc/input.cu); the random draws
(crop origin, which augmentations fire, their parameters) stay on the host and are passed in, so a run is reproducible
from a ``numpy.random.RandomState`` exactly like the reference's.

Volume layout follows the reference: ``image[h, w, d]`` (d fastest); outputs are ``[1, D, H, W]``.
"""
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import p as _p

_SRC_CODE = {torch.float32: 0, torch.int16: 1, torch.uint8: 2}


def _src(t: torch.Tensor):
    if t.dtype not in _SRC_CODE:
        t = t.float()
    return t.contiguous(), _SRC_CODE[t.dtype]


def padded_shape(shape, crop):
    """pad_image (:269-282): every axis grows to at least crop + 5."""
    return tuple(max(int(s), int(c) + 5) for s, c in zip(shape, crop))


def draw_crop_origin(shape, crop, rng: np.random.RandomState):
    """The reference's random crop (:378-380): randint(padded - crop) per axis, in the order h, w, d."""
    ps = padded_shape(shape, crop)
    return tuple(int(rng.randint(ps[i] - crop[i])) for i in range(3))


def prepare_patch(image: torch.Tensor, label: Optional[torch.Tensor], crop: Sequence[int], origin: Sequence[int],
                  modality: str = "ct", atlas: Optional[torch.Tensor] = None, label_dtype=torch.float32):
    """image / label [h, w, d] on the device (int16 / uint8 / fp32 as stored), ``crop`` = (crop_h, crop_w, crop_d),
    ``origin`` = (b, c, a).  -> (image [1, D, H, W] fp32, label [1, D, H, W] or None, catlas [K, D, H, W] or None) with
    D = crop_d, H = crop_h, W = crop_w."""
    _lib.require_device()
    L = _lib.lib()
    assert modality in ("ct", "mri") and image.is_cuda and image.dim() == 3
    dev = image.device
    st = _lib.stream_ptr()
    h, w, d = (int(v) for v in image.shape)
    if label is not None and tuple(label.shape) != (h, w, d):          # :360-368: common leading extents
        h, w, d = (min(a, b) for a, b in zip(image.shape, label.shape))
        image, label = image[:h, :w, :d], label[:h, :w, :d]
    ch, cw, cd = (int(v) for v in crop)
    b, c, a = (int(v) for v in origin)
    img, code = _src(image)
    out = torch.empty((1, cd, ch, cw), dtype=torch.float32, device=dev)
    moments = None
    mode = 0
    if modality == "mri":
        mode = 1
        moments = torch.empty(2, dtype=torch.float64, device=dev)
        _lib.check(L.mmpl_volume_moments(_p(img), code, img.numel(), _p(moments), st), "volume_moments")
    _lib.check(L.mmpl_prepare_patch(_p(img), code, _p(out), 0, h, w, d, b, c, a, ch, cw, cd, mode, _p(moments), st),
               "prepare_patch")
    lab_out = None
    if label is not None:
        lab, lcode = _src(label)
        u8 = label_dtype == torch.uint8
        lab_out = torch.empty((1, cd, ch, cw), dtype=torch.uint8 if u8 else torch.float32, device=dev)
        _lib.check(L.mmpl_prepare_patch(_p(lab), lcode, _p(lab_out), int(u8), h, w, d, b, c, a, ch, cw, cd, 2, None, st),
                   "prepare_patch(label)")
    cat = None
    if atlas is not None:
        at = atlas.to(dev, torch.float32).contiguous()
        k, ha, wa, da = (int(v) for v in at.shape)
        cat = torch.empty((k, cd, ch, cw), dtype=torch.float32, device=dev)
        _lib.check(L.mmpl_atlas_patch(_p(at), _p(cat), k, ha, wa, da, h, w, d, b, c, a, ch, cw, cd, st), "atlas_patch")
    return out, lab_out, cat


def draw_augment_params(rng: np.random.RandomState):
    """One sample's draws for get_train_transform (:33-52), in its order; each entry None = the transform does not fire."""
    p = {}
    p["noise_std"] = float(rng.uniform(0, 0.1)) if rng.uniform() < 0.1 else None             # GaussianNoiseTransform
    p["blur_sigma"] = None                                                                   # GaussianBlurTransform
    if rng.uniform() < 0.2 and rng.uniform() <= 0.5:
        p["blur_sigma"] = float(rng.uniform(0.5, 1.0))
    p["mult"] = float(rng.uniform(0.75, 1.25)) if rng.uniform() < 0.15 else None              # BrightnessMultiplicative
    p["add"] = None                                                                          # BrightnessTransform
    if rng.uniform() < 0.15 and rng.uniform() <= 0.5:
        p["add"] = float(rng.normal(0.0, 0.1))
    p["contrast"] = None                                                                     # ContrastAugmentation
    if rng.uniform() < 0.15:
        p["contrast"] = float(rng.uniform(0.75, 1.0) if rng.uniform() < 0.5 else rng.uniform(1.0, 1.25))
    p["seed"] = int(rng.randint(0, 2 ** 31 - 1))
    return p


def gaussian_taps(sigma: float):
    """scipy.ndimage.gaussian_filter's kernel (order 0, truncate 4.0): radius = int(4 sigma + 0.5), normalised."""
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return (w / w.sum()).astype(np.float32), radius


def augment_patch(image: torch.Tensor, params: dict) -> torch.Tensor:
    """In-place intensity augmentation of one patch [1, D, H, W] fp32 with the draws of ``draw_augment_params``."""
    _lib.require_device()
    L = _lib.lib()
    assert image.is_cuda and image.dtype == torch.float32 and image.is_contiguous()
    st = _lib.stream_ptr()
    n = image.numel()
    noise = params.get("noise_std") or 0.0
    if noise > 0.0:
        _lib.check(L.mmpl_augment_patch(_p(image), n, float(noise), int(params.get("seed", 0)), 1.0, 0.0, 1.0, None, st),
                   "augment(noise)")
    if params.get("blur_sigma"):
        taps, radius = gaussian_taps(params["blur_sigma"])
        taps_d = torch.from_numpy(taps).to(image.device)
        d, h, w = image.shape[-3:]
        tmp = torch.empty_like(image)
        src, dst = image, tmp
        for axis in range(3):
            _lib.check(L.mmpl_blur_axis(_p(src), _p(dst), d, h, w, axis, _p(taps_d), radius, st), "blur_axis")
            src, dst = dst, src
        if src is not image:
            image.copy_(src)
    mult = params.get("mult") if params.get("mult") is not None else 1.0
    add = params.get("add") if params.get("add") is not None else 0.0
    if mult != 1.0 or add != 0.0:
        _lib.check(L.mmpl_augment_patch(_p(image), n, 0.0, 0, float(mult), float(add), 1.0, None, st), "augment(brightness)")
    if params.get("contrast") is not None and params["contrast"] != 1.0:
        stats = torch.empty(4, dtype=torch.float64, device=image.device)
        _lib.check(L.mmpl_patch_stats(_p(image), n, _p(stats), st), "patch_stats")
        _lib.check(L.mmpl_augment_patch(_p(image), n, 0.0, 0, 1.0, 0.0, float(params["contrast"]), _p(stats), st),
                   "augment(contrast)")
    return image


class PatchPipeline:
    """Per-sample pipeline object: ``pipeline(image, label, modality)`` -> the (image, label, catlas) triple the
    reference's dataset yields, computed on the device.  ``rng`` drives the crop origin and the augmentation draws."""

    def __init__(self, crop, atlas: Optional[torch.Tensor] = None, train: bool = True, augment: bool = True, seed: int = 0,
                 label_dtype=torch.float32):
        self.crop = tuple(int(c) for c in crop)
        self.atlas = atlas
        self.train, self.augment = train, augment
        self.rng = np.random.RandomState(seed)
        self.label_dtype = label_dtype

    def __call__(self, image, label, modality="ct"):
        dev = torch.device("cuda", torch.cuda.current_device())
        image = torch.as_tensor(image).to(dev, non_blocking=True)
        label = torch.as_tensor(label).to(dev, non_blocking=True)
        shape = tuple(min(a, b) for a, b in zip(image.shape, label.shape))
        origin = draw_crop_origin(shape, self.crop, self.rng) if self.train else (0, 0, 0)
        img, lab, cat = prepare_patch(image, label, self.crop, origin, modality, self.atlas, self.label_dtype)
        if self.train and self.augment:
            augment_patch(img, draw_augment_params(self.rng))
        return img, lab, cat
