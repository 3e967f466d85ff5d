"""Drop-in for the sliding-window inference and metrics of the reference's ``evaluate_amos.py``:
``_get_gaussian`` (:184-197), ``multi_net`` (:199-209), ``predict_sliding`` (:211-279), ``get_dice`` (:128-154),
``dice_score`` / ``senc_score`` / ``spec_score`` (:92-126) -- same names and call signatures.

What changes underneath: the reference copies every tile to the GPU and its logits back, then blends on the CPU in
float64 numpy (:242, :259-276).  Here the volume stays on the device: tiles are sliced on the GPU, the Gaussian-weighted
accumulation is one fused kernel per tile (mmpl_sw_blend, fp64 accumulators by default like the reference), and
normalise + argmax + per-class counting is one more (mmpl_sw_finalize).  With ``world_size > 1`` tiles are dealt
round-robin to the ranks and the accumulators are summed with one all-reduce (``predict_sliding_sharded``).
"""
from math import ceil

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib
from ._lib import p as _p


def _get_gaussian(patch_size, sigma_scale=1. / 8) -> np.ndarray:
    """Reference evaluate_amos.py:184-197 (scipy on the host; computed once per tile size)."""
    from scipy.ndimage import gaussian_filter

    tmp = np.zeros(patch_size)
    center_coords = [i // 2 for i in patch_size]
    sigmas = [i * sigma_scale for i in patch_size]
    tmp[tuple(center_coords)] = 1
    g = gaussian_filter(tmp, sigmas, 0, mode='constant', cval=0)
    g = (g / np.max(g) * 1).astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


_gauss_cache = {}


def _gaussian_device(tile_size, device):
    key = (tuple(int(t) for t in tile_size), str(device))
    if key not in _gauss_cache:
        _gauss_cache[key] = torch.from_numpy(_get_gaussian(key[0])).to(device).contiguous()
    return _gauss_cache[key]


def multi_net(net_list, img, task_id):
    """Mean of the networks' logits (reference :199-209)."""
    pred = net_list[0](img, task_id)
    for i in range(1, len(net_list)):
        pred = pred + net_list[i](img, task_id)
    if len(net_list) > 1:
        pred = pred / len(net_list)
    return pred


def tile_origins(image_size, tile_size):
    """(d1, y1, x1) of every window in the reference's dep -> row -> col order (:215-239); the H/W stride is derived
    from tile_size[1] only, as in the reference (:217)."""
    overlap = 1 / 4
    strideHW = ceil(tile_size[1] * (1 - overlap))
    strideD = ceil(tile_size[0] * (1 - overlap))
    tile_deps = int(ceil((image_size[2] - tile_size[0]) / strideD) + 1)
    tile_rows = int(ceil((image_size[3] - tile_size[1]) / strideHW) + 1)
    tile_cols = int(ceil((image_size[4] - tile_size[2]) / strideHW) + 1)
    out = []
    for dep in range(tile_deps):
        for row in range(tile_rows):
            for col in range(tile_cols):
                d1, x1, y1 = int(dep * strideD), int(col * strideHW), int(row * strideHW)
                d2 = min(d1 + tile_size[0], image_size[2])
                x2 = min(x1 + tile_size[2], image_size[4])
                y2 = min(y1 + tile_size[1], image_size[3])
                out.append((max(int(d2 - tile_size[0]), 0), max(int(y2 - tile_size[1]), 0), max(int(x2 - tile_size[2]), 0)))
    return out


def _tile_logits(net_list, img, task_id, tta):
    pred = multi_net(net_list, img, task_id)
    if tta:                                                                       # reference :247-255
        # a network may hand out a static buffer that its next call overwrites (engine.GraphedInference): take a
        # private copy of the un-flipped prediction before asking for the next one
        pred = pred.clone()
        for dims in ([2], [3], [4], [2, 3], [2, 4], [3, 4], [2, 3, 4]):
            pred += torch.flip(multi_net(net_list, torch.flip(img, dims), task_id), dims)
        pred /= 8.
    return pred


def _accumulate(net_list, image, tile_size, classes, task_id, tta, acc_dtype, rank=0, world=1):
    _lib.require_device()
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(image)
    image = image.to(dev, torch.float32)
    B, _, D, H, W = image.shape
    g = _gaussian_device(tile_size, dev)
    acc = torch.zeros((B, classes, D, H, W), dtype=acc_dtype, device=dev)
    wsum = torch.zeros((B, D, H, W), dtype=acc_dtype, device=dev)
    nbytes = acc.element_size()
    td, th, tw = (int(t) for t in tile_size)
    for i, (d1, y1, x1) in enumerate(tile_origins(image.shape, tile_size)):
        if i % world != rank:
            continue
        img = image[:, :, d1:d1 + td, y1:y1 + th, x1:x1 + tw].contiguous()
        with torch.no_grad():
            pred = _tile_logits(net_list, img, task_id, tta).float().contiguous()
        assert tuple(pred.shape) == (B, classes, td, th, tw), f"network returned {tuple(pred.shape)}"
        for b in range(B):
            _lib.check(L.mmpl_sw_blend(_p(acc[b]), _p(wsum[b]), _p(pred[b]), _p(g), classes, D, H, W, td, th, tw,
                                       d1, y1, x1, nbytes, _lib.stream_ptr()), "sw_blend")
    return acc, wsum


def predict_sliding(args, net_list, image, tile_size, classes, task_id, tta=False, acc_dtype=torch.float64):
    """Reference evaluate_amos.py:211-279.  Returns ``full_probs / count_predictions`` as a [B,C,D,H,W] tensor of
    ``acc_dtype`` (float64 like the reference) on the current CUDA device."""
    acc, wsum = _accumulate(net_list, image, tile_size, classes, task_id, tta, acc_dtype)
    return acc / wsum.unsqueeze(1)


def predict_sliding_dice(args, net_list, image, tile_size, classes, task_id, label=None, tta=False,
                         acc_dtype=torch.float64, num_class=None, sharded=False):
    """Fused variant: blend, normalise, argmax and per-class Dice counts without materialising the normalised logit
    volume.  Returns (dices, senc, spec, argmax uint8 [B,D,H,W]).  ``sharded=True`` deals tiles round-robin over the
    ranks of the default process group and sums the accumulators with one all-reduce."""
    world = dist.get_world_size() if (sharded and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    acc, wsum = _accumulate(net_list, image, tile_size, classes, task_id, tta, acc_dtype, rank, world)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(wsum, op=dist.ReduceOp.SUM)
    return _finalize(acc, wsum, label, classes, num_class if num_class is not None else classes - 1)


def _finalize(acc, wsum, label, classes, num_class):
    L = _lib.lib()
    B = acc.shape[0]
    vox = acc[0, 0].numel()
    dev = acc.device
    amax = torch.empty(acc.shape[0:1] + acc.shape[2:], dtype=torch.uint8, device=dev)
    counts = torch.zeros((B, 3, classes), dtype=torch.int64, device=dev)
    lab = None
    if label is not None:
        if isinstance(label, np.ndarray):
            label = torch.from_numpy(label)
        lab = label.to(dev, torch.float32).reshape(B, -1).contiguous()
    for b in range(B):
        _lib.check(L.mmpl_sw_finalize(_p(acc[b]), None if wsum is None else _p(wsum[b]), None if lab is None else _p(lab[b]),
                                      None, _p(amax[b]), _p(counts[b]) if lab is not None else None, classes, vox,
                                      acc.element_size(), _lib.stream_ptr()), "sw_finalize")
    if lab is None:
        return None, None, None, amax
    inter, npred, ntgt = counts[:, 0].double(), counts[:, 1].double(), counts[:, 2].double()
    dices = [(2 * inter[:, l] / (npred[:, l] + ntgt[:, l] + 1)).mean() for l in range(1, num_class + 1)]
    senc = [(inter[:, l] / (ntgt[:, l] + 1)).mean() for l in range(1, num_class + 1)]
    spec = [(inter[:, l] / (npred[:, l] + 1)).mean() for l in range(1, num_class + 1)]
    return dices, senc, spec, amax


def dice_score(preds, labels):
    assert preds.shape[0] == labels.shape[0], "predict & target batch size don't match"
    predict = preds.contiguous().view(preds.shape[0], -1).double()
    target = labels.contiguous().view(labels.shape[0], -1).double()
    num = torch.sum(torch.mul(predict, target), dim=1)
    den = torch.sum(predict, dim=1) + torch.sum(target, dim=1) + 1
    return (2 * num / den).mean()


def spec_score(preds, labels):
    predict = preds.contiguous().view(preds.shape[0], -1).double()
    target = labels.contiguous().view(labels.shape[0], -1).double()
    return (torch.sum(predict * target, dim=1) / (torch.sum(predict, dim=1) + 1)).mean()


def senc_score(preds, labels):
    predict = preds.contiguous().view(preds.shape[0], -1).double()
    target = labels.contiguous().view(labels.shape[0], -1).double()
    return (torch.sum(predict * target, dim=1) / (torch.sum(target, dim=1) + 1)).mean()


def get_dice(preds, labels, t_id, atlas=None, num_class=13):
    """Reference evaluate_amos.py:128-154: argmax over classes, then per class l = 1..num_class Dice / sensitivity /
    specificity-like scores with +1 smoothing -- one fused kernel (argmax of softmax == argmax of logits).  With an
    ``atlas`` prior [B, num_class, D, H, W] the per-class prediction is ``softmax(preds)[:, l+1] + 0.15 > 1 - atlas[:, l]``
    instead (:142-151; a validation-time option of the reference, plain tensor ops here).  Returns
    (dices, senc, spec, preds) with ``preds`` the argmax volume."""
    _lib.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = preds.to(dev)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    x = x.contiguous()
    dices, senc, spec, amax = _finalize(x, None, labels, x.shape[1], num_class)
    if atlas is not None:
        preds_r = F.softmax(x, dim=1)
        lab = labels.to(dev)
        atlas = atlas.to(dev)
        dices, senc, spec = [], [], []
        for l in range(num_class):
            cpred = (preds_r[:, l + 1] + 0.15) > (1 - atlas[:, l])
            tgt = (lab == (l + 1)).reshape(cpred.shape) if lab.numel() == cpred.numel() else lab == (l + 1)
            dices.append(dice_score(cpred, tgt))
            senc.append(senc_score(cpred, tgt))
            spec.append(spec_score(cpred, tgt))
    return dices, senc, spec, amax.long()
