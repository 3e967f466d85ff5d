"""Drop-in for the sliding-window inference and metrics of the reference's ``evaluate_amos.py``:
``_get_gaussian`` (:184-197), ``multi_net`` (:199-209), ``predict_sliding`` (:211-279), ``get_dice`` (:128-154),
``dice_score`` / ``senc_score`` / ``spec_score`` (:92-126) -- same names and call signatures.

What changes underneath: the reference copies every tile to the GPU and its logits back, then blends on the CPU in
float64 numpy (:242, :259-276).  Here the volume stays on the device: tiles are sliced on the GPU, the Gaussian-weighted
accumulation is one fused kernel per tile (mmpl_sw_blend, fp64 accumulators by default like the reference), and
normalise + argmax + per-class counting is one more (mmpl_sw_finalize).  With ``world_size > 1`` tiles are dealt
to the ranks in contiguous runs; the production path sends every touched accumulator plane to the owner of its depth slab
(a reduce-scatter restricted to the non-zero planes), finalises each slab where it lands and all-gathers the uint8 mask
(``_sliding_blend``).
"""
import os
from math import ceil

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib
from ._lib import p as _p


def _get_gaussian(patch_size, sigma_scale=1. / 8) -> np.ndarray:
    """Importance map of a tile (reference evaluate_amos.py:184-197): a unit impulse at the tile centre smoothed by
    scipy's ``gaussian_filter`` with sigma = size * sigma_scale per axis (zero boundary), scaled to a maximum of 1, cast to
    fp32, exact zeros lifted to the smallest positive value.  Host-side, once per tile size; the scipy call and the order
    of the casts are what make the map bit-identical to the reference's."""
    from scipy.ndimage import gaussian_filter

    impulse = np.zeros(patch_size)
    impulse[tuple(n // 2 for n in patch_size)] = 1
    blurred = gaussian_filter(impulse, [n * sigma_scale for n in patch_size], 0, mode='constant', cval=0)
    weight = (blurred / np.max(blurred) * 1).astype(np.float32)
    weight[weight == 0] = np.min(weight[weight != 0])
    return weight


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


def _axis_starts(extent, window, step):
    """Window starts along one axis: ceil((extent - window) / step) + 1 windows at multiples of ``step``, each pulled
    back so that it ends inside the volume (reference :218-239)."""
    count = int(ceil((extent - window) / step) + 1)
    return [max(min(i * step + window, extent) - window, 0) for i in range(count)]


def tile_origins(image_size, tile_size):
    """(d, y, x) origin of every window in the reference's depth -> row -> column order (:215-239), 25 % overlap; the
    in-plane step is derived from tile_size[1] alone, as in the reference (:217)."""
    step_hw = int(ceil(tile_size[1] * 0.75))
    step_d = int(ceil(tile_size[0] * 0.75))
    return [(d, y, x)
            for d in _axis_starts(image_size[2], tile_size[0], step_d)
            for y in _axis_starts(image_size[3], tile_size[1], step_hw)
            for x in _axis_starts(image_size[4], tile_size[2], step_hw)]


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


def _my_tiles(tiles, rank, world):
    """Contiguous run of the tile list for ``rank``: neighbouring tiles share depth levels, so a rank needs (and uploads)
    only the depth range of the volume its run covers."""
    per = (len(tiles) + world - 1) // world
    return tiles[rank * per:(rank + 1) * per]


def _accumulate(net_list, image, tile_size, classes, task_id, tta, acc_dtype, rank=0, world=1):
    """Generic path (any callable networks, TTA, fp64 or fp32 class-major accumulators like the reference)."""
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
    for d1, y1, x1 in _my_tiles(tile_origins(image.shape, tile_size), rank, world):
        img = image[:, :, d1:d1 + td, y1:y1 + th, x1:x1 + tw].contiguous()
        with torch.no_grad():
            pred = _tile_logits(net_list, img, task_id, tta).float().contiguous()
        assert tuple(pred.shape) == (B, classes, td, th, tw), f"network returned {tuple(pred.shape)}"
        for b in range(B):
            _lib.check(L.mmpl_sw_blend(_p(acc[b]), _p(wsum[b]), _p(pred[b]), _p(g), classes, D, H, W, td, th, tw,
                                       d1, y1, x1, nbytes, 0, _lib.stream_ptr()), "sw_blend")
    return acc, wsum


def predict_sliding(args, net_list, image, tile_size, classes, task_id, tta=False, acc_dtype=torch.float64):
    """Reference evaluate_amos.py:211-279.  Returns ``full_probs / count_predictions`` as a [B,C,D,H,W] tensor of
    ``acc_dtype`` (float64 like the reference) on the current CUDA device."""
    acc, wsum = _accumulate(net_list, image, tile_size, classes, task_id, tta, acc_dtype)
    return acc / wsum.unsqueeze(1)


def _blender_for(net_list, volume_dhw, tile_size, classes, world):
    """The fused tile-blend engine for ``net_list`` if it has one: an ``engine.GraphedSlidingWindow`` of matching
    geometry, or a bare model with ``blend_tile`` (eager launches)."""
    if len(net_list) != 1:
        return None
    net = net_list[0]
    if hasattr(net, "blend_tile") and hasattr(net, "acc"):
        ok = (net.volume_dhw == tuple(volume_dhw) and net.tile == tuple(int(t) for t in tile_size)
              and net.classes == classes and net.world == world)
        return net if ok else None
    if isinstance(net, torch.nn.Module) and hasattr(net, "blend_supported") and not net.training and net.blend_supported():
        return _EagerBlender(net, volume_dhw, tile_size, classes, world)
    return None


class _EagerBlender:
    """Same contract as engine.GraphedSlidingWindow with per-kernel launches (any volume shape, no capture cost): up to
    ``tile_batch`` tiles of the volume per forward, accumulated tile by tile in list order."""

    def __init__(self, model, volume_dhw, tile, classes, world, tile_batch=8):
        dev = torch.device("cuda", torch.cuda.current_device())
        D, H, W = volume_dhw
        self.model, self.world, self.tile, self.tile_batch = model, world, tuple(int(t) for t in tile), int(tile_batch)
        self.dpad = (D + world - 1) // world * world
        self.acc = torch.zeros((1, self.dpad, classes, H, W), dtype=torch.float32, device=dev)
        self._sinks = {}

    def _sink(self, tiles):
        from . import ops

        if tiles not in self._sinks:
            origin_dev = torch.zeros((tiles, 3), dtype=torch.int32, device=self.acc.device)
            self._sinks[tiles] = ops.BlendSink(self.acc, _gaussian_device(self.tile, self.acc.device), origin_dev,
                                               self.tile, d_outer=True)
        return self._sinks[tiles]

    def reset(self, lo=0, hi=None):
        self.acc[0, lo:hi].zero_()

    def blend_tile(self, img, origin_dev_row):
        self.blend_tiles([img], origin_dev_row.reshape(1, 3))

    def blend_tiles(self, imgs, origins_dev):
        for i in range(0, len(imgs), self.tile_batch):
            chunk = imgs[i:i + self.tile_batch]
            sink = self._sink(len(chunk))
            sink.origin_dev.copy_(origins_dev[i:i + len(chunk)].reshape(len(chunk), 3), non_blocking=True)
            with torch.no_grad():
                self.model.blend_tile(chunk[0] if len(chunk) == 1 else torch.cat(chunk), sink)


_SW_TRACE = os.environ.get("MMPL_SW_TRACE", "0") == "1"


class _PhaseTrace:
    """MMPL_SW_TRACE=1: device time of the phases of one sharded sliding-window volume (CUDA events on the current stream),
    printed per rank to stderr.  Diagnostic only (it synchronises at the end of the volume)."""

    def __init__(self):
        self.ev = [("start", torch.cuda.Event(enable_timing=True))]
        self.ev[0][1].record()

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append((name, e))

    def report(self, rank):
        import sys

        torch.cuda.synchronize()
        parts = [f"{n} {a.elapsed_time(b):.2f}" for (_, a), (n, b) in zip(self.ev[:-1], self.ev[1:])]
        sys.stderr.write(f"[sw rank {rank}] ms: " + ", ".join(parts) + "\n")


def _depth_range(tiles, td):
    """Depth planes [lo, hi) touched by a run of tiles ((0, 0) for an empty run)."""
    if not tiles:
        return 0, 0
    return min(t[0] for t in tiles), max(t[0] for t in tiles) + td


def _overlap(rng, lo, hi):
    a, b = max(rng[0], lo), min(rng[1], hi)
    return (a, b) if b > a else (a, a)


def _exchange_plan(ranges, rank, slab, world):
    """Who sends which accumulator planes to whom.  ``ranges[r]`` = depth planes [lo, hi) rank r accumulated into; rank o
    owns [o*slab, (o+1)*slab).  -> (sends, recvs) for ``rank``: lists of (peer, lo, hi) in absolute plane indices --
    ``sends``: my planes inside the peer's slab; ``recvs``: the peer's planes inside my slab.  Both sides derive the same
    (lo, hi) from the same tile list, so message sizes agree without any handshake."""
    sends, recvs = [], []
    for r in range(world):
        if r == rank:
            continue
        lo, hi = _overlap(ranges[rank], r * slab, (r + 1) * slab)
        if hi > lo:
            sends.append((r, lo, hi))
        lo, hi = _overlap(ranges[r], rank * slab, (rank + 1) * slab)
        if hi > lo:
            recvs.append((r, lo, hi))
    return sends, recvs


def _sliding_blend(blender, image, tile_size, classes, label, num_class, rank, world):
    """Production path of predict_sliding_dice.  Every rank: upload the depth range its tiles need, run
    forward + classifier + Gaussian accumulation per tile into a depth-major fp32 accumulator.  Then the exchange along
    depth: rank o owns the depth slab [o*slab, (o+1)*slab) and receives, from every rank whose tiles touched it, exactly the
    planes of that slab the sender's depth range covers (one grouped NCCL send/recv batch -- a reduce-scatter restricted to
    the non-zero planes: with contiguous tile runs a rank's accumulator is empty outside ~1/world of the depth, so this
    moves 3-4x fewer bytes than ``reduce_scatter_tensor`` over the whole accumulator).  The owner adds the received planes,
    takes argmax + Dice counts on its slab, the uint8 mask is all-gathered and 3 x C counters are all-reduced.  The weight
    sum is never formed: argmax and Dice do not depend on a positive per-voxel normaliser."""
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    trace = _PhaseTrace() if _SW_TRACE else None
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(image)
    B, _, D, H, W = image.shape
    assert B == 1, "the fused sliding-window path handles one volume per call"
    td, th, tw = (int(t) for t in tile_size)
    tiles = tile_origins(image.shape, tile_size)
    runs = [_my_tiles(tiles, r, world) for r in range(world)]
    ranges = [_depth_range(m, td) for m in runs]
    mine = runs[rank]
    acc = blender.acc[0]                                   # [Dpad, C, H, W]
    dpad = acc.shape[0]
    slab = dpad // world
    z0 = rank * slab
    dlo, dhi = ranges[rank]
    # only the planes this rank writes (its tiles) or owns (its slab) are read later: zero their hull, not 5 GB
    blender.reset(min(dlo, z0) if mine else z0, max(dhi, z0 + slab))
    if mine:
        part = image[:, :, dlo:dhi].to(dev, torch.float32, non_blocking=True)       # contiguous depth range (B = 1)
        origins = torch.tensor(mine, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
        views = [part[:, :, d1 - dlo:d1 - dlo + td, y1:y1 + th, x1:x1 + tw] for d1, y1, x1 in mine]
        if hasattr(blender, "blend_tiles"):         # several tiles per forward, accumulated in this order
            blender.blend_tiles(views, origins)
        else:
            for i, tile in enumerate(views):
                blender.blend_tile(tile, origins[i])
    if trace:
        trace.mark("tiles")
    if world > 1:
        sends, recvs = _exchange_plan(ranges, rank, slab, world)
        p2p, received = [], []
        for r, lo, hi in sends:
            p2p.append(dist.P2POp(dist.isend, acc[lo:hi], r))
        for r, lo, hi in recvs:
            buf = torch.empty((hi - lo, classes, H, W), dtype=torch.float32, device=dev)
            p2p.append(dist.P2POp(dist.irecv, buf, r))
            received.append((lo, hi, buf))
        if p2p:
            for req in dist.batch_isend_irecv(p2p):
                req.wait()
        if trace:
            trace.mark("exchange")
        for lo, hi, buf in received:
            _lib.check(L.mmpl_accumulate_f32(_p(acc[lo:hi]), _p(buf), buf.numel(), _lib.stream_ptr()), "accumulate_f32")
    if trace:
        trace.mark("accumulate")
    mine_acc = acc[z0:z0 + slab]
    amax_slab = torch.empty((slab, H, W), dtype=torch.uint8, device=dev)
    counts = torch.zeros((3, classes), dtype=torch.int64, device=dev)
    lab = None
    if label is not None:
        if isinstance(label, np.ndarray):
            label = torch.from_numpy(label)
        lab_full = label.reshape(D, H, W)
        lab = lab_full[z0:min(z0 + slab, D)].to(dev, non_blocking=True)
        lab = lab.contiguous() if lab.dtype == torch.uint8 else lab.float().contiguous()
    valid = max(min(z0 + slab, D) - z0, 0)                 # planes of this slab inside the volume (the rest is padding)
    if valid > 0:
        _lib.check(L.mmpl_sw_finalize(_p(mine_acc), None, None if lab is None else _p(lab),
                                      int(lab is not None and lab.dtype == torch.uint8), None, _p(amax_slab),
                                      _p(counts) if lab is not None else None, classes, valid * H * W, H * W, 4,
                                      _lib.stream_ptr()), "sw_finalize")
    if valid < slab:
        amax_slab[valid:].zero_()
    if trace:
        trace.mark("finalize")
    if world > 1:
        amax = torch.empty((dpad, H, W), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(amax, amax_slab)
        if lab is not None:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        amax = amax[:D]
    else:
        amax = amax_slab[:D]
    if trace:
        trace.mark("gather")
        trace.report(rank)
    amax = amax.unsqueeze(0)
    if label is None:
        return None, None, None, amax
    return _scores(counts.unsqueeze(0), num_class) + (amax,)


def predict_sliding_dice(args, net_list, image, tile_size, classes, task_id, label=None, tta=False,
                         acc_dtype=torch.float64, num_class=None, sharded=False):
    """Fused variant: blend, normalise, argmax and per-class Dice counts without materialising the normalised logit
    volume.  Returns (dices, senc, spec, argmax uint8 [B,D,H,W]).

    ``acc_dtype=torch.float32`` with a single bf16 network that supports it (a ``unet3D_baseline`` in eval mode, or an
    ``engine.GraphedSlidingWindow``) takes the production path: classifier + Gaussian accumulation in one kernel, and with
    ``sharded=True`` a contiguous run of tiles per rank, plane exchange along depth, local argmax/Dice, all-gather of the
    uint8 mask (``_sliding_blend``).  Otherwise the generic path (fp64 accumulators like the reference, TTA, several
    networks); sharded: tiles split over the ranks, accumulators summed with all-reduce."""
    world = dist.get_world_size() if (sharded and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    num_class = num_class if num_class is not None else classes - 1
    shape = image.shape
    if acc_dtype == torch.float32 and not tta and shape[0] == 1:
        blender = _blender_for(net_list, tuple(int(v) for v in shape[2:]), tile_size, classes, world)
        if blender is not None:
            return _sliding_blend(blender, image, tile_size, classes, label, num_class, rank, world)
    acc, wsum = _accumulate(net_list, image, tile_size, classes, task_id, tta, acc_dtype, rank, world)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(wsum, op=dist.ReduceOp.SUM)
    return _finalize(acc, wsum, label, classes, num_class)


def _scores(counts, num_class):
    """counts int64 [B][3][C] = |P & T|, |P|, |T| -> per-class lists (dice_score / senc_score / spec_score, :92-126)."""
    inter, npred, ntgt = counts[:, 0].double(), counts[:, 1].double(), counts[:, 2].double()
    dices = [(2 * inter[:, l] / (npred[:, l] + ntgt[:, l] + 1)).mean() for l in range(1, num_class + 1)]
    senc = [(inter[:, l] / (ntgt[:, l] + 1)).mean() for l in range(1, num_class + 1)]
    spec = [(inter[:, l] / (npred[:, l] + 1)).mean() for l in range(1, num_class + 1)]
    return dices, senc, spec


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
        lab = label.to(dev).reshape(B, -1)
        lab = lab.contiguous() if lab.dtype == torch.uint8 else lab.float().contiguous()
    u8 = int(lab is not None and lab.dtype == torch.uint8)
    for b in range(B):
        _lib.check(L.mmpl_sw_finalize(_p(acc[b]), None if wsum is None else _p(wsum[b]), None if lab is None else _p(lab[b]),
                                      u8, None, _p(amax[b]), _p(counts[b]) if lab is not None else None, classes, vox, 0,
                                      acc.element_size(), _lib.stream_ptr()), "sw_finalize")
    if lab is None:
        return None, None, None, amax
    return _scores(counts, num_class) + (amax,)


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
