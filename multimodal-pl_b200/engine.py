"""Drop-in for the reference's ``engine.py`` with the data-parallel hooks actually implemented.

The reference ``Engine`` (engine.py:10-77) is a stub: ``data_parallel`` returns the model unchanged (:30-32),
``all_reduce_tensor`` is ``torch.mean`` (:57-58), samplers are ``None`` (:44).  The authors' log shows the code was
run as 3-process DDP (run_files/amos_ours_77.txt:4-6).  Here the same method names are backed by one process per GPU
(``torchrun`` env: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*), NCCL over NVLink for the gradient all-reduce
(gloo on CPU-only hosts, used by the tests), and a fused SGD step on a flat parameter buffer.

Training is embarrassingly parallel over samples (GroupNorm has no cross-sample statistics) with ONE exchange step:
the gradient all-reduce.  Gradients live in one flat fp32 buffer cut into buckets in reverse registration order;
each bucket is all-reduced asynchronously as soon as its last gradient has been produced, so the collective overlaps
the rest of the backward pass.
"""
import argparse
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class DataParallelModel(torch.nn.Module):
    """Wraps a module for bucketed, overlapped gradient averaging.  ``.module`` is the wrapped model (the train loop
    uses ``model.module.renew_token``, train_amos_atlas_final.py:391).

    Gradients live in ONE flat fp32 buffer (the backward kernels write into it directly, ``ops._grad_dst``), cut into
    buckets in reverse registration order = the order in which backward produces them.  As soon as the last gradient of
    a bucket exists its NCCL all-reduce is issued asynchronously (``_on_grad``), so the exchange of the big low-resolution
    layers runs underneath the long full-resolution tail of the backward pass; only the last, small bucket (stem,
    layer0..2) is exposed.  ``bucket_step`` (set by ``GraphedTrainStep`` / callers that own the optimizer) is applied to
    each bucket right after its all-reduce has landed, so the optimizer step of bucket i overlaps the exchange of i+1.
    All of it is stream-ordered and capturable: inside a CUDA graph the collectives become graph nodes."""

    def __init__(self, module: torch.nn.Module, world_size: int, bucket_mb: float = 8.0, average: bool = True):
        super().__init__()
        self.module = module
        self.world_size = world_size
        self.average = average
        params = [p for p in module.parameters() if p.requires_grad]
        self._params = params
        if world_size > 1:
            for p in module.parameters():
                dist.broadcast(p.data, src=0)
            for b in module.buffers():
                dist.broadcast(b.data, src=0)
        total = sum(p.numel() for p in params)
        dev = params[0].device if params else torch.device("cpu")
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        # gradients are produced roughly in reverse registration order: bucket 0 holds the LAST parameters
        self._buckets: List[dict] = []
        limit = int(bucket_mb * 1024 * 1024 / 4)
        off = total
        cur = {"hi": total, "lo": total, "pending": 0, "count": 0, "handle": None}
        for p in reversed(params):
            off -= p.numel()
            # the backward kernels write parameter gradients straight into this slot (ops._grad_dst); gradients that
            # arrive any other way are folded in by _adopt_grad
            p._mmpl_grad_slot = (self.flat_grad, off, p.numel())
            p.grad = None
            p._mmpl_bucket = len(self._buckets)
            cur["lo"] = off
            cur["count"] += 1
            if cur["hi"] - cur["lo"] >= limit:
                self._buckets.append(cur)
                cur = {"hi": off, "lo": off, "pending": 0, "count": 0, "handle": None}
        if cur["count"]:
            self._buckets.append(cur)
        self._callback_queued = False
        self.sync_in_backward = True      # False: the caller reduces flat_grad itself (all_reduce_flat)
        self.bucket_step = None           # callable(lo, hi): optimizer step on flat range [lo, hi) once it is reduced
        if world_size > 1:
            for p in params:
                p.register_post_accumulate_grad_hook(self._on_grad)
        self._reset()

    def _reset(self):
        for b in self._buckets:
            b["pending"] = b["count"]
            b["handle"] = None
        self._callback_queued = False

    def all_reduce_flat(self):
        """One blocking all-reduce of the whole flat gradient buffer (the un-overlapped fallback)."""
        ops.join_side_stream()          # weight gradients are finished on a side stream (ops._ws_bwd_launch)
        for p in self._params:
            _adopt_grad(p)
        if self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            if self.average:
                self.flat_grad.mul_(1.0 / self.world_size)

    def _on_grad(self, p):
        if not self.sync_in_backward:
            return
        _adopt_grad(p)
        if not self._callback_queued:
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)
            self._callback_queued = True
        b = self._buckets[p._mmpl_bucket]
        b["pending"] -= 1
        if b["pending"] == 0:
            ops.join_side_stream()      # weight gradients are finished on a side stream (ops._ws_bwd_launch)
            view = self.flat_grad[b["lo"]:b["hi"]]
            b["handle"] = dist.all_reduce(view, op=dist.ReduceOp.SUM, async_op=True)

    def _finish(self):
        # buckets that never completed (a parameter without gradient, e.g. eam*.proj of unet3D_with_feam3) are reduced
        # here without having passed through _on_grad's join: order the side-stream writes before any collective
        ops.join_side_stream()
        for b in self._buckets:
            view = self.flat_grad[b["lo"]:b["hi"]]
            if b["handle"] is not None:
                b["handle"].wait()
            else:
                dist.all_reduce(view, op=dist.ReduceOp.SUM)
            if self.average:
                view.mul_(1.0 / self.world_size)
            if self.bucket_step is not None:
                self.bucket_step(b["lo"], b["hi"])
        self._reset()

    def zero_grad(self, set_to_none: bool = True):
        _zero_flat(self._params, self.flat_grad)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def _adopt_grad(p):
    """Make ``p.grad`` the view of p's slot in the flat gradient buffer.  The library's backward kernels write there
    directly and autograd adopts that view (nothing to do); a gradient produced any other way (foreign autograd
    function, cloned by autograd) is copied into the slot."""
    slot = getattr(p, "_mmpl_grad_slot", None)
    if slot is None or p.grad is None:
        return
    flat, off, n = slot
    if p.grad.data_ptr() != flat.data_ptr() + 4 * off or p.grad.dtype != torch.float32:
        view = flat[off:off + n].view_as(p)
        view.copy_(p.grad)
        p.grad = view


def _zero_flat(params, flat):
    """Start of a step: clear the flat buffer (parameters that receive no gradient must contribute zero) and drop the
    per-parameter views so the next backward writes in place instead of accumulating."""
    flat.zero_()
    for p in params:
        p.grad = None


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD(lr, momentum, weight_decay) semantics (train_amos_atlas_final.py:132-135) as ONE kernel launch
    over flat parameter / gradient / momentum buffers (mmpl_sgd_step).  ``param_groups[0]['lr']`` stays writable for
    the reference's poly schedule (utils.py:56-60).  Parameters are re-pointed into the flat buffer (values kept)."""

    def __init__(self, params, lr=1e-2, momentum=0.9, weight_decay=1e-4, flat_grad: Optional[torch.Tensor] = None):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        own_grad = flat_grad is None
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev) if own_grad else flat_grad
        assert self.flat_grad.numel() == total
        off = 0
        for p in params:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1).float())
            p.data = self.flat_param[off:off + n].view_as(p)
            if own_grad:
                p._mmpl_grad_slot = (self.flat_grad, off, n)
                p.grad = None
            else:
                slot = getattr(p, "_mmpl_grad_slot", None)
                assert slot is not None and slot[0] is self.flat_grad and slot[1] == off, \
                    "flat_grad must be laid out in parameter order (use DataParallelModel.flat_grad)"
            off += n
        self._params = params
        self.momentum_buf = torch.zeros(total, dtype=torch.float32, device=dev)
        self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._lr_host = None
        self._steps = 0

    def zero_grad(self, set_to_none: bool = True):
        _zero_flat(self._params, self.flat_grad)

    def _sync_lr(self):
        """Mirror param_groups[0]['lr'] into the device scalar the kernel reads (outside any graph capture)."""
        g = self.param_groups[0]
        if self._lr_host != g["lr"]:
            self._lr_dev.fill_(float(g["lr"]))
            self._lr_host = g["lr"]

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, flat_range=None, last: bool = True):
        """``flat_range=(lo, hi)`` restricts the step to that range of the flat buffers (one reduced gradient bucket);
        ``last=False`` marks a partial step that further ranges of the same optimisation step will follow."""
        g = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self._sync_lr()
        for p in self._params:      # gradients that did not land in the flat buffer by themselves
            _adopt_grad(p)
        _lib.require_device()
        # torch.optim.SGD skips parameters whose .grad is None (no weight decay, no momentum update): e.g. eam*.proj of
        # unet3D_with_feam3 never receives a gradient.  One launch per maximal run of parameters that have one -- a
        # single launch for the backbone.  Under CUDA-graph capture the runs are those of the captured step.
        first = int(self._steps == 0)
        for lo, hi in self._active_ranges():
            if flat_range is not None:
                lo, hi = max(lo, flat_range[0]), min(hi, flat_range[1])
                if lo >= hi:
                    continue
            _lib.check(_lib.lib().mmpl_sgd_step(self.flat_param.data_ptr() + 4 * lo, self.flat_grad.data_ptr() + 4 * lo,
                                                self.momentum_buf.data_ptr() + 4 * lo, hi - lo,
                                                self._lr_dev.data_ptr(), float(g["momentum"]), float(g["weight_decay"]),
                                                float(grad_scale), first, _lib.stream_ptr()), "sgd_step")
        if last:
            self._steps += 1
        ops.bump_weights_epoch()       # weights changed behind autograd's back: see ops._check_ws_stamp

    def _active_ranges(self):
        """[lo, hi) element ranges of the flat buffers covering exactly the parameters with a gradient this step."""
        out, off, start = [], 0, None
        for p in self._params:
            n = p.numel()
            if p.grad is not None:
                if start is None:
                    start = off
            elif start is not None:
                out.append((start, off))
                start = None
            off += n
        if start is not None:
            out.append((start, off))
        return out


class GraphedTrainStep:
    """One train step (zero_grad -> forward -> loss -> backward -> gradient exchange -> fused SGD) captured ONCE into a
    CUDA graph and replayed per batch: ~250 kernel launches per step cost one graph launch on the host, so the GPU
    never waits for Python.  The reference drives every op from the Python loop (train_amos_atlas_final.py:258-378).

    ``loss_fn(logits, labels) -> scalar``.  With world_size > 1 the NCCL exchange is INSIDE the graph: the bucketed
    asynchronous all-reduces of ``DataParallelModel`` are captured on NCCL's stream as parallel branches that overlap the
    rest of the backward pass, and the SGD step of each bucket is captured right behind its all-reduce.
    (``comm_in_graph=False`` / MMPL_GRAPH_NCCL=0 keeps the exchange outside: one all-reduce of the flat buffer after the
    replay.)  The constructor runs ``warmup`` real optimisation steps on the example batch (capture needs warmed-up
    allocators, lazily initialised kernels and an initialised NCCL communicator).

    Inputs: ``step(image, label)`` copies them into the graph's static buffers.  Host batches go through a staging
    buffer filled on a copy stream: ``step.stage(next_image, next_label)`` may be called while the previous step still
    runs, so the PCIe transfer of batch i+1 overlaps the compute of batch i; ``step.run_staged()`` then costs one
    device-to-device copy.  Labels may be uint8 (a quarter of the fp32 bytes).

    ``fused_loss(module, image, label) -> loss`` (optional) replaces ``loss_fn(module(image)[0], label)``; with
    ``unet3D_baseline.forward_partial_loss`` the step never materialises logits."""

    def __init__(self, dp_model: "DataParallelModel", loss_fn, optimizer: "FusedSGD", image, label, warmup: int = 3,
                 comm_in_graph: Optional[bool] = None, fused_loss=None):
        self.dp, self.loss_fn, self.opt = dp_model, loss_fn, optimizer
        # fused_loss(module, image, label) -> loss replaces  loss_fn(module(image)[0], label): e.g.
        # lambda m, x, y: m.forward_partial_loss(x, y, masks, lut=luts, per_sample=True)  (no logits tensor at all)
        self.fused_loss = fused_loss
        self.world = dp_model.world_size
        if comm_in_graph is None:
            comm_in_graph = os.environ.get("MMPL_GRAPH_NCCL", "1") != "0"
        self.comm_in_graph = bool(comm_in_graph) and self.world > 1
        dev = dp_model.flat_grad.device
        self.static_image = image.to(dev, copy=True)
        self.static_label = label.to(dev, copy=True)
        self._stage_img = torch.empty_like(self.static_image)
        self._stage_lab = torch.empty_like(self.static_label)
        self._copy_stream = torch.cuda.Stream()
        self._stage_ready, self._stage_free = torch.cuda.Event(), torch.cuda.Event()
        self._stage_free.record()
        self._prev_sync = dp_model.sync_in_backward
        self.opt._sync_lr()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._body(in_graph=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        try:   # gradients are views of a flat buffer created on the default stream: the mismatch is intentional
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        try:
            self._capture()
        except Exception:
            if not self.comm_in_graph:
                raise
            # NCCL refused to be captured (old NCCL, a watchdog interfering, ...): exchange outside the graph instead
            self.comm_in_graph = False
            torch.cuda.synchronize()
            self._capture()

    def _capture(self):
        self.graph = torch.cuda.CUDAGraph()
        # capture on a HIGH-priority stream: the critical path (forward, dgrad, GroupNorm backward) then wins the SMs
        # over the weight-gradient branch that ops forks onto its default-priority side stream
        hp = None
        if os.environ.get("MMPL_GRAPH_PRIORITY", "1") != "0":
            lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
            hp = torch.cuda.Stream(priority=hi)
        # thread_local: NCCL's watchdog thread polls events while we capture; only OUR thread's calls are policed
        with torch.cuda.graph(self.graph, stream=hp, capture_error_mode="thread_local"):
            self.static_loss = self._body(in_graph=True)

    def close(self):
        """Release the captured graph.  With world_size > 1 the graph holds NCCL kernels: call this (or drop the last
        reference to the object) BEFORE ``dist.destroy_process_group()``, which otherwise waits on the communicator the
        graph still pins."""
        torch.cuda.synchronize()
        graph, self.graph = getattr(self, "graph", None), None
        if graph is not None:
            graph.reset()

    def _bucket_step(self, lo, hi):
        self.opt.step(grad_scale=1.0 if self.dp.average else 1.0 / self.world, flat_range=(lo, hi), last=False)

    def _body(self, in_graph: bool):
        overlapped = self.world > 1 and self.comm_in_graph
        self.dp.sync_in_backward = overlapped
        self.dp.bucket_step = self._bucket_step if overlapped else None
        try:
            self.opt.zero_grad()
            if self.fused_loss is not None:
                loss = self.fused_loss(self.dp.module, self.static_image, self.static_label)
            else:
                logits = self.dp(self.static_image, self.static_label)
                logits = logits[0] if isinstance(logits, (tuple, list)) else logits
                loss = self.loss_fn(logits, self.static_label)
            loss.backward()      # overlapped: bucket all-reduces + per-bucket SGD are issued from the autograd hooks
        finally:
            self.dp.bucket_step = None
            self.dp.sync_in_backward = self._prev_sync
        if overlapped:
            self.opt._steps += 1
        elif self.world == 1:
            self.opt.step()
        elif not in_graph:
            self._exchange_outside()
        return loss

    def _exchange_outside(self):
        self.dp.all_reduce_flat()
        self.opt.step(grad_scale=1.0 if self.dp.average else 1.0 / self.world)

    def stage(self, image, label):
        """Start copying the next batch (pinned host or device tensors) into the staging buffers on the copy stream."""
        cs = self._copy_stream
        cs.wait_event(self._stage_free)          # the previous batch has left the staging buffers
        with torch.cuda.stream(cs):
            self._stage_img.copy_(image, non_blocking=True)
            self._stage_lab.copy_(label, non_blocking=True)
            self._stage_ready.record(cs)

    def run_staged(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_ready)
        self.static_image.copy_(self._stage_img, non_blocking=True)
        self.static_label.copy_(self._stage_lab, non_blocking=True)
        self._stage_free.record(cur)
        return self._replay()

    def _replay(self):
        self.opt._sync_lr()
        self.graph.replay()
        if self.world > 1 and not self.comm_in_graph:
            self._exchange_outside()
        return self.static_loss

    def __call__(self, image, label):
        if image.is_cuda and label.is_cuda:
            self.static_image.copy_(image, non_blocking=True)
            self.static_label.copy_(label, non_blocking=True)
            return self._replay()
        self.stage(image, label)
        return self.run_staged()


class GraphedInference:
    """Eval-mode forward of ``model`` captured ONCE into a CUDA graph for the shape of ``example`` and replayed per
    call: the ~140 kernel launches of a tile cost one graph launch on the host.  Use as a network in
    ``evaluate.predict_sliding(..., net_list=[GraphedInference(model, tile)], ...)``: ``net(img, task_id) -> logits``.
    The returned tensor is a static buffer that the next call overwrites (the sliding-window blend consumes it first,
    in stream order).  Inputs of any other shape fall back to the eager module.  The weights are treated as frozen:
    their standardised copies are computed once here, not per tile (``ops.frozen_weights``) -- build a new
    GraphedInference after the weights change."""

    def __init__(self, model: torch.nn.Module, example: torch.Tensor, warmup: int = 2):
        self.model = model
        self.static_in = example.detach().clone()
        was_training = model.training
        model.eval()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad(), ops.frozen_weights():
            for _ in range(max(warmup, 1)):
                model(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad(), ops.frozen_weights():
            out = model(self.static_in)
            self.static_out = out[0] if isinstance(out, (tuple, list)) else out
        if was_training:
            model.train()

    def __call__(self, img, task_id=None):
        if tuple(img.shape) != tuple(self.static_in.shape) or self.model.training:
            with torch.no_grad():
                out = self.model(img)
            return out[0] if isinstance(out, (tuple, list)) else out
        self.static_in.copy_(img, non_blocking=True)
        self.graph.replay()
        return self.static_out


def plan_tile_batches(run: int, lanes: int = 2, tile_batch: Optional[int] = None, cap: int = 8):
    """(tiles per forward, streams) for a rank that owns ``run`` tiles of a volume.  ``tile_batch`` given: clamped to the
    run.  Otherwise: as few batches of at most ``cap`` tiles as cover the run, their number rounded up to a multiple of
    ``lanes`` so that every stream gets the same number of batches, all of (nearly) the same size -- 96 tiles: 12 x 8; a
    rank's 12 tiles of an 8-GPU job: 2 x 6, one batch per stream; 36 tiles: 6 x 6.  Never more streams than batches."""
    run, lanes = max(int(run), 1), max(int(lanes), 1)
    if tile_batch is None:
        nb = (run + cap - 1) // cap
        nb = min((nb + lanes - 1) // lanes * lanes, run)
        tile_batch = (run + nb - 1) // nb
    tile_batch = max(min(int(tile_batch), run), 1)
    return tile_batch, max(min(lanes, (run + tile_batch - 1) // tile_batch), 1)


class GraphedSlidingWindow:
    """One sliding-window step -- eval-mode forward of a batch of ``tile_batch`` tiles of the volume, classifier, Gaussian
    weighting and accumulation into the volume accumulator -- captured ONCE into CUDA graphs and replayed per tile batch;
    the tile origins are device int32 rows the graphs read, so the same graphs serve all 96 tiles of a 300x512x512 volume.
    Tiles of a batch are accumulated in order, one ``mmpl_cls_blend`` launch each, so the accumulator sees the reference's
    tile order (evaluate_amos.py:228-276) whatever the batch size; batching only gives the low-resolution levels of the
    network (1 152 / 9 216 voxels per tile: fewer work items than SMs) more rows per launch.  One more graph pair serves
    the remainder of a rank's run in a single replay, and a single-tile pair ``blend_tile``.

    ``lanes`` = 2 (default) alternates the tile batches between two streams, each with its own graphs and buffers: the
    bandwidth-bound kernels of one batch (GroupNorm+ReLU, up-sampling) run under the tensor-bound convolutions of the other.
    A step is therefore two graphs -- the network up to the classifier's input, and the accumulation launches -- and the
    accumulation of batch k waits for the accumulation of batch k-1 on the other stream: the order in which overlapping
    tiles are added stays the reference's.

    Owns the fp32 accumulator ``acc`` [1, Dpad, C, H, W] (depth-major; Dpad = D rounded up to a multiple of
    ``world_size`` so that every rank owns a depth slab of the same size).  Used by ``evaluate.predict_sliding_dice``.  The weights are
    treated as frozen (``ops.frozen_weights``): build a new object after they change."""

    class _Captured:
        __slots__ = ("graph_f", "graph_b", "static_in", "origin_dev", "sink", "feat", "launches", "done")

    def __init__(self, model: torch.nn.Module, volume_dhw, tile, classes: int, world_size: int = 1, warmup: int = 2,
                 device=None, tile_batch: int = None, lanes: int = None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.model, self.tile, self.classes = model, tuple(int(t) for t in tile), int(classes)
        self.volume_dhw = tuple(int(v) for v in volume_dhw)
        D, H, W = self.volume_dhw
        self.world = max(int(world_size), 1)
        self.dpad = (D + self.world - 1) // self.world * self.world
        from .evaluate import tile_origins

        # tiles per forward: as given, else MMPL_SW_TILE_BATCH, else up to 8 (beyond that the low-resolution levels are
        # saturated: 370 -> 447 -> 490 -> 511 tiles/s for 1 / 2 / 4 / 8 on one B200, one lane).  A rank's run of ``run``
        # tiles is served by full batches, then ONE replay of a remainder graph (run % tile_batch tiles).
        ntiles = len(tile_origins((1, 1) + self.volume_dhw, self.tile))
        self.run = (ntiles + self.world - 1) // self.world
        if lanes is None:
            lanes = int(os.environ.get("MMPL_SW_LANES", "0")) or 2
        if tile_batch is None:
            tile_batch = int(os.environ.get("MMPL_SW_TILE_BATCH", "0")) or None
        self.tile_batch, self.lanes = plan_tile_batches(self.run, lanes, tile_batch)
        self.acc = torch.zeros((1, self.dpad, classes, H, W), dtype=torch.float32, device=dev)
        self._dev, self._warmup = dev, max(int(warmup), 1)
        was_training = model.training
        model.eval()
        if not model.blend_supported():
            raise RuntimeError("GraphedSlidingWindow needs the bf16 compute dtype and a 32/64-channel classifier")
        sizes = sorted({self.tile_batch, max(self.run % self.tile_batch, 1)})
        self._graphs = [{t: self._capture(t) for t in sizes} for _ in range(self.lanes)]
        if 1 not in self._graphs[0]:
            self._graphs[0][1] = self._capture(1)
        self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.lanes)] if self.lanes > 1 else []
        one = self._graphs[0][1]
        self.static_in, self.origin_dev, self.sink = one.static_in, one.origin_dev, one.sink
        self.launches_per_tile = one.launches          # library kernels one single-tile replay executes
        self.tiles_replayed = 0
        self.launches_replayed = 0                     # library kernels executed by all replays so far
        self.acc.zero_()
        if was_training:
            model.train()

    def _capture(self, tiles: int):
        from .evaluate import _gaussian_device

        c = GraphedSlidingWindow._Captured()
        c.origin_dev = torch.zeros((tiles, 3), dtype=torch.int32, device=self._dev)
        c.static_in = torch.zeros((tiles, 1) + self.tile, dtype=torch.float32, device=self._dev)
        c.sink = ops.BlendSink(self.acc, _gaussian_device(self.tile, self._dev), c.origin_dev, self.tile, d_outer=True)
        c.done = torch.cuda.Event()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad(), ops.frozen_weights():
            for _ in range(self._warmup):
                self.model.blend_accumulate(self.model.blend_features(c.static_in), c.sink)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        c.graph_f, c.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        launches0 = _lib.launch_count()
        with torch.cuda.graph(c.graph_f), torch.no_grad(), ops.frozen_weights():
            c.feat = self.model.blend_features(c.static_in)
        with torch.cuda.graph(c.graph_b), torch.no_grad():
            self.model.blend_accumulate(c.feat, c.sink)
        c.launches = _lib.launch_count() - launches0
        return c

    def close(self):
        """Release the captured graphs (before the process group goes away, like GraphedTrainStep.close)."""
        self._graphs = []

    def reset(self, lo=0, hi=None):
        """Zero the accumulator planes [lo, hi) (default: all) before a volume."""
        self.acc[0, lo:hi].zero_()

    def _fill(self, c, imgs, origins_dev):
        for j, im in enumerate(imgs):
            c.static_in[j:j + 1].copy_(im, non_blocking=True)
        c.origin_dev.copy_(origins_dev.reshape(len(imgs), 3), non_blocking=True)

    def _count(self, c, tiles):
        self.tiles_replayed += tiles
        self.launches_replayed += c.launches

    def blend_tile(self, img, origin_dev_row):
        """``img`` [1,1,td,th,tw] on the device; ``origin_dev_row`` a device int32[3] = (d0, h0, w0) of the tile."""
        c = self._graphs[0][1]
        self._fill(c, [img], origin_dev_row)
        c.graph_f.replay()
        c.graph_b.replay()
        self._count(c, 1)

    def blend_tiles(self, imgs, origins_dev):
        """``imgs``: a list of T device views [1,1,td,th,tw]; ``origins_dev`` device int32[T,3].  Full batches, then the
        remainder, in list order; the batches alternate between the lanes."""
        T, i, k = len(imgs), 0, 0
        cur = torch.cuda.current_stream()
        if self.lanes > 1:
            for s in self._streams:            # the tiles, the origins and the zeroed accumulator come from ``cur``
                s.wait_stream(cur)
        prev_done = None
        while i < T:
            tb = min(self.tile_batch, T - i)
            lane = k % self.lanes
            if tb not in self._graphs[lane]:       # a run length this object was not built for: single tiles
                tb, lane = 1, 0
            c = self._graphs[lane][tb]
            if self.lanes > 1:
                s = self._streams[lane]
                with torch.cuda.stream(s):
                    self._fill(c, imgs[i:i + tb], origins_dev[i:i + tb])
                    c.graph_f.replay()
                    if prev_done is not None:
                        s.wait_event(prev_done)    # accumulate after the previous batch (reference tile order)
                    c.graph_b.replay()
                    c.done.record(s)
                prev_done = c.done
            else:
                self._fill(c, imgs[i:i + tb], origins_dev[i:i + tb])
                c.graph_f.replay()
                c.graph_b.replay()
            self._count(c, tb)
            i += tb
            k += 1
        for s in self._streams:
            cur.wait_stream(s)


def extant_file(x):
    if not os.path.exists(x):
        raise argparse.ArgumentTypeError("{0} does not exist".format(x))
    return x


class Engine(object):
    """Same surface as the reference Engine (engine.py:10-77): context manager, ``.args``, ``.distributed``,
    ``.local_rank``, ``.world_size``, ``.devices``, ``data_parallel``, ``get_train_loader``, ``get_test_loader``,
    ``all_reduce_tensor``."""

    def __init__(self, custom_parser=None):
        self.devices = None
        if custom_parser is None:
            self.parser = argparse.ArgumentParser()
        else:
            assert isinstance(custom_parser, argparse.ArgumentParser)
            self.parser = custom_parser
        self.inject_default_parser()
        self.args = self.parser.parse_args()
        self.continue_state_object = self.args.continue_fpath
        self.world_size = _env_int("WORLD_SIZE", 1)
        self.local_rank = _env_int("LOCAL_RANK", 0)
        self.rank = _env_int("RANK", 0)
        self.distributed = self.world_size > 1
        self.devices = [self.local_rank]
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank)
        if self.distributed and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {"device_id": torch.device("cuda", self.local_rank)} if backend == "nccl" else {}
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size, **kw)

    def data_parallel(self, model):
        return DataParallelModel(model, self.world_size)

    def _loader(self, dataset, batch_size, shuffle, collate_fn=None):
        sampler = None
        if self.distributed:
            sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=self.world_size,
                                                                      rank=self.rank, shuffle=shuffle)
        loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size,
                                             num_workers=getattr(self.args, "num_workers", 0), drop_last=False,
                                             shuffle=(shuffle and sampler is None), pin_memory=torch.cuda.is_available(),
                                             sampler=sampler, collate_fn=collate_fn)
        return loader, sampler

    def get_train_loader(self, train_dataset, collate_fn=None):
        return self._loader(train_dataset, getattr(self.args, "batch_size", 1), True, collate_fn)

    def get_test_loader(self, test_dataset):
        return self._loader(test_dataset, 1, False)

    def all_reduce_tensor(self, tensor, norm=True):
        """Mean over ranks of the (scalar) tensor; single process: ``torch.mean`` like the reference stub (:57-58)."""
        if not self.distributed:
            return torch.mean(tensor)
        t = tensor.detach().clone().float()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if norm:
            t = t / self.world_size
        return torch.mean(t)

    def inject_default_parser(self):
        p = self.parser
        p.add_argument('-d', '--devices', default='', help='set data parallel training')
        p.add_argument('-c', '--continue', type=extant_file, metavar="FILE", dest="continue_fpath",
                       help='continue from one certain checkpoint')

    def __enter__(self):
        return self

    def __exit__(self, type, value, tb):
        if dist.is_initialized():
            try:
                dist.barrier()
            except Exception:
                pass
        if type is not None:
            print("A exception occurred during Engine initialization, give up running process")
            return False
