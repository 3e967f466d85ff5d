"""Autograd glue between PyTorch tensors and the C-ABI kernels.

Activations travel as logical NCDHW tensors in ``torch.channels_last_3d`` memory format, i.e. physically NDHWC --
the layout the kernels want -- so the reference's module interfaces (NCDHW in, NCDHW out) are kept without any
transpose kernels.  Gradients use the same convention.  Every function enqueues on torch's current stream.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import p as _p

_CL = torch.channels_last_3d

_cfg = {"dtype": torch.bfloat16, "conv_algo": os.environ.get("MMPL_CONV_ALGO", "auto"),
        "fuse_gn_bwd": os.environ.get("MMPL_FUSE_GN_BWD", "1") != "0",
        "ws_bwd_side_stream": os.environ.get("MMPL_WS_BWD_SIDE_STREAM", "1") != "0",
        "wgrad_side_stream": os.environ.get("MMPL_WGRAD_SIDE_STREAM", "1") != "0",
        "ws_fwd_side_stream": os.environ.get("MMPL_WS_FWD_SIDE_STREAM", "1") != "0",
        "fuse_gn_bwd_cls": {"0": False, "1": True}.get(os.environ.get("MMPL_FUSE_GN_BWD_CLS", ""), None)}


def set_fuse_gn_bwd(on: bool):
    """Fold the reduction pass of every GroupNorm+ReLU backward into the epilogue of the kernel that produces its dY
    (conv dgrad / classifier backward).  On by default; off = separate reduction kernel (used by the parity tests)."""
    _cfg["fuse_gn_bwd"] = bool(on)


def set_compute_dtype(dt: torch.dtype):
    """torch.bfloat16 (tcgen05 path, default) or torch.float32 (exact CUDA-core path for argmax/Dice parity)."""
    _lib.dtype_code(dt)
    _cfg["dtype"] = dt


def get_compute_dtype() -> torch.dtype:
    return _cfg["dtype"]


def set_conv_algo(algo: str):
    """'auto' (tcgen05 where supported), 'direct' (CUDA cores everywhere) or 'tcgen05' (fail if unsupported)."""
    assert algo in ("auto", "direct", "tcgen05")
    _cfg["conv_algo"] = algo


_PROF = {"on": False, "events": []}


def enable_conv_profile(on: bool):
    """bench.py: bracket every tcgen05 conv launch (fprop / dgrad / wgrad) with CUDA events on the launching stream."""
    _PROF["on"] = bool(on)
    _PROF["events"] = []
    return _PROF


def collect_conv_profile():
    """Totals over every tcgen05 conv launch, plus the same per kernel key (op, Cin, Cout, k, stride, voxels); the
    key with the largest total time is the dominant kernel of the step."""
    torch.cuda.synchronize()
    per = {}
    tot_ms, tot_fl = 0.0, 0.0
    for a, b, f, key in _PROF["events"]:
        t = a.elapsed_time(b)
        tot_ms += t
        tot_fl += f
        e = per.setdefault(key, [0.0, 0.0, 0])
        e[0] += t
        e[1] += f
        e[2] += 1
    # the dominant KERNEL is one template instantiation: fprop / fprop+res / dgrad / dgrad+gn launches of the same
    # (op, Cin, Cout, k, stride, voxels) problem are the same conv_tc kernel -> group by the first six key fields
    inst = {}
    for k, v in per.items():
        e = inst.setdefault(tuple(k[:6]), [0.0, 0.0, 0])
        e[0] += v[0]
        e[1] += v[1]
        e[2] += v[2]
    dom = max(inst.items(), key=lambda kv: kv[1][0]) if inst else (None, [0.0, 0.0, 0])
    # roofline classes: the 3x3x3 convolutions with >= 32 input channels are tensor bound; the Cin = 1 stem
    # and the 1x1x1 convolutions (<= 43 FLOP per byte at 64 -> 128 channels) sit far below the ridge: HBM / issue bound
    k3 = [v for k, v in per.items() if k[3] == 3 and not str(k[0]).startswith("stem")]
    return {"ms": tot_ms, "launches": len(_PROF["events"]), "flops": float(tot_fl),
            "k3": {"ms": sum(v[0] for v in k3), "flops": float(sum(v[1] for v in k3)), "launches": sum(v[2] for v in k3)},
            "per_key": {str(k): {"ms": v[0], "flops": float(v[1]), "launches": v[2]} for k, v in per.items()},
            "dominant": {"key": dom[0], "ms": dom[1][0], "flops": float(dom[1][1]), "launches": dom[1][2]}}


class _timed:
    """Context manager recording an event pair around a tcgen05 launch when profiling is on."""

    def __init__(self, algo, flops=0, key=None):
        self.on = _PROF["on"] and algo != _lib.ALGO_DIRECT
        self.flops = flops
        self.key = key

    def __enter__(self):
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            _PROF["events"].append((self.a, self.b, self.flops, self.key))
        return False


def _grad_dst(param, shape):
    """Where a backward kernel writes a parameter gradient.  If the parameter has a slot in a flat gradient buffer
    (engine.DataParallelModel / FusedSGD set ``_mmpl_grad_slot``) and no gradient yet this step, the kernel writes
    straight into a fresh view of that slot and autograd's AccumulateGrad adopts the view as ``param.grad`` -- no copy,
    no add kernel.  Otherwise (gradient accumulation, parameters without a slot, non-fp32 masters) a new tensor."""
    if _grad_is_direct(param):
        flat, off, n = param._mmpl_grad_slot
        return flat[off:off + n].view(shape)
    return torch.empty(shape, dtype=torch.float32, device=param.device if param is not None else None)


def _grad_is_direct(param) -> bool:
    """True if ``_grad_dst`` hands out the flat-buffer view, i.e. autograd will adopt the gradient without launching
    anything that reads it."""
    slot = getattr(param, "_mmpl_grad_slot", None) if param is not None else None
    return slot is not None and param.grad is None and param.dtype == torch.float32


# ---- side stream for the weight-standardisation backward --------------------------------------------------------
# dW of a convolution has no consumer inside the backward pass (only the optimizer / gradient all-reduce read it), but
# its two small kernels (wgrad -> ws_weight_bwd) would sit on the critical path 35 times per step.  ws_weight_bwd is
# therefore launched on a side stream that forks after the wgrad kernel and is joined once, when the backward pass
# ends (autograd engine callback) -- inside a CUDA-graph capture this becomes a parallel branch of the graph.
_SIDE = {"stream": None, "pending": False, "keep": []}


def _side_stream():
    if _SIDE["stream"] is None:
        _SIDE["stream"] = torch.cuda.Stream()
    return _SIDE["stream"]


def join_side_stream():
    """Make the current stream wait for everything launched on the side stream (no-op if nothing is pending)."""
    if _SIDE["pending"]:
        torch.cuda.current_stream().wait_stream(_SIDE["stream"])
        _SIDE["pending"] = False
    _SIDE["keep"].clear()      # tensors the side stream was still reading may be recycled from here on


def _ws_bwd_launch(g_hat, w_hat, inv_std, cout, cin, taps, standardise, dw, off_critical_path):
    L = _lib.lib()
    if not (off_critical_path and _cfg["ws_bwd_side_stream"]):
        _lib.check(L.mmpl_ws_weight_bwd(_p(g_hat), _p(w_hat), _p(inv_std), cout, cin, taps, standardise, _p(dw),
                                        _lib.stream_ptr()), "ws_weight_bwd")
        return
    cur, side = torch.cuda.current_stream(), _side_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        _lib.check(L.mmpl_ws_weight_bwd(_p(g_hat), _p(w_hat), _p(inv_std), cout, cin, taps, standardise, _p(dw),
                                        _lib.stream_ptr()), "ws_weight_bwd")
    g_hat.record_stream(side)
    _side_mark_pending()


def _side_mark_pending():
    if not _SIDE["pending"]:
        _SIDE["pending"] = True
        torch.autograd.Variable._execution_engine.queue_callback(join_side_stream)


_STATS_POOL = {"buf": None, "off": 0}


# dX tensors whose producer (conv dgrad / classifier backward) already accumulated the GroupNorm-backward sums:
# data_ptr -> (workspace data_ptr, head).  Consumed (popped) by the GNReLUFn.backward that autograd runs next and cleared
# by every begin_forward.  A hand-off that is lost or does not match (autograd summed two gradients into a new tensor, a
# second model ran a forward in between) only costs the fusion: GNReLUFn.backward then zeroes the workspace and runs its
# own reduction pass (mmpl_gn_relu_bwd, reduced = 0), so interleaving several networks in one loop is safe.
_GN_REDUCED = {}


def begin_forward(device, doubles: int = 1 << 16):
    """One zero-fill for all GroupNorm accumulators of a forward pass (statistics [N][16][2] and backward workspaces
    [N][C][4] per layer) instead of one per layer.  A fresh pool per forward: the slices live until the backward."""
    # sized by what the previous forward asked for (the wide cfg5 network needs more than the default)
    doubles = max(doubles, _STATS_POOL.get("demand", 0))
    _STATS_POOL["buf"] = torch.zeros(doubles, dtype=torch.float64, device=device)
    _STATS_POOL["off"] = 0
    _STATS_POOL["demand"] = 0
    _GN_REDUCED.clear()


def _zero_stats(numel: int, device) -> torch.Tensor:
    _STATS_POOL["demand"] = _STATS_POOL.get("demand", 0) + numel
    buf, off = _STATS_POOL["buf"], _STATS_POOL["off"]
    if buf is not None and buf.device == device and off + numel <= buf.numel():
        _STATS_POOL["off"] = off + numel
        return buf[off:off + numel]
    return torch.zeros(numel, dtype=torch.float64, device=device)


class _StatsBox:
    """The GroupNorm raw sums a producer's epilogue accumulated, handed out of an autograd Function as a NON-tensor output.
    As a tensor output (even marked non-differentiable) autograd materialises a zero gradient for it in every backward:
    one 64-element fill launch per convolution per step, 32 nodes on the critical path of the captured step."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def to_cl(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """Logical NCDHW tensor -> compute dtype with physically dense NDHWC storage (no copy if already so)."""
    dtype = dtype or _cfg["dtype"]
    if x.dtype != dtype:
        x = x.to(dtype)
    v = x.permute(0, 2, 3, 4, 1)
    if not v.is_contiguous():
        x = v.contiguous().permute(0, 4, 1, 2, 3)
    return x


def empty_cl(n, c, d, h, w, dtype, device) -> torch.Tensor:
    return torch.empty((n, d, h, w, c), dtype=dtype, device=device).permute(0, 4, 1, 2, 3)


def _tc_supported(dtype, kred, nout) -> bool:
    """tcgen05 conv kernels: bf16, reduction channels 32 or a multiple of 64, output channels 32/64/128/256 (or a
    multiple of 256); 32 reduction channels only with 32 or 64 outputs."""
    if dtype != torch.bfloat16:
        return False
    ok_in = kred == 32 or kred % 64 == 0
    ok_out = nout in (32, 64, 128, 256) or (nout > 256 and nout % 256 == 0)
    if kred == 32 and nout not in (32, 64):
        return False
    return ok_in and ok_out


_FALLBACK_WARNED = set()


def _algo(dtype, kred, nout) -> int:
    mode = _cfg["conv_algo"]
    if mode == "direct":
        return _lib.ALGO_DIRECT
    if _tc_supported(dtype, kred, nout):
        return _lib.ALGO_TCGEN05
    if mode == "tcgen05":
        raise RuntimeError(f"tcgen05 conv path does not cover dtype={dtype} {kred}->{nout}")
    if dtype == torch.bfloat16 and (kred, nout) not in _FALLBACK_WARNED:
        # a 10-100x performance cliff must not be silent (e.g. the refiner's 24/48/96-channel layers)
        _FALLBACK_WARNED.add((kred, nout))
        import warnings

        warnings.warn(f"multimodal-pl_b200: no tcgen05 kernel for a bf16 {kred}->{nout} channel convolution (reduction "
                      "channels 32 or a multiple of 64, outputs 32/64/128/256 or a multiple of 256); it runs on the "
                      "CUDA-core direct kernel, 10-100x slower", RuntimeWarning, stacklevel=3)
    return _lib.ALGO_DIRECT


def _out_dim(i, k, s):
    return (i + 2 * (k // 2) - k) // s + 1


# --------------------------------------------------------------------------------------------------------------
# Standardised weights.  The reference recomputes w_hat in every Conv3d.forward (unet3D.py:22-26), 8 ATen launches per
# convolution.  Here each weight owns persistent buffers (w_hat, inv_std, the two tap-major packings) and ALL
# convolutions of a network are refreshed by ONE launch (mmpl_ws_weight_fwd_batched) at the top of the model's forward
# (``prepare_ws``).  The refresh is unconditional -- weights may be changed behind autograd's back (``p.data`` updates,
# raw-pointer optimizers, CUDA-graph replays), so nothing is ever assumed fresh across forwards: ``prepare_ws`` marks
# each entry as ready for exactly one use, and a convolution that finds its entry not ready (module called on its own,
# or a weight shared by two layers) refreshes it itself.
_WS_TABLES = {}


class _WsEntry:
    __slots__ = ("w_hat", "inv_std", "pf", "pd", "key", "shape_key", "on_side", "stamp")


# The per-weight buffers are rewritten in place by the next refresh, behind autograd's version counters.  Each refresh
# stamps the entry with (optimizer epoch, weight._version); a convolution remembers the stamp it ran with and its backward
# refuses to run on buffers that were meanwhile refreshed from CHANGED weights (stock PyTorch raises "modified by an
# inplace operation" in that situation).  A refresh from unchanged weights rewrites the same values and is harmless.
_WEIGHTS_EPOCH = [0]


def bump_weights_epoch():
    """Called by optimizers that update weights through raw pointers (engine.FusedSGD)."""
    _WEIGHTS_EPOCH[0] += 1


def _ws_stamp(weight):
    return (_WEIGHTS_EPOCH[0], weight._version)


def _ws_key(weight, dt, standardise, stem_kch):
    return (weight.data_ptr(), dt, bool(standardise), stem_kch)


def _ws_entry(weight, dt, stem_kch=0, packed=True):
    shape_key = (tuple(weight.shape), dt, weight.device, stem_kch, packed)
    e = getattr(weight, "_mmpl_ws", None)
    if e is None or e.shape_key != shape_key:
        cout, cin = weight.shape[0], weight.shape[1]
        taps = weight[0, 0].numel()
        dev = weight.device
        e = _WsEntry()
        e.w_hat = torch.empty(weight.shape, dtype=torch.float32, device=dev)
        e.inv_std = torch.empty(cout, dtype=torch.float32, device=dev)
        if stem_kch:
            e.pf = torch.zeros((cout, stem_kch), dtype=dt, device=dev)
            e.pd = None
        elif packed:
            e.pf = torch.empty(taps * cout * cin, dtype=dt, device=dev)
            e.pd = torch.empty(taps * cout * cin, dtype=dt, device=dev)
        else:
            e.pf = e.pd = None
        e.key = None
        e.on_side = False
        e.stamp = None
        e.shape_key = shape_key
        weight._mmpl_ws = e
    return e


def _ws_cacheable(weight):
    return weight.dtype == torch.float32 and weight.is_contiguous() and weight.is_cuda


def _ws_refresh_one(weight, e, dt, standardise, stem_kch):
    """Single-convolution refresh (fallback when prepare_ws was not called, e.g. a module used on its own)."""
    L = _lib.lib()
    cout, cin = weight.shape[0], weight.shape[1]
    taps = weight[0, 0].numel()
    w32 = weight.detach()
    st = _lib.stream_ptr()
    if stem_kch:
        tab = _ws_table([(weight, standardise, stem_kch, e)], dt)
        _lib.check(L.mmpl_ws_weight_fwd_batched(_p(tab[0]), tab[1], tab[2], _lib.dtype_code(dt), st), "ws_weight_fwd")
    else:
        _lib.check(L.mmpl_ws_weight_fwd(_p(w32), cout, cin, taps, int(standardise), _p(e.w_hat), _p(e.inv_std), _p(e.pf),
                                        _p(e.pd), _lib.dtype_code(dt), st), "ws_weight_fwd")


def _ws_table(items, dt):
    """Device table of mmpl_ws_entry records (include/mmpl_b200.h) for ``items`` = [(weight, standardise, stem_kch,
    entry)], cached by the pointers it contains."""
    import numpy as np

    rec = np.dtype([("w", "<u8"), ("w_hat", "<u8"), ("inv_std", "<u8"), ("pf", "<u8"), ("pd", "<u8"), ("cout", "<i4"),
                    ("cin", "<i4"), ("taps", "<i4"), ("standardise", "<i4"), ("first_block", "<i4"), ("stem_kch", "<i4")])
    assert rec.itemsize == 64
    arr = np.zeros(len(items), dtype=rec)
    first = 0
    for i, (weight, standardise, stem_kch, e) in enumerate(items):
        cout, cin = weight.shape[0], weight.shape[1]
        arr[i] = (weight.data_ptr(), e.w_hat.data_ptr(), e.inv_std.data_ptr(), 0 if e.pf is None else e.pf.data_ptr(),
                  0 if e.pd is None else e.pd.data_ptr(), cout, cin, weight[0, 0].numel(), int(standardise), first,
                  stem_kch)
        first += (cout + 7) // 8          # a block of the batched kernel serves 8 out-channels
    key = arr.tobytes()
    hit = _WS_TABLES.get(key)
    if hit is None:
        if len(_WS_TABLES) > 16:
            _WS_TABLES.clear()
        dev = items[0][0].device
        t = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        hit = (t, len(items), first)
        _WS_TABLES[key] = hit
    return hit


def _stem_plan(dt, cout):
    """-> (use_tc, tc_fwd, kch) for the Cin = 1 stem under the current dtype / algo / stem mode."""
    mode = _STEM["mode"]
    kch = 64 if mode in ("split", "fused") else 32
    use_tc = (mode != "direct" and _algo(dt, kch, cout) == _lib.ALGO_TCGEN05
              and _tc_wgrad_supported(dt, 1, 1, kch, cout))
    return use_tc, use_tc and mode != "fp32fwd", kch


class frozen_weights:
    """Context manager for inference with weights that do not change: inside it the standardised weights are computed
    by the first forward and reused by the following ones (the reference recomputes them in every Conv3d.forward,
    unet3D.py:22-26; with frozen weights that is the same values 96 times per volume).  Leave the context -- or call
    ``ops.invalidate_weights()`` -- before the weights change."""

    def __enter__(self):
        self.prev = _cfg.get("ws_frozen", False)
        _cfg["ws_frozen"] = True
        return self

    def __exit__(self, *exc):
        _cfg["ws_frozen"] = self.prev
        return False


def prepare_ws(convs):
    """Refresh the standardised weights of ``convs`` = [(weight, standardise, is_stem)] with one launch.  Called at
    the top of unet3D_baseline.forward."""
    dt = _cfg["dtype"]
    items, keys = [], []
    for weight, standardise, is_stem in convs:
        if not _ws_cacheable(weight):
            continue
        if is_stem:
            _, tc_fwd, kch = _stem_plan(dt, weight.shape[0])
            stem_kch = kch if tc_fwd else 0
            e = _ws_entry(weight, dt, stem_kch, packed=False)
        else:
            stem_kch = 0
            e = _ws_entry(weight, dt)
        items.append((weight, standardise, stem_kch, e))
        keys.append(_ws_key(weight, dt, standardise, stem_kch))
    if not items:
        return
    if _cfg.get("ws_frozen", False) and all(it[3].key == k for it, k in zip(items, keys)):
        return                       # frozen weights: what the previous forward computed is still valid
    _lib.require_device()
    L = _lib.lib()
    code = _lib.dtype_code(dt)
    first = [it for it in items if it[2] or it[0].shape[1] == 1]        # the stem: needed immediately
    rest = [it for it in items if not (it[2] or it[0].shape[1] == 1)]
    if _cfg["ws_fwd_side_stream"] and first and rest:
        # the stem's weights on the current stream; everything else on a side stream, concurrently with the stem
        # (im2col + 1x1x1 conv are HBM-bound).  The first convolution that needs them joins (``_ws_get``).
        tab = _ws_table(first, dt)
        _lib.check(L.mmpl_ws_weight_fwd_batched(_p(tab[0]), tab[1], tab[2], code, _lib.stream_ptr()),
                   "ws_weight_fwd_batched")
        if _WS_SIDE["stream"] is None:
            _WS_SIDE["stream"] = torch.cuda.Stream()
        side = _WS_SIDE["stream"]
        side.wait_stream(torch.cuda.current_stream())
        tab = _ws_table(rest, dt)
        with torch.cuda.stream(side):
            _lib.check(L.mmpl_ws_weight_fwd_batched(_p(tab[0]), tab[1], tab[2], code, _lib.stream_ptr()),
                       "ws_weight_fwd_batched")
        _WS_SIDE["pending"] = True
        side_ids = {id(it[3]) for it in rest}
    else:
        tab = _ws_table(items, dt)
        _lib.check(L.mmpl_ws_weight_fwd_batched(_p(tab[0]), tab[1], tab[2], code, _lib.stream_ptr()),
                   "ws_weight_fwd_batched")
        side_ids = set()
    for (w_, _, _, e), key in zip(items, keys):
        e.key = key
        e.on_side = id(e) in side_ids
        e.stamp = _ws_stamp(w_)


_WS_SIDE = {"stream": None, "pending": False}


def _ws_get(weight, dt, standardise, stem_kch=0, packed=True):
    """Standardised-weight buffers of ``weight`` for one use: those ``prepare_ws`` just refreshed, else refreshed here."""
    if _ws_cacheable(weight):
        e = _ws_entry(weight, dt, stem_kch, packed)
        if e.key != _ws_key(weight, dt, standardise, stem_kch):
            _ws_refresh_one(weight, e, dt, standardise, stem_kch)
            e.stamp = _ws_stamp(weight)
        elif e.on_side and _WS_SIDE["pending"]:
            torch.cuda.current_stream().wait_stream(_WS_SIDE["stream"])      # join the side-stream refresh once
            _WS_SIDE["pending"] = False
        if not _cfg.get("ws_frozen", False):
            e.key = None    # consumed: the next forward refreshes again
        e.on_side = False
        return e
    # non-fp32 / non-contiguous master weight: one-off buffers
    w32 = weight.detach().float().contiguous()
    e = _ws_entry(w32, dt, stem_kch, packed)
    _ws_refresh_one(w32, e, dt, standardise, stem_kch)
    e.stamp = _ws_stamp(w32)
    return e


# --------------------------------------------------------------------------------------------------------------
class WSConv3dFn(torch.autograd.Function):
    """Weight-standardised (or plain) k^3 convolution, optional fused residual add.  unet3D.py:16-27."""

    @staticmethod
    def forward(ctx, x, weight, residual, stride, standardise, want_stats=False):
        _lib.require_device()
        L = _lib.lib()
        dt = _cfg["dtype"]
        ctx.gn_bwd = getattr(x, "_mmpl_gn_bwd", None)      # x = relu(gn(.)): its backward reduction rides on our dgrad
        presplit = bool(getattr(x, "_mmpl_psplit", False))   # x's memory already is the parity-split tensor (gn_relu_dual)
        x = to_cl(x, dt)
        cout, cin, k = weight.shape[0], weight.shape[1], weight.shape[2]
        taps = k * k * k
        n, _, d, h, w = x.shape
        assert x.shape[1] == cin, f"conv: input has {x.shape[1]} channels, weight expects {cin}"
        dev = x.device
        code = _lib.dtype_code(dt)
        st = _lib.stream_ptr()
        ws = _ws_get(weight, dt, standardise)
        w_hat, inv_std, pf, pd = ws.w_hat, ws.inv_std, ws.pf, ws.pd
        do, ho, wo = _out_dim(d, k, stride), _out_dim(h, k, stride), _out_dim(w, k, stride)
        y = empty_cl(n, cout, do, ho, wo, dt, dev)
        res = None
        if residual is not None:
            res = to_cl(residual, dt)
            assert res.shape == y.shape
        algo = _algo(dt, cin, cout)
        src = x
        if presplit and not (algo == _lib.ALGO_TCGEN05 and stride == 2 and k == 3):
            raise RuntimeError("a parity-split activation (gn_relu_dual(psplit1=True)) can only feed the stride-2 3x3x3 "
                               "tensor-core convolution")
        if algo == _lib.ALGO_TCGEN05 and stride == 2 and k == 3:
            # stride-2 3x3x3 on tensor cores reads a parity-split copy of the input: written directly by the GroupNorm
            # kernel that produced x (presplit), else by one extra streaming pass
            if presplit:
                src = x.permute(0, 2, 3, 4, 1).reshape(8 * n, do, ho, wo, cin)
                assert src.data_ptr() == x.data_ptr()
            else:
                src = torch.empty((8 * n, do, ho, wo, cin), dtype=dt, device=dev)
                _lib.check(L.mmpl_parity_split(_p(x), _p(src), n, d, h, w, cin, code, st), "parity_split")
            algo = _lib.ALGO_TCGEN05_PSPLIT
        flops = 2 * n * do * ho * wo * cout * cin * taps
        # fprop and stride-1 dgrad of a Cin==Cout layer are the same kernel instantiation on the same problem size
        ctx_key = ("conv_tc", min(cin, cout), max(cin, cout), k, stride, n * d * h * w)
        # GroupNorm(16) raw sums of the output for the next GN+ReLU, produced by the conv epilogue
        stats = _zero_stats(n * 16 * 2, dev) if (want_stats and cout % 16 == 0) else None
        with _timed(algo, flops, ctx_key + ("fprop+res" if res is not None else "fprop",)):
            _lib.check(L.mmpl_conv3d_fprop(_p(src), _p(pf), _p(res), _p(y), n, d, h, w, cin, cout, k, stride, code, algo,
                                           _p(stats), st), "conv3d_fprop")
        # stride-2 3x3x3 on tensor cores: the parity-split copy is what wgrad reads, so keep it instead of x
        keep = src if (algo == _lib.ALGO_TCGEN05_PSPLIT and _tc_wgrad_supported(dt, k, stride, cin, cout)) else x
        ctx.save_for_backward(keep, w_hat, inv_std, pd)
        ctx.ws_entry, ctx.ws_stamp = ws, ws.stamp
        ctx.x_is_psplit = keep is not x
        ctx.meta = (n, d, h, w, cin, cout, k, stride, int(standardise), residual is not None, weight.dtype)
        ctx.weight = weight
        ctx.flops = flops
        ctx.key = ctx_key
        if want_stats:
            return y, _StatsBox(stats)
        return y

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        L = _lib.lib()
        x, w_hat, inv_std, pd = ctx.saved_tensors
        _check_ws_stamp(ctx)
        n, d, h, w, cin, cout, k, stride, standardise, has_res, wdtype = ctx.meta
        dt = pd.dtype
        code = _lib.dtype_code(dt)
        st = _lib.stream_ptr()
        dy = to_cl(dy, dt)
        dev = x.device
        dx = dw = dres = None
        # The weight-gradient chain (wgrad -> ws_weight_bwd) has no consumer inside the backward pass.  When autograd
        # will simply adopt dW (flat-buffer slot), the chain runs on the side stream, forked HERE -- before the data
        # gradient is enqueued -- so the critical path (dgrad -> GN backward -> previous layer) does not wait for it and
        # its CTAs fill the SMs that tails and low-resolution kernels leave idle.
        direct = ctx.needs_input_grad[1] and _grad_is_direct(ctx.weight)
        wgrad_side = (direct and ctx.needs_input_grad[0] and _cfg["wgrad_side_stream"] and _cfg["ws_bwd_side_stream"]
                      and not _PROF["on"])
        if wgrad_side:
            _side_stream().wait_stream(torch.cuda.current_stream())
        if ctx.needs_input_grad[0]:
            dx = empty_cl(n, cin, d, h, w, dt, dev)
            algo = _algo(dt, cout, cin)
            fuse, fused = None, None
            if ctx.gn_bwd is not None and algo == _lib.ALGO_TCGEN05 and x.dtype == dt:
                # x (our forward input, or its parity-split copy) is a = relu(gn(.)) itself
                gb, gws, ghead = ctx.gn_bwd
                fuse = _lib.GnBwdFuse(_p(x), _p(gb), _p(gws), int(ctx.x_is_psplit), int(ghead))
                fused = ctypes.c_int(0)
            with _timed(algo, ctx.flops, ctx.key + ("dgrad+gn" if fuse is not None else "dgrad",)):
                _lib.check(L.mmpl_conv3d_dgrad(_p(dy), _p(pd), None, _p(dx), n, d, h, w, cin, cout, k, stride, code,
                                               algo, ctypes.byref(fuse) if fuse is not None else None,
                                               ctypes.byref(fused) if fused is not None else None, st), "conv3d_dgrad")
            if fused is not None and fused.value:
                _GN_REDUCED[dx.data_ptr()] = (gws.data_ptr(), ghead)
        if ctx.needs_input_grad[1]:
            taps = k * k * k
            algo = _lib.ALGO_DIRECT
            if ctx.x_is_psplit:
                algo = _lib.ALGO_TCGEN05_PSPLIT
            elif _cfg["conv_algo"] != "direct" and _tc_wgrad_supported(dt, k, stride, cin, cout) and not (
                    stride == 2 and k == 3):
                algo = _lib.ALGO_TCGEN05
            dw = _grad_dst(ctx.weight, w_hat.shape)

            def chain(on_side):
                g_hat = torch.empty(taps * cout * cin, dtype=torch.float32, device=dev)
                wsb = int(L.mmpl_conv3d_wgrad_workspace(n, d, h, w, cin, cout, k, stride, algo))
                ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev) if wsb else None
                with _timed(algo, ctx.flops, ("wgrad_tc",) + ctx.key[1:]):
                    _lib.check(L.mmpl_conv3d_wgrad(_p(x), _p(dy), _p(g_hat), n, d, h, w, cin, cout, k, stride, code, algo,
                                                   _p(ws), wsb, _lib.stream_ptr()), "conv3d_wgrad")
                if on_side:
                    _lib.check(L.mmpl_ws_weight_bwd(_p(g_hat), _p(w_hat), _p(inv_std), cout, cin, taps, standardise,
                                                    _p(dw), _lib.stream_ptr()), "ws_weight_bwd")
                    _SIDE["keep"].append((x, dy, g_hat, ws))     # alive until the join
                else:
                    _ws_bwd_launch(g_hat, w_hat, inv_std, cout, cin, taps, standardise, dw, direct)

            if wgrad_side:
                with torch.cuda.stream(_side_stream()):
                    chain(True)
                _side_mark_pending()
            else:
                chain(False)
            dw = dw.to(wdtype)
        if has_res and ctx.needs_input_grad[2]:
            dres = dy
        return dx, dw, dres, None, None, None


def _check_ws_stamp(ctx):
    if ctx.ws_entry.stamp != ctx.ws_stamp:
        raise RuntimeError("the weights of this convolution were changed and re-standardised between its forward and "
                           "its backward (optimizer step or in-place update followed by another forward): the saved "
                           "standardised weights no longer exist -- run backward before updating the weights")


_TC_WGRAD = {"enabled": os.environ.get("MMPL_TC_WGRAD", "1") != "0"}


def _tc_wgrad_supported(dtype, k, stride, cin, cout) -> bool:
    if not (_TC_WGRAD["enabled"] and dtype == torch.bfloat16 and _cfg["conv_algo"] != "direct"):
        return False
    if stride == 2 and k == 3 and cout % 64 != 0:
        return False
    if cin == 32:
        return cout == 32 or cout % 64 == 0
    return cin % 64 == 0 and (cout == 32 or cout % 64 == 0)


def psplit_consumer(cin, cout, k, stride) -> bool:
    """True if ``ws_conv3d`` with these parameters reads a parity-split input on the current path, i.e. its producer may
    write that layout directly (``gn_relu_dual(psplit1=True)``)."""
    dt = _cfg["dtype"]
    return (k == 3 and stride == 2 and _PSPLIT_DIRECT and _algo(dt, cin, cout) == _lib.ALGO_TCGEN05
            and _tc_wgrad_supported(dt, k, stride, cin, cout))


_PSPLIT_DIRECT = os.environ.get("MMPL_PSPLIT_DIRECT", "1") != "0"


def ws_conv3d(x, weight, stride=1, standardise=True, residual=None, want_stats=False):
    """want_stats=True additionally computes the GroupNorm(16) statistics of the output (fused into the tcgen05
    epilogue) and attaches them to the returned tensor, where ``gn_relu`` picks them up."""
    if not want_stats:
        return WSConv3dFn.apply(x, weight, residual, int(stride), bool(standardise), False)
    y, box = WSConv3dFn.apply(x, weight, residual, int(stride), bool(standardise), True)
    if box.t is not None:
        y._mmpl_gn_stats = (box.t, 16)
    return y


# --------------------------------------------------------------------------------------------------------------
_STEM = {"mode": os.environ.get("MMPL_STEM", "fused")}


def set_stem_mode(mode: str):
    """bf16 stem: 'fused' (default: tcgen05, the 27-tap hi+lo bf16 operand tile is built in shared memory from the fp32
    image, nothing is expanded in HBM), 'split' (tcgen05 on an expanded [N,D,H,W,64] hi+lo copy of the image), 'tc32'
    (expanded, image rounded to bf16, K = 32), 'fp32fwd' (CUDA-core fp32 forward, tcgen05 weight gradient) or 'direct'
    (CUDA cores only)."""
    assert mode in ("fused", "split", "tc32", "fp32fwd", "direct")
    _STEM["mode"] = mode


class StemConvFn(torch.autograd.Function):
    """conv3x3x3(1 -> base) on the fp32 image (unet3D.py:594, :666).

    bf16 / tcgen05 path: the image is expanded once into a bf16 [N,D,H,W,K] tensor holding its 27 shifted copies
    (mmpl_stem_im2col; K = 64 carries each fp32 value as hi + lo bf16 parts so the image is not rounded); forward and
    weight gradient are then K -> base 1x1x1 convolutions on the tensor-core kernels (HBM-bound) instead of 864
    CUDA-core FMAs per voxel.  fp32 / 'direct' path: CUDA-core kernels on the fp32 image."""

    @staticmethod
    def forward(ctx, image, weight, standardise):
        _lib.require_device()
        L = _lib.lib()
        dt = _cfg["dtype"]
        img = image.detach().float().contiguous()
        n, c, d, h, w = img.shape
        assert c == 1 and tuple(weight.shape[1:]) == (1, 3, 3, 3)
        cout = weight.shape[0]
        dev = img.device
        st = _lib.stream_ptr()
        y = empty_cl(n, cout, d, h, w, dt, dev)
        use_tc, tc_fwd, kch = _stem_plan(dt, cout)
        ws = _ws_get(weight, dt, standardise, kch if tc_fwd else 0, packed=False)
        w_hat, inv_std, pf = ws.w_hat, ws.inv_std, ws.pf
        fused = tc_fwd and _STEM["mode"] == "fused" and cout in (32, 64)
        if fused:
            # the K = 64 operand tile is built in shared memory inside the kernel (csrc/stem_tc.cu)
            stats = _zero_stats(n * 16 * 2, dev) if cout % 16 == 0 else None
            flops = 2 * n * d * h * w * cout * 27          # algorithmic (the tensor core multiplies K = 64: hi + lo + padding)
            with _timed(_lib.ALGO_TCGEN05, flops, ("stem_tc", 1, cout, 3, 1, n * d * h * w, "fprop")):
                _lib.check(L.mmpl_stem_tc_fwd(_p(img), _p(pf), _p(y), _p(stats), n, d, h, w, cout, st), "stem_tc_fwd")
            ctx.save_for_backward(img, w_hat, inv_std)
        elif tc_fwd:
            x27 = torch.empty((n, d, h, w, kch), dtype=dt, device=dev)
            _lib.check(L.mmpl_stem_im2col(_p(img), _p(x27), n, d, h, w, kch, st), "stem_im2col")
            code = _lib.dtype_code(dt)
            flops = 2 * n * d * h * w * cout * kch
            # GroupNorm(16) raw sums of the stem output (layer0.0.gn1) come from the conv epilogue
            stats = _zero_stats(n * 16 * 2, dev) if cout % 16 == 0 else None
            with _timed(_lib.ALGO_TCGEN05, flops, ("conv_tc", kch, cout, 1, 1, n * d * h * w)):
                _lib.check(L.mmpl_conv3d_fprop(_p(x27), _p(pf), None, _p(y), n, d, h, w, kch, cout, 1, 1, code,
                                               _lib.ALGO_TCGEN05, _p(stats), st), "conv3d_fprop(stem)")
            ctx.save_for_backward(x27, w_hat, inv_std)
        else:
            _lib.check(L.mmpl_stem_conv_fwd(_p(img), _p(w_hat), _p(y), n, d, h, w, cout, _lib.dtype_code(dt), st),
                       "stem_conv_fwd")
            ctx.save_for_backward(img, w_hat, inv_std)
        ctx.ws_entry, ctx.ws_stamp = ws, ws.stamp
        ctx.meta = (n, d, h, w, cout, int(standardise), weight.dtype, use_tc, tc_fwd, kch, fused)
        ctx.weight = weight
        return y, _StatsBox(stats if tc_fwd else None)

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        L = _lib.lib()
        src, w_hat, inv_std = ctx.saved_tensors
        _check_ws_stamp(ctx)
        n, d, h, w, cout, standardise, wdtype, use_tc, tc_fwd, kch, fused = ctx.meta
        dy = to_cl(dy)
        st = _lib.stream_ptr()
        dev = src.device
        if fused:
            g_hat = torch.empty(27 * cout, dtype=torch.float32, device=dev)             # tap-major [27][cout]
            flops = 2 * n * d * h * w * cout * 27
            with _timed(_lib.ALGO_TCGEN05, flops, ("stem_tc", 1, cout, 3, 1, n * d * h * w, "wgrad")):
                _lib.check(L.mmpl_stem_tc_wgrad(_p(src), _p(dy), _p(g_hat), n, d, h, w, cout, st), "stem_tc_wgrad")
        elif use_tc:
            if not tc_fwd:      # forward ran on the fp32 image: expand it now (bf16 rounding is fine for dW)
                x27 = torch.empty((n, d, h, w, kch), dtype=dy.dtype, device=dev)
                _lib.check(L.mmpl_stem_im2col(_p(src), _p(x27), n, d, h, w, kch, st), "stem_im2col")
                src = x27
            code = _lib.dtype_code(dy.dtype)
            g32 = torch.empty(cout * kch, dtype=torch.float32, device=dev)            # [1 tap][cout][kch]
            flops = 2 * n * d * h * w * cout * kch
            with _timed(_lib.ALGO_TCGEN05, flops, ("wgrad_tc", kch, cout, 1, 1, n * d * h * w)):
                _lib.check(L.mmpl_conv3d_wgrad(_p(src), _p(dy), _p(g32), n, d, h, w, kch, cout, 1, 1, code,
                                               _lib.ALGO_TCGEN05, None, 0, st), "conv3d_wgrad(stem)")
            g2 = g32.view(cout, kch)
            g_hat = g2[:, :27] + g2[:, 32:59] if kch == 64 else g2[:, :27]
            g_hat = g_hat.t().contiguous()                                              # tap-major [27][cout]
        else:
            g_hat = torch.empty(27 * cout, dtype=torch.float32, device=dev)
            wsb = int(L.mmpl_stem_conv_wgrad_workspace(n, d, h, w))
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            _lib.check(L.mmpl_stem_conv_wgrad(_p(src), _p(dy), _p(g_hat), n, d, h, w, cout, _lib.dtype_code(dy.dtype),
                                              _p(ws), wsb, st), "stem_conv_wgrad")
        direct = _grad_is_direct(ctx.weight)
        dw = _grad_dst(ctx.weight, w_hat.shape)
        _ws_bwd_launch(g_hat, w_hat, inv_std, cout, 1, 27, standardise, dw, direct)
        return None, dw.to(wdtype), None


def stem_conv(image, weight, standardise=True):
    y, box = StemConvFn.apply(image, weight, bool(standardise))
    if box.t is not None:
        y._mmpl_gn_stats = (box.t, 16)
    return y


# --------------------------------------------------------------------------------------------------------------
class GNReLUFn(torch.autograd.Function):
    """GroupNorm(groups, C) + ReLU, optionally two affine heads on the same input (gn1 and downsample.0 share the
    block input, unet3D.py:59-60 and :69 / :645-646).

    ``alias=True`` additionally returns the input itself as an extra output.  Callers route the *other* use of the
    block input through it (the identity residual of unet3D.py:69-71, or the encoder skip of :669-677), so that its
    gradient arrives here and is added inside the backward kernel instead of by a separate elementwise add."""

    @staticmethod
    def forward(ctx, x, gamma, beta, gamma2, beta2, groups, eps, stats_in=None, alias=False, ws=None, real_cpg=0,
                compact2=False, psplit1=False):
        _lib.require_device()
        L = _lib.lib()
        dt = _cfg["dtype"]
        x = to_cl(x, dt)
        n, c, d, h, w = x.shape
        spatial = d * h * w
        dev = x.device
        have_stats = stats_in is not None and stats_in.numel() == n * groups * 2
        stats = stats_in if have_stats else _zero_stats(n * groups * 2, dev)
        code = _lib.dtype_code(dt)
        st = _lib.stream_ptr()
        g1, b1 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        dual = gamma2 is not None
        g2 = gamma2.detach().float().contiguous() if dual else None
        b2 = beta2.detach().float().contiguous() if dual else None
        if not have_stats:
            _lib.check(L.mmpl_gn_stats(_p(x), _p(stats), n, spatial, c, groups, code, st), "gn_stats")
        y = empty_cl(n, c, d, h, w, dt, dev)
        # psplit1: the first head's MEMORY is the parity-split tensor P[pc*N + n][d/2][h/2][w/2][c] the stride-2 3x3x3
        # tensor-core convolution reads (same element count; the logical NCDHW shape is kept so that the gradient autograd
        # hands back -- a plain NDHWC tensor -- has the shape it expects).  Only that convolution may consume it.
        psplit1 = bool(psplit1)
        assert not psplit1 or (d % 2 == 0 and h % 2 == 0 and w % 2 == 0 and dt == torch.bfloat16)
        # compact2: the second head lives on the even voxels only (input of a 1x1x1 stride-2 convolution)
        compact2 = bool(dual and compact2)
        layout = (1 if psplit1 else 0) | (2 if compact2 else 0)
        cdims = (d, h, w) if compact2 else (0, 0, 0)
        y2 = None
        if dual:
            y2 = empty_cl(n, c, (d + 1) // 2, (h + 1) // 2, (w + 1) // 2, dt, dev) if compact2 else empty_cl(n, c, d, h, w, dt, dev)
        _lib.check(L.mmpl_gn_relu_fwd(_p(x), _p(stats), _p(g1), _p(b1), _p(y), _p(g2), _p(b2), _p(y2), n, spatial, c,
                                      groups, int(real_cpg), d if layout else 0, h if layout else 0, w if layout else 0,
                                      layout, eps, code, st), "gn_relu_fwd")
        ctx.save_for_backward(x, stats, g1, b1, g2, b2)
        ctx.real_cpg = int(real_cpg)
        ctx.cdims = cdims
        ctx.y2_shape = tuple(y2.shape) if dual else None
        ctx.meta = (n, c, spatial, groups, eps, dual, gamma.dtype, bool(alias))
        ctx.params = (gamma, beta, gamma2, beta2)
        ctx.ws = ws
        # what a consumer convolution needs to fold the first backward pass of this node into its dgrad epilogue
        ctx.fuse_info = None
        if ws is not None:
            ctx.fuse_info = [(b1, ws, 0)] + ([(b2, ws, 1)] if dual else [])
        outs = [y]
        if dual:
            outs.append(y2)
        if alias:
            outs.append(x.view_as(x))
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        L = _lib.lib()
        x, stats, g1, b1, g2, b2 = ctx.saved_tensors
        n, c, spatial, groups, eps, dual, pdtype, alias = ctx.meta
        grads = list(grads)
        dy = grads.pop(0)
        dy2 = grads.pop(0) if dual else None
        dres = grads.pop(0) if alias else None
        dt = x.dtype
        dev = x.device
        code = _lib.dtype_code(dt)
        st = _lib.stream_ptr()
        if dy is None:
            dy = torch.zeros_like(x)
        dy = to_cl(dy, dt)
        if dual:
            if dy2 is None:
                dy2 = torch.zeros(ctx.y2_shape, dtype=dt, device=dev)
            dy2 = to_cl(dy2, dt)
            assert tuple(dy2.shape) == ctx.y2_shape
        if dres is not None:
            dres = to_cl(dres, dt)
        dx = torch.empty_like(x)
        pg, pb, pg2, pb2 = ctx.params
        dg1, db1 = _grad_dst(pg, (c,)), _grad_dst(pb, (c,))
        dg2 = _grad_dst(pg2, (c,)) if dual else None
        db2 = _grad_dst(pb2, (c,)) if dual else None
        ws = ctx.ws
        reduced = 0
        if ws is not None:
            # the producers of dy (and dy2) may already have accumulated the reduction sums into ws
            t1 = _GN_REDUCED.pop(dy.data_ptr(), None)
            t2 = _GN_REDUCED.pop(dy2.data_ptr(), None) if dual else None
            if t1 == (ws.data_ptr(), 0) and (not dual or t2 == (ws.data_ptr(), 1)):
                reduced = 1
        else:
            ws = torch.empty(n * c * 6 + 2, dtype=torch.float64, device=dev)
        _lib.check(L.mmpl_gn_relu_bwd(_p(x), _p(stats), _p(g1), _p(b1), _p(dy), _p(g2), _p(b2), _p(dy2) if dual else None,
                                      _p(dres), _p(dx), _p(dg1), _p(db1), _p(dg2), _p(db2), _p(ws), reduced, n, spatial, c,
                                      groups, ctx.real_cpg, ctx.cdims[0], ctx.cdims[1], ctx.cdims[2],
                                      2 if ctx.cdims[2] else 0, eps, code, st), "gn_relu_bwd")
        if dual:
            return (dx, dg1.to(pdtype), db1.to(pdtype), dg2.to(pdtype), db2.to(pdtype)) + (None,) * 8
        return (dx, dg1.to(pdtype), db1.to(pdtype)) + (None,) * 10


def _attached_stats(x, groups):
    st = getattr(x, "_mmpl_gn_stats", None)
    return st[0] if (st is not None and st[1] == groups) else None


def _gn_bwd_ws(x):
    """Backward workspace [N][C][4] (fp64, zero) of a GroupNorm node, taken from the forward's zero pool."""
    if not (_cfg["fuse_gn_bwd"] and torch.is_grad_enabled() and x.is_cuda):
        return None
    return _zero_stats(x.shape[0] * x.shape[1] * 6 + 2, x.device)     # [N][C][6] + ticket (mmpl_gn_relu_bwd)


def _tag_gn_outputs(outs, n_heads):
    """Attach to each GN+ReLU output what its consumer needs to fuse this node's backward reduction (head h)."""
    first = outs[0] if isinstance(outs, tuple) else outs
    fn = first.grad_fn
    info = getattr(fn, "fuse_info", None) if fn is not None else None
    if info:
        ys = outs if isinstance(outs, tuple) else (outs,)
        for h in range(n_heads):
            ys[h]._mmpl_gn_bwd = info[h]
    return outs


def gn_relu(x, gamma, beta, groups=16, eps=1e-5, alias=False, real_cpg=0):
    """-> y, or (y, x_alias) with alias=True.  ``real_cpg``: channels per group that carry data (zero-padded layouts)."""
    outs = GNReLUFn.apply(x, gamma, beta, None, None, int(groups), float(eps), _attached_stats(x, groups), bool(alias),
                          _gn_bwd_ws(x), int(real_cpg))
    return _tag_gn_outputs(outs, 1)


def gn_relu_dual(x, gamma, beta, gamma2, beta2, groups=16, eps=1e-5, alias=False, real_cpg=0, compact2=False,
                 psplit1=False):
    """-> (y, y2), or (y, y2, x_alias) with alias=True.  ``compact2``: y2 only on the even voxels,
    [N, C, ceil(D/2), ceil(H/2), ceil(W/2)] == the full y2[:, :, ::2, ::2, ::2] (what a 1x1x1 stride-2 convolution reads).
    ``psplit1``: y is laid out in memory as the parity-split tensor of the stride-2 3x3x3 tensor-core convolution (even
    extents, bf16 tensor-core path; see ``psplit_consumer``) -- pass it to that convolution and nothing else."""
    psplit1 = bool(psplit1) and all(int(v) % 2 == 0 for v in x.shape[2:]) and _cfg["dtype"] == torch.bfloat16
    outs = GNReLUFn.apply(x, gamma, beta, gamma2, beta2, int(groups), float(eps), _attached_stats(x, groups),
                          bool(alias), _gn_bwd_ws(x), int(real_cpg), bool(compact2), psplit1)
    if psplit1:
        outs[0]._mmpl_psplit = True       # read by WSConv3dFn.forward

    return _tag_gn_outputs(outs, 2)


# --------------------------------------------------------------------------------------------------------------
class Upsample2xAddFn(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='trilinear') followed by ``+ skip`` (unet3D.py:608, :686-687)."""

    @staticmethod
    def forward(ctx, x_lo, skip):
        _lib.require_device()
        L = _lib.lib()
        dt = _cfg["dtype"]
        x_lo, skip = to_cl(x_lo, dt), to_cl(skip, dt)
        n, c, d, h, w = x_lo.shape
        assert tuple(skip.shape) == (n, c, 2 * d, 2 * h, 2 * w), f"skip {tuple(skip.shape)} vs 2x of {tuple(x_lo.shape)}"
        y = empty_cl(n, c, 2 * d, 2 * h, 2 * w, dt, x_lo.device)
        # GroupNorm(16) raw sums of the output (consumed by the next block's gn1 / downsample.0), when the channel
        # vectors of a row tile evenly over a 256-thread block
        vn = 8 if dt == torch.bfloat16 else 4
        fuse = c % 16 == 0 and c <= 512 and 256 % (c // vn) == 0
        stats = _zero_stats(n * 16 * 2, x_lo.device) if fuse else None
        _lib.check(L.mmpl_upsample2x_add_fwd(_p(x_lo), _p(skip), _p(y), n, d, h, w, c, _lib.dtype_code(dt), _p(stats),
                                             _lib.stream_ptr()), "upsample2x_add_fwd")
        ctx.meta = (n, c, d, h, w, dt)
        return y, _StatsBox(stats)

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        L = _lib.lib()
        n, c, d, h, w, dt = ctx.meta
        dy = to_cl(dy, dt)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = empty_cl(n, c, d, h, w, dt, dy.device)
            _lib.check(L.mmpl_upsample2x_bwd(_p(dy), _p(dx), n, d, h, w, c, _lib.dtype_code(dt), _lib.stream_ptr()),
                       "upsample2x_bwd")
        return dx, (dy if ctx.needs_input_grad[1] else None)


def upsample2x_add(x_lo, skip):
    y, box = Upsample2xAddFn.apply(x_lo, skip)
    if box.t is not None:
        y._mmpl_gn_stats = (box.t, 16)
    return y


# --------------------------------------------------------------------------------------------------------------
class ClassifierFn(torch.autograd.Function):
    """nn.Conv3d(base, classes, 1) with bias -> fp32 NCDHW logits (unet3D.py:632, :713)."""

    @staticmethod
    def forward(ctx, a, weight, bias):
        _lib.require_device()
        L = _lib.lib()
        dt = _cfg["dtype"]
        ctx.gn_bwd = getattr(a, "_mmpl_gn_bwd", None)      # a = relu(gn(.)): its backward reduction rides on cls_bwd
        a = to_cl(a, dt)
        n, cin, d, h, w = a.shape
        classes = weight.shape[0]
        spatial = d * h * w
        wc = weight.detach().float().reshape(classes, cin).contiguous()
        b = bias.detach().float().contiguous()
        logits = torch.empty((n, classes, d, h, w), dtype=torch.float32, device=a.device)
        _lib.check(L.mmpl_cls_fwd(_p(a), _p(wc), _p(b), _p(logits), n, spatial, cin, classes, _lib.dtype_code(dt),
                                  _lib.stream_ptr()), "cls_fwd")
        ctx.save_for_backward(a, wc)
        ctx.meta = (n, cin, d, h, w, classes, weight.dtype, tuple(weight.shape))
        ctx.params = (weight, bias)
        return logits

    @staticmethod
    def backward(ctx, dl):
        L = _lib.lib()
        a, wc = ctx.saved_tensors
        n, cin, d, h, w, classes, wdtype, wshape = ctx.meta
        dl = dl.float().contiguous()
        da = torch.empty_like(a)
        dwc = _grad_dst(ctx.params[0], wshape)
        db = _grad_dst(ctx.params[1], (classes,))
        gb = gws = None
        # The backward reduction of precls_conv.0/1 rides on the bf16 warp-MMA kernel (8 extra loads + 12 FMAs per 16x8
        # fragment); the fp32 CUDA-core kernel is issue-bound, there the stand-alone reduction pass is cheaper.
        fuse_cls = _cfg["fuse_gn_bwd_cls"] if _cfg["fuse_gn_bwd_cls"] is not None else a.dtype == torch.bfloat16
        if fuse_cls and ctx.gn_bwd is not None and ctx.gn_bwd[2] == 0 and cin in (32, 64):
            gb, gws, _ = ctx.gn_bwd
        _lib.check(L.mmpl_cls_bwd(_p(a), _p(wc), _p(dl), _p(da), _p(dwc), _p(db), _p(gb), _p(gws), n, d * h * w, cin,
                                  classes, _lib.dtype_code(a.dtype), _lib.stream_ptr()), "cls_bwd")
        if gws is not None:
            _GN_REDUCED[da.data_ptr()] = (gws.data_ptr(), 0)
        return da, dwc.to(wdtype), db.to(wdtype)


def classifier(a, weight, bias):
    return ClassifierFn.apply(a, weight, bias)


class LayerNormRowsFn(torch.autograd.Function):
    """LayerNorm over the channel axis of a channels-last volume, WITHOUT affine: xhat = (x - mean_c) * rstd_c per voxel.
    The per-voxel rows of ``EAM.norm2`` (unet3D.py:197); its affine is folded into the attention matrix (csrc/eam.cu)."""

    @staticmethod
    def forward(ctx, x, eps):
        _lib.require_device()
        L = _lib.lib()
        dt = _cfg["dtype"]
        x = to_cl(x, dt)
        n, c, d, h, w = x.shape
        rows = n * d * h * w
        y = empty_cl(n, c, d, h, w, dt, x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        _lib.check(L.mmpl_ln_rows_fwd(_p(x), _p(y), _p(rstd), rows, c, float(eps), _lib.dtype_code(dt), _lib.stream_ptr()),
                   "ln_rows_fwd")
        ctx.save_for_backward(y, rstd)
        return y

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        y, rstd = ctx.saved_tensors
        g = to_cl(g, y.dtype)
        dx = torch.empty_like(y)
        _lib.check(L.mmpl_ln_rows_bwd(_p(y), _p(rstd), _p(g), _p(dx), rstd.numel(), y.shape[1], _lib.dtype_code(y.dtype),
                                      _lib.stream_ptr()), "ln_rows_bwd")
        return dx, None


def layer_norm_rows(x, eps=1e-5):
    return LayerNormRowsFn.apply(x, float(eps))


@torch.no_grad()
def renew_tokens(token, feature, mask, alpha):
    """In-place EMA of the class tokens (unet3D_with_feam3.renew_token, unet3D.py:1051-1068) for ONE feature level:
    token [ntok, C] fp32 on the device; feature [N, C, D, H, W]; mask [N, 1, Dm, Hm, Wm] class ids (float or uint8) at
    any resolution (nearest-neighbour sampled at the feature resolution like F.interpolate(mode='nearest')).  Two
    launches, no host synchronisation; classes absent from the (down-sampled) mask keep their token."""
    _lib.require_device()
    L = _lib.lib()
    x = feature.detach()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = to_cl(x, x.dtype)
    n, c, d, h, w = x.shape
    m = mask.detach()
    u8 = m.dtype == torch.uint8
    m = m.contiguous() if u8 else m.float().contiguous()
    dm, hm, wm = m.shape[-3:]
    assert m.numel() == n * dm * hm * wm, f"mask {tuple(mask.shape)} vs features {tuple(feature.shape)}"
    assert token.dtype == torch.float32 and token.is_contiguous() and token.shape[1] == c
    ntok = token.shape[0]
    sums = torch.empty(ntok * c, dtype=torch.float32, device=x.device)
    cnt = torch.empty(ntok, dtype=torch.float32, device=x.device)
    st = _lib.stream_ptr()
    _lib.check(L.mmpl_token_stats(_p(x), _p(m), int(u8), _p(sums), _p(cnt), n, d, h, w, c, dm, hm, wm, ntok,
                                  _lib.dtype_code(x.dtype), st), "token_stats")
    _lib.check(L.mmpl_token_ema(_p(token), _p(sums), _p(cnt), ntok, c, float(alpha), st), "token_ema")
    return token


class SpaceToDepth2Fn(torch.autograd.Function):
    """[N,C,D,H,W] -> [N,cp,D/2,H/2,W/2] with channel (pd*4+ph*2+pw)*C + c (zero beyond 8C): the input side of the
    4x4x4 stride-2 -> 3x3x3 stride-1 rewrite of the discriminator's convolutions (csrc/aux_nets.cu)."""

    @staticmethod
    def forward(ctx, x, cp):
        _lib.require_device()
        dt = _cfg["dtype"]
        x = to_cl(x, dt)
        n, c, d, h, w = x.shape
        y = empty_cl(n, cp, d // 2, h // 2, w // 2, dt, x.device)
        _lib.check(_lib.lib().mmpl_space_to_depth2(_p(x), _p(y), n, d, h, w, c, cp, 0, _lib.dtype_code(dt), _lib.stream_ptr()),
                   "space_to_depth2")
        ctx.meta = (n, c, d, h, w, cp, dt)
        return y

    @staticmethod
    def backward(ctx, g):
        n, c, d, h, w, cp, dt = ctx.meta
        g = to_cl(g, dt)
        dx = empty_cl(n, c, d, h, w, dt, g.device)
        _lib.check(_lib.lib().mmpl_space_to_depth2(_p(g), _p(dx), n, d, h, w, c, cp, 1, _lib.dtype_code(dt), _lib.stream_ptr()),
                   "depth_to_space2")
        return dx, None


def space_to_depth2(x, cp):
    return SpaceToDepth2Fn.apply(x, int(cp))


class BiasLeakyReLUFn(torch.autograd.Function):
    """leaky_relu(x + bias[c], slope) on a channels-last activation: nn.Conv3d bias + nn.LeakyReLU(0.2) of the
    discriminator blocks (unet3D.py:1912-1936), one pass forward, one backward (dx and dbias)."""

    @staticmethod
    def forward(ctx, x, bias, slope):
        _lib.require_device()
        dt = _cfg["dtype"]
        x = to_cl(x, dt)
        n, c, d, h, w = x.shape
        b = bias.detach().float().contiguous()
        y = torch.empty_like(x)
        _lib.check(_lib.lib().mmpl_bias_lrelu_fwd(_p(x), _p(b), _p(y), n * d * h * w, c, float(slope), _lib.dtype_code(dt),
                                                  _lib.stream_ptr()), "bias_lrelu_fwd")
        ctx.save_for_backward(y)
        ctx.meta = (n * d * h * w, c, float(slope), bias.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        rows, c, slope, bdtype = ctx.meta
        g = to_cl(g, y.dtype)
        dx = torch.empty_like(y)
        db = torch.empty(c, dtype=torch.float32, device=y.device)
        _lib.check(_lib.lib().mmpl_bias_lrelu_bwd(_p(y), _p(g), _p(dx), _p(db), rows, c, slope, _lib.dtype_code(y.dtype),
                                                  _lib.stream_ptr()), "bias_lrelu_bwd")
        return dx, db.to(bdtype), None


def bias_leaky_relu(x, bias, slope=0.2):
    return BiasLeakyReLUFn.apply(x, bias, float(slope))


class Upsample2xNCDHWFn(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='trilinear') of an fp32 NCDHW tensor (the refiner's logits, unet3D.py:1621)."""

    @staticmethod
    def forward(ctx, x):
        _lib.require_device()
        x = x.float().contiguous()
        n, c, d, h, w = x.shape
        y = torch.empty((n, c, 2 * d, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mmpl_upsample2x_ncdhw_fwd(_p(x), _p(y), n * c, d, h, w, _lib.stream_ptr()), "upsample2x_ncdhw_fwd")
        ctx.meta = (n, c, d, h, w)
        return y

    @staticmethod
    def backward(ctx, g):
        n, c, d, h, w = ctx.meta
        g = g.float().contiguous()
        dx = torch.empty((n, c, d, h, w), dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib().mmpl_upsample2x_ncdhw_bwd(_p(g), _p(dx), n * c, d, h, w, _lib.stream_ptr()), "upsample2x_ncdhw_bwd")
        return dx


def upsample2x_ncdhw(x):
    return Upsample2xNCDHWFn.apply(x)


class BlendSink:
    """Where the sliding-window classifier accumulates (predict_sliding, evaluate_amos.py:261-276): fp32 accumulator
    ``acc`` [B, D, C, H, W] (depth-major, ``d_outer``) or [B, C, D, H, W], optional weight sum ``wsum`` [B, D, H, W], the
    Gaussian importance map of the tile and the tile origin as a DEVICE int32[3] (so a captured graph serves all tiles).
    ``origin_dev`` may also be int32[T, 3] with B = 1: a batch of T tiles of the SAME volume goes through the network in one
    forward and is accumulated tile by tile, in order, into ``acc[0]`` (stream-ordered launches, so overlapping tiles add in
    the order the reference visits them)."""

    def __init__(self, acc, gauss, origin_dev, tile, d_outer=True, wsum=None):
        assert acc.dtype == torch.float32 and acc.is_contiguous() and gauss.dtype == torch.float32
        assert origin_dev.dtype == torch.int32 and origin_dev.numel() % 3 == 0 and origin_dev.is_cuda
        assert origin_dev.is_contiguous() and (origin_dev.numel() == 3 or acc.shape[0] == 1), \
            "a tile batch (origin_dev [T, 3]) accumulates into ONE volume"
        self.tiles = origin_dev.numel() // 3
        self.acc, self.gauss, self.origin_dev, self.wsum = acc, gauss.contiguous(), origin_dev, wsum
        self.tile = tuple(int(t) for t in tile)
        self.d_outer = bool(d_outer)
        if d_outer:
            self.B, self.D, self.C, self.H, self.W = acc.shape
        else:
            self.B, self.C, self.D, self.H, self.W = acc.shape


def classifier_blend_supported(a_channels, classes, dtype) -> bool:
    return dtype == torch.bfloat16 and a_channels in (32, 64) and classes <= 16


@torch.no_grad()
def classifier_blend(a, weight, bias, sink: BlendSink):
    """acc += gauss * (classifier(a)) at the sink's tile origin: nn.Conv3d(base, classes, 1) (unet3D.py:632) fused with
    the Gaussian-weighted accumulation of predict_sliding -- no fp32 logits tile is written.  bf16 activations only."""
    _lib.require_device()
    L = _lib.lib()
    a = to_cl(a, torch.bfloat16)
    n, cin, d, h, w = a.shape
    classes = weight.shape[0]
    batch_of_tiles = sink.tiles > 1          # n tiles of one volume: acc[0], origin row i
    assert (d, h, w) == sink.tile and n == (sink.tiles if batch_of_tiles else sink.B) and classes == sink.C, \
        "tile / batch / classes do not match the sink"
    wc = weight.detach().float().reshape(classes, cin).contiguous()
    b = bias.detach().float().contiguous()
    a_rows = a.permute(0, 2, 3, 4, 1)      # the NDHWC storage
    origins = sink.origin_dev.reshape(-1, 3)
    for i in range(n):
        v = 0 if batch_of_tiles else i
        _lib.check(L.mmpl_cls_blend(_p(a_rows[i]), _p(wc), _p(b), _p(sink.gauss), _p(sink.acc[v]),
                                    None if sink.wsum is None else _p(sink.wsum[v]),
                                    _p(origins[i if batch_of_tiles else 0]), classes,
                                    sink.D, sink.H, sink.W, d, h, w, cin, int(sink.d_outer), _lib.stream_ptr()), "cls_blend")


# --------------------------------------------------------------------------------------------------------------
class PartialLossFn(torch.autograd.Function):
    """EDiceLoss_partial.forward with soft_max=True (loss_partial.py:71-99): one fused forward pass, one fused
    backward pass, no host synchronisation.  ``target`` may be float class ids (the reference) or uint8.
    ``class_weight`` [C] (= mask[0], pooled over the batch like the reference) or [N, C] with ``per_sample`` (the
    reference formula per sample with that sample's weights, averaged over the batch); ``lut`` likewise."""

    @staticmethod
    def forward(ctx, logits, target, class_weight, lut, uce, per_sample):
        _lib.require_device()
        L = _lib.lib()
        z = logits.detach().float().contiguous()
        n, c = z.shape[0], z.shape[1]
        spatial = z[0, 0].numel()
        t = target.detach()
        u8 = t.dtype == torch.uint8
        t = t.contiguous() if u8 else t.float().contiguous()
        assert t.numel() == n * spatial, f"target {tuple(target.shape)} does not match logits {tuple(logits.shape)}"
        dev = z.device
        groups = n if per_sample else 1
        cw = class_weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        assert cw.numel() == groups * c, f"class weights {tuple(class_weight.shape)} for {groups} group(s) of {c} classes"
        lt = None if lut is None else lut.detach().to(device=dev, dtype=torch.float32).contiguous()
        assert lt is None or lt.numel() == groups * c
        sums = torch.empty(groups * 4 * c + 1, dtype=torch.float64, device=dev)   # [G][4][C] sums + the ticket slot
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(L.mmpl_partial_loss_fwd(_p(z), _p(t), int(u8), _p(cw), _p(lt), int(per_sample), _p(sums), _p(loss), n,
                                           spatial, c, int(uce), _lib.stream_ptr()), "partial_loss_fwd")
        ctx.save_for_backward(z, t, cw, lt, sums)
        ctx.meta = (n, spatial, c, int(uce), logits.dtype, int(u8), int(per_sample))
        return loss

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        z, t, cw, lt, sums = ctx.saved_tensors
        n, spatial, c, uce, ldtype, u8, per_sample = ctx.meta
        g = gout.detach().float().contiguous()
        dz = torch.empty_like(z)
        _lib.check(L.mmpl_partial_loss_bwd(_p(z), _p(t), u8, _p(cw), _p(lt), per_sample, _p(sums), _p(g), _p(dz), n,
                                           spatial, c, uce, _lib.stream_ptr()), "partial_loss_bwd")
        return dz.to(ldtype), None, None, None, None, None


def partial_label_loss(logits, target, class_weight, lut=None, uce=True, per_sample=False):
    return PartialLossFn.apply(logits, target, class_weight, lut, bool(uce), bool(per_sample))


class ClassifierPartialLossFn(torch.autograd.Function):
    """``partial_label_loss(classifier(a, weight, bias), target, ...)`` as two launches (csrc/cls_loss.cu): the fp32 logits
    and their gradient only ever exist in the registers of the warp MMAs.  Forward: a + labels -> per-class sums -> loss.
    Backward: logits recomputed from ``a``, dA / dW / db and the GroupNorm-backward reduction of the node that produced
    ``a`` -- everything ClassifierFn.backward does."""

    @staticmethod
    def forward(ctx, a, weight, bias, target, class_weight, lut, uce, per_sample):
        _lib.require_device()
        L = _lib.lib()
        ctx.gn_bwd = getattr(a, "_mmpl_gn_bwd", None)
        a = to_cl(a, torch.bfloat16)
        n, cin, d, h, w = a.shape
        classes = weight.shape[0]
        spatial = d * h * w
        dev = a.device
        wc = weight.detach().float().reshape(classes, cin).contiguous()
        b = bias.detach().float().contiguous()
        t = target.detach()
        u8 = t.dtype == torch.uint8
        t = t.contiguous() if u8 else t.float().contiguous()
        assert t.numel() == n * spatial, f"target {tuple(target.shape)} does not match activations {tuple(a.shape)}"
        groups = n if per_sample else 1
        cw = class_weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        assert cw.numel() == groups * classes, \
            f"class weights {tuple(class_weight.shape)} for {groups} group(s) of {classes} classes"
        lt = None if lut is None else lut.detach().to(device=dev, dtype=torch.float32).contiguous()
        assert lt is None or lt.numel() == groups * classes
        sums = torch.empty(groups * 4 * classes + 1, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(L.mmpl_cls_loss_fwd(_p(a), _p(wc), _p(b), _p(t), int(u8), _p(cw), _p(lt), int(per_sample), _p(sums),
                                       _p(loss), n, spatial, cin, classes, int(uce), _lib.stream_ptr()), "cls_loss_fwd")
        ctx.save_for_backward(a, wc, b, t, cw, lt, sums)
        ctx.meta = (n, cin, d, h, w, classes, weight.dtype, tuple(weight.shape), int(u8), int(uce), int(per_sample))
        ctx.params = (weight, bias)
        return loss

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        a, wc, b, t, cw, lt, sums = ctx.saved_tensors
        n, cin, d, h, w, classes, wdtype, wshape, u8, uce, per_sample = ctx.meta
        g = gout.detach().float().contiguous()
        da = torch.empty_like(a)
        dwc = _grad_dst(ctx.params[0], wshape)
        db = _grad_dst(ctx.params[1], (classes,))
        gb = gws = None
        fuse_cls = _cfg["fuse_gn_bwd_cls"] if _cfg["fuse_gn_bwd_cls"] is not None else True
        if fuse_cls and ctx.gn_bwd is not None and ctx.gn_bwd[2] == 0:
            gb, gws, _ = ctx.gn_bwd
        _lib.check(L.mmpl_cls_loss_bwd(_p(a), _p(wc), _p(b), _p(t), u8, _p(cw), _p(lt), per_sample, _p(sums), _p(g), _p(da),
                                       _p(dwc), _p(db), _p(gb), _p(gws), n, d * h * w, cin, classes, uce,
                                       _lib.stream_ptr()), "cls_loss_bwd")
        if gws is not None:
            _GN_REDUCED[da.data_ptr()] = (gws.data_ptr(), 0)
        return da, dwc.to(wdtype), db.to(wdtype), None, None, None, None, None


def classifier_partial_loss_supported(cin: int, classes: int, dtype=None) -> bool:
    dtype = _cfg["dtype"] if dtype is None else dtype
    return dtype == torch.bfloat16 and cin in (32, 64) and 1 <= classes <= 16


def classifier_partial_loss(a, weight, bias, target, class_weight, lut=None, uce=True, per_sample=False):
    """Loss of the classifier applied to ``a`` without materialising the logits when the fused kernels cover the shape
    (bf16 activations, 32/64 channels, <= 16 classes); the two-step composition otherwise -- same value, same gradients."""
    if classifier_partial_loss_supported(weight.shape[1], weight.shape[0]):
        return ClassifierPartialLossFn.apply(a, weight, bias, target, class_weight, lut, bool(uce), bool(per_sample))
    return partial_label_loss(classifier(a, weight, bias), target, class_weight, lut, uce, per_sample)


# --------------------------------------------------------------------------------------------------------------
class MaskedDiceFn(torch.autograd.Function):
    """DiceLoss._dice_loss over a voxel gate, optionally on sigmoid(x) and with the BCE-with-logits term of
    EDiceLoss_full2.forward (loss_partial.py:24-36, :150-170): one fused pass forward, one backward."""

    @staticmethod
    def forward(ctx, x, target, gate, sigmoid, uce):
        _lib.require_device()
        L = _lib.lib()
        xf = x.detach().float().contiguous()
        tf = target.detach().float().contiguous()
        v = xf.numel()
        assert tf.numel() == v, f"masked dice: score {tuple(x.shape)} vs target {tuple(target.shape)}"
        gf = None
        if gate is not None:
            gf = gate.detach().to(torch.float32).contiguous()
            assert gf.numel() == v, f"masked dice: gate {tuple(gate.shape)} vs score {tuple(x.shape)}"
        dev = xf.device
        sums = torch.empty(5, dtype=torch.float64, device=dev)             # I, Y, Z, E + the kernel's ticket slot
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(L.mmpl_masked_dice_fwd(_p(xf), _p(tf), _p(gf), _p(sums), _p(loss), v, int(sigmoid), int(uce),
                                          _lib.stream_ptr()), "masked_dice_fwd")
        ctx.save_for_backward(xf, tf, gf, sums)
        ctx.meta = (v, int(sigmoid), int(uce), x.dtype, tuple(x.shape), target.dtype, tuple(target.shape))
        return loss

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        xf, tf, gf, sums = ctx.saved_tensors
        v, sigmoid, uce, xdt, xshape, tdt, tshape = ctx.meta
        g = gout.detach().float().contiguous()
        dx = torch.empty_like(xf)
        dt = torch.empty_like(tf) if ctx.needs_input_grad[1] else None
        _lib.check(L.mmpl_masked_dice_bwd(_p(xf), _p(tf), _p(gf), _p(sums), _p(g), _p(dx), _p(dt), v, sigmoid, uce,
                                          _lib.stream_ptr()), "masked_dice_bwd")
        return (dx.view(xshape).to(xdt) if ctx.needs_input_grad[0] else None,
                dt.view(tshape).to(tdt) if dt is not None else None, None, None, None)


def masked_dice(x, target, gate=None, sigmoid=False, uce=False):
    return MaskedDiceFn.apply(x, target, gate, bool(sigmoid), bool(uce))
