"""CPU oracle for the multimodal-PL hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32/fp64 PyTorch + numpy on the CPU, the algorithm of the reference's dense hot
path so that the CUDA kernels can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package never does.

Parity pin: the reference (TThuraya/multimodal-PL) ships no tests and no golden vectors.
The oracle is therefore pinned against *outputs of the reference itself*: ``oracle/make_golden.py`` imports the
unmodified reference modules from ``/root/reference`` in the authoring container and writes small fixtures to
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them against this file on every CPU test run.

All arithmetic below the Python level is PyTorch ATen (torch 2.11.0+cu128, oneDNN 3.10 on CPU) exactly as in
the reference; this file only re-expresses *which* ATen ops are applied in *which* order, citing the reference.
Citations are ``path:line`` relative to the reference repository root.
"""
from __future__ import annotations

import csv
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------------------
# Weight standardisation + convolution                                                    unet3D.py:16-27
# ----------------------------------------------------------------------------------------------------------------


def ws_weight(w: torch.Tensor) -> torch.Tensor:
    """Per-out-channel standardised weight (unet3D.py:22-26): nested means over dims 1..4, unbiased variance of
    the centred weight, eps 1e-12 under the square root."""
    m = w.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True).mean(dim=3, keepdim=True).mean(dim=4, keepdim=True)
    c = w - m
    std = torch.sqrt(torch.var(c.view(c.size(0), -1), dim=1) + 1e-12).view(-1, 1, 1, 1, 1)
    return c / std


def ws_conv3d(x: torch.Tensor, w: torch.Tensor, stride: int, padding: int) -> torch.Tensor:
    """Conv3d.forward (unet3D.py:21-27): F.conv3d with the standardised weight, no bias, dilation 1, groups 1."""
    return F.conv3d(x, ws_weight(w), None, stride, padding, 1, 1)


def gn_relu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 16) -> torch.Tensor:
    """nn.GroupNorm(16, C) (eps 1e-5, affine) followed by ReLU (unet3D.py:44,47,49,59-60,64-65)."""
    return F.relu(F.group_norm(x, groups, gamma, beta, 1e-5))


# ----------------------------------------------------------------------------------------------------------------
# Storage-precision emulation (test infrastructure for the bf16 production path; nothing in the reference)
# ----------------------------------------------------------------------------------------------------------------
# The reference computes everything in fp32.  The B200 path stores every activation tensor between
# kernels -- and every gradient tensor between backward kernels -- in bf16 while accumulating in fp32, and feeds the
# tensor cores bf16 copies of the standardised weights.  ``store_dtype=torch.bfloat16`` makes this oracle round at
# exactly those points (round-to-nearest-even, like the kernels' cvt.rn.bf16x2.f32), so the CUDA path can be held to a
# real per-tensor gradient tolerance instead of the "a few ReLU gates resolve differently" argument: both sides then
# see the same stored values up to fp32 summation order.


class _RoundStored(torch.autograd.Function):
    """y = round_to(dtype)(x) in the forward pass; the incoming gradient is rounded the same way in the backward pass
    (a tensor that is stored in bf16 has its gradient stored in bf16 by the backward kernel that produces it)."""

    @staticmethod
    def forward(ctx, x, dtype, round_grad):
        ctx.dtype, ctx.round_grad = dtype, round_grad
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        if ctx.round_grad:
            g = g.to(ctx.dtype).to(g.dtype)
        return g, None, None


def stored(x: torch.Tensor, store_dtype: Optional[torch.dtype], round_grad: bool = True) -> torch.Tensor:
    """Identity for ``store_dtype`` None / float32, else the value as it is after a round trip through HBM."""
    if store_dtype is None or store_dtype == x.dtype:
        return x
    return _RoundStored.apply(x, store_dtype, round_grad)


# ----------------------------------------------------------------------------------------------------------------
# Topology of unet3D_baseline                                                            unet3D.py:585-718
# ----------------------------------------------------------------------------------------------------------------


def backbone_spec(base: int = 32) -> List[Tuple[str, int, int, int]]:
    """(prefix, inplanes, planes, stride) for every NoBottleneck in forward order, layers=[1,2,2,2,2]
    (unet3D.py:596-626, _make_layer :641-661).  ``base`` = 32 reproduces the reference; 64 is the wide stress
    config (BASELINE.json configs[4]) built from the same recipe with doubled widths."""
    b = base
    enc = [
        ("layer0.0.", b, b, 1),
        ("layer1.0.", b, 2 * b, 2), ("layer1.1.", 2 * b, 2 * b, 1),
        ("layer2.0.", 2 * b, 4 * b, 2), ("layer2.1.", 4 * b, 4 * b, 1),
        ("layer3.0.", 4 * b, 8 * b, 2), ("layer3.1.", 8 * b, 8 * b, 1),
        ("layer4.0.", 8 * b, 8 * b, 2), ("layer4.1.", 8 * b, 8 * b, 1),
    ]
    dec = [
        ("x8_resb.0.", 8 * b, 4 * b, 1),
        ("x4_resb.0.", 4 * b, 2 * b, 1),
        ("x2_resb.0.", 2 * b, b, 1),
        ("x1_resb.0.", b, b, 1),
    ]
    return enc + dec


def state_dict_shapes(base: int = 32, num_classes: int = 16) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape for the 107 tensors of unet3D_baseline (probed from the reference)."""
    shapes: Dict[str, Tuple[int, ...]] = {"conv1.weight": (base, 1, 3, 3, 3)}
    for p, cin, cout, stride in backbone_spec(base):
        shapes[p + "gn1.weight"] = (cin,)
        shapes[p + "gn1.bias"] = (cin,)
        shapes[p + "conv1.weight"] = (cout, cin, 3, 3, 3)
        shapes[p + "gn2.weight"] = (cout,)
        shapes[p + "gn2.bias"] = (cout,)
        shapes[p + "conv2.weight"] = (cout, cout, 3, 3, 3)
        if stride != 1 or cin != cout:
            shapes[p + "downsample.0.weight"] = (cin,)
            shapes[p + "downsample.0.bias"] = (cin,)
            shapes[p + "downsample.2.weight"] = (cout, cin, 1, 1, 1)
    shapes["fusionConv.0.weight"] = (8 * base,)
    shapes["fusionConv.0.bias"] = (8 * base,)
    shapes["fusionConv.2.weight"] = (8 * base, 8 * base, 1, 1, 1)
    shapes["precls_conv.0.weight"] = (base,)
    shapes["precls_conv.0.bias"] = (base,)
    shapes["precls_conv.2.weight"] = (num_classes, base, 1, 1, 1)
    shapes["precls_conv.2.bias"] = (num_classes,)
    return shapes


def synth_state_dict(base: int = 32, num_classes: int = 16, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic random weights keyed by tensor name (independent of module construction order, so the same
    values can be loaded into the reference model, this oracle and the CUDA path)."""
    sd: Dict[str, torch.Tensor] = {}
    for i, (k, shp) in enumerate(sorted(state_dict_shapes(base, num_classes).items())):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if k.endswith("gn1.weight") or k.endswith("gn2.weight") or k.endswith(".0.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias") and len(shp) == 1 and not k.startswith("precls_conv.2"):
            t = 0.1 * torch.randn(shp, generator=g)
        elif k.startswith("precls_conv.2"):
            bound = 1.0 / math.sqrt(base)
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        else:
            fan_in = int(np.prod(shp[1:]))
            t = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        sd[k] = t.float()
    return sd


def synth_feam3_state_dict(num_classes: int = 16, seed: int = 0):
    """Deterministic weights for unet3D_with_feam3 (reference unet3D.py:938-1017): the baseline backbone weights of
    ``synth_state_dict`` plus the three deep-supervision heads and the three EAM modules; also returns the three class
    tokens (plain tensors in the reference, :1006-1011).  -> (state_dict, [token1, token2, token3])."""
    sd = synth_state_dict(32, num_classes, seed)
    extra = {}
    for name, c in (("1", 128), ("2", 64), ("3", 32)):
        extra[f"deepout{name}.0.weight"] = (c,)
        extra[f"deepout{name}.0.bias"] = (c,)
        extra[f"deepout{name}.2.weight"] = (num_classes, c, 1, 1, 1)
        extra[f"deepout{name}.2.bias"] = (num_classes,)
    for name, c in (("eam84", 128), ("eam42", 64), ("eam21", 32)):
        extra[name + ".kv.weight"] = (2 * c, c)
        extra[name + ".q.weight"] = (c, c)
        extra[name + ".proj.weight"] = (c, c)
        extra[name + ".proj.bias"] = (c,)
        for nrm in ("norm2", "norm3"):
            extra[f"{name}.{nrm}.weight"] = (c,)
            extra[f"{name}.{nrm}.bias"] = (c,)
    for i, (k, shp) in enumerate(sorted(extra.items())):
        g = torch.Generator().manual_seed(seed * 100003 + 5000 + i)
        if k.endswith(".0.weight") or k.endswith("norm2.weight") or k.endswith("norm3.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            t = torch.randn(shp, generator=g) / math.sqrt(int(np.prod(shp[1:])))
        sd[k] = t.float()
    tokens = [torch.randn((num_classes - 1, c), generator=torch.Generator().manual_seed(seed * 100003 + 7000 + j))
              for j, c in enumerate((128, 64, 32))]
    return sd, tokens


def no_bottleneck(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, stride: int,
                  store_dtype: Optional[torch.dtype] = None, groups: int = 16) -> torch.Tensor:
    """NoBottleneck.forward (unet3D.py:56-73): pre-activation residual block; the residual branch is
    downsample(x) = WSconv1x1(relu(GN(x))) computed from the block input when present (:68-69, :643-649).
    ``store_dtype``: see ``stored`` (each tensor a kernel writes to HBM is rounded; the residual is added to the fp32
    accumulator before the one rounding of the block output)."""
    st = store_dtype

    def conv(a, key, s_, pad):
        return F.conv3d(a, stored(ws_weight(sd[key]), st, round_grad=False), None, s_, pad, 1, 1)

    out = stored(gn_relu(x, sd[p + "gn1.weight"], sd[p + "gn1.bias"], groups), st)
    out = stored(conv(out, p + "conv1.weight", stride, 1), st)
    out = stored(gn_relu(out, sd[p + "gn2.weight"], sd[p + "gn2.bias"], groups), st)
    out = conv(out, p + "conv2.weight", 1, 1)
    if (p + "downsample.2.weight") in sd:
        res = stored(gn_relu(x, sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"], groups), st)
        res = stored(conv(res, p + "downsample.2.weight", stride, 0), st)
    else:
        res = x
    return stored(out + res, st)


def upsample2x_add(x: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
    """nn.Upsample(scale_factor=2, mode='trilinear') (align_corners=False) then ``x + skip``
    (unet3D.py:608, :686-687)."""
    return F.interpolate(x, scale_factor=2, mode="trilinear") + skip


def unet3d_forward(sd: Dict[str, torch.Tensor], image: torch.Tensor, base: int = 32, feats: Optional[dict] = None,
                   store_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """unet3D_baseline.forward (unet3D.py:663-718), returning the logits [B, num_classes, D, H, W].  When ``feats``
    is a dict, the stage outputs are stored in it (debugging aid for gradient localisation).
    ``store_dtype=torch.bfloat16`` emulates the storage precision of the B200 bf16 path (see ``stored``); the default
    is the reference's fp32 arithmetic."""
    spec = {p: (cin, cout, s) for p, cin, cout, s in backbone_spec(base)}
    st = store_dtype

    def blk(x, p):
        return no_bottleneck(x, sd, p, spec[p][2], st)

    def conv(x, w, stride, padding):           # Conv3d.forward; the tensor cores read a bf16 copy of the weight
        return stored(F.conv3d(x, stored(ws_weight(w), st, round_grad=False), None, stride, padding, 1, 1), st)

    def act(x, gamma, beta):
        return stored(gn_relu(x, gamma, beta), st)

    def up(x, skip):
        return stored(upsample2x_add(x, skip), st)

    def keep(name, t):
        if feats is not None:
            feats[name] = t
        return t

    x = keep("stem", conv(image, sd["conv1.weight"], 1, 1))                            # :666
    x = keep("layer0", blk(x, "layer0.0."))
    skip0 = x                                                                         # :667-668
    x = keep("layer1", blk(blk(x, "layer1.0."), "layer1.1."))
    skip1 = x                                                                         # :670-671
    x = keep("layer2", blk(blk(x, "layer2.0."), "layer2.1."))
    skip2 = x
    x = keep("layer3", blk(blk(x, "layer3.0."), "layer3.1."))
    skip3 = x
    x = keep("layer4", blk(blk(x, "layer4.0."), "layer4.1."))                         # :679
    x = keep("fusion", conv(act(x, sd["fusionConv.0.weight"], sd["fusionConv.0.bias"]), sd["fusionConv.2.weight"], 1, 0))
    x = keep("x8", blk(keep("up8", up(x, skip3)), "x8_resb.0."))                      # :686-688
    x = keep("x4", blk(keep("up4", up(x, skip2)), "x4_resb.0."))
    x = keep("x2", blk(keep("up2", up(x, skip1)), "x2_resb.0."))
    x = keep("x1", blk(keep("up1", up(x, skip0)), "x1_resb.0."))                      # :707-709
    x = act(x, sd["precls_conv.0.weight"], sd["precls_conv.0.bias"])
    return F.conv3d(x, sd["precls_conv.2.weight"], sd["precls_conv.2.bias"])         # :629-633, :713 (fp32 logits)


def eam_attention_logits(sd: Dict[str, torch.Tensor], p: str, x_tokens: torch.Tensor, class_token: torch.Tensor,
                         heads: int = 4) -> torch.Tensor:
    """The part of EAM.forward (unet3D.py:190-212) the model uses: attn = q k^T per head (UNSCALED, :203), with
    k = kv(norm2(x))[..., :C], q = q(norm3(token)).  x_tokens [B, N, C], class_token [B, Nt, C] -> [B, heads, Nt, N]."""
    B, N, C = x_tokens.shape
    xn = F.layer_norm(x_tokens, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    tn = F.layer_norm(class_token, (C,), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    kv = F.linear(xn, sd[p + "kv.weight"]).reshape(B, N, 2, heads, C // heads).permute(2, 0, 3, 1, 4)
    q = F.linear(tn, sd[p + "q.weight"]).reshape(B, class_token.shape[1], heads, C // heads).permute(0, 2, 1, 3)
    return q @ kv[0].transpose(-2, -1)


def unet3d_feam3_forward(sd: Dict[str, torch.Tensor], tokens: Sequence[torch.Tensor], image: torch.Tensor):
    """unet3D_with_feam3.forward in train mode (unet3D.py:1095-1190, use_cm all True, deep_up False):
    -> (logits, atten_map[3], deep_map[3], feature_stored[3])."""
    spec = {p: (cin, cout, s) for p, cin, cout, s in backbone_spec(32)}

    def blk(x, p):
        return no_bottleneck(x, sd, p, spec[p][2])

    def head(x, p):        # GroupNorm -> ReLU -> 1x1x1 conv with bias (:969-993)
        return F.conv3d(gn_relu(x, sd[p + "0.weight"], sd[p + "0.bias"]), sd[p + "2.weight"], sd[p + "2.bias"])

    x = blk(ws_conv3d(image, sd["conv1.weight"], 1, 1), "layer0.0.")
    skip0 = x
    x = blk(blk(x, "layer1.0."), "layer1.1.")
    skip1 = x
    x = blk(blk(x, "layer2.0."), "layer2.1.")
    skip2 = x
    x = blk(blk(x, "layer3.0."), "layer3.1.")
    skip3 = x
    x = blk(blk(x, "layer4.0."), "layer4.1.")
    x = ws_conv3d(gn_relu(x, sd["fusionConv.0.weight"], sd["fusionConv.0.bias"]), sd["fusionConv.2.weight"], 1, 0)
    attn, deep, feats = [], [], []
    for resb, skip, dp, eam, tok in (("x8_resb.0.", skip3, "deepout1.", "eam84.", tokens[0]),
                                     ("x4_resb.0.", skip2, "deepout2.", "eam42.", tokens[1]),
                                     ("x2_resb.0.", skip1, "deepout3.", "eam21.", tokens[2])):
        x = blk(upsample2x_add(x, skip), resb)
        deep.append(head(x, dp))
        feats.append(x.detach().clone())
        B, C = x.shape[0], x.shape[1]
        x_t = x.view(B, C, -1).permute(0, 2, 1)                                        # :1131
        a = eam_attention_logits(sd, eam, x_t, tok.view(1, tok.shape[0], C).detach())
        attn.append(a.mean(1).reshape((B, tok.shape[0]) + tuple(x.shape[2:])))        # :1134
    x = blk(upsample2x_add(x, skip0), "x1_resb.0.")
    return head(x, "precls_conv."), attn, deep, feats


def get_loss_refine(output, deep_out, target, class_weight, attns, refine_output, label_t, confi=0.10, aux_weight=1.0,
                    weight_feature=0.1):
    """get_loss with a refiner output (losses.py:107-178): base partial-label term + deep-supervision terms +
    pseudo-label gated-Dice terms; single-sample batches (the reference indexes mask[0])."""
    num_classes = output.shape[1] - 1
    loss = partial_label_loss(output, target.squeeze(1), class_weight)
    aux = 0.0
    weights = [0.125, 0.25, 0.5, 1]
    for idx, l in enumerate(deep_out):
        ct = F.interpolate(target, l.shape[2:], mode="nearest").float()
        aux = aux + partial_label_loss(l, ct.squeeze(1), class_weight, uce=False) * weights[idx]
    rp = torch.softmax(refine_output, 1)
    cmask = torch.logical_or(rp > (1 - confi), rp < confi).float()
    sup = sum(1 for v in label_t if v)
    maps = list(attns) + [torch.softmax(output, 1)[:, 1:]]
    for idx, l in enumerate(maps):
        for gan in range(num_classes):
            if label_t[gan]:
                continue
            cd = binary_gated_dice(l[:, gan:gan + 1], rp[gan:gan + 1, 1], cmask[gan:gan + 1, 1:], uce=False,
                                   sigmoid=(idx != 3))
            aux = aux + cd / (num_classes - sup) * weights[idx] * weight_feature
    return loss + aux * aux_weight


# ----------------------------------------------------------------------------------------------------------------
# Partial-label loss                                                                      loss_partial.py:10-99
# ----------------------------------------------------------------------------------------------------------------


def dice_term(score: torch.Tensor, target01: torch.Tensor) -> torch.Tensor:
    """DiceLoss._dice_loss with an all-ones mask (loss_partial.py:24-36): pooled over batch and voxels,
    squared-score denominator, smooth 1e-5 top and bottom."""
    smooth = 1e-5
    inter = torch.sum(score * target01)
    y = torch.sum(target01 * target01)
    z = torch.sum(score * score)
    return 1 - (2 * inter + smooth) / (z + y + smooth)


def partial_label_loss(logits: torch.Tensor, target: torch.Tensor, class_weight: Sequence[float],
                       uce: bool = True) -> torch.Tensor:
    """EDiceLoss_partial.forward(inputs, target, mask=[w,...], soft_max=True, uce) (loss_partial.py:71-99).

    ``target`` is [B, D, H, W] float class ids; ``class_weight`` is ``mask[0]`` (the reference uses the first
    sample's vector for the whole batch, :87/:92).  dice = sum_c w_c * dice_c / C (:49-57); ce = sum_c w_c *
    BCELoss(p_c, [target == c]) with PyTorch's log clamp at -100 (:90-92)."""
    C = logits.shape[1]
    p = torch.softmax(logits, dim=1)
    dice = 0.0
    for c in range(C):                                   # DiceLoss.forward loop, loss_partial.py:49-56
        t = (target == c).float()
        dice = dice + dice_term(p[:, c], t) * float(class_weight[c])
    dice = dice / C
    if not uce:
        return dice
    ce = 0.0
    for c in range(C):                                   # loss_partial.py:91-92
        t = (target == c).float()
        ce = ce + F.binary_cross_entropy(p[:, c].float(), t) * float(class_weight[c])
    return dice + ce


def dice_loss_class(probs: torch.Tensor, target: torch.Tensor, weight: Optional[Sequence[float]] = None,
                    gate: Optional[torch.Tensor] = None) -> torch.Tensor:
    """DiceLoss.forward(inputs=probabilities, target, weight, softmax=False, mask=gate) (loss_partial.py:38-57):
    per-class ``_dice_loss`` over the voxels where ``gate[:, c]`` is true, weighted sum divided by the class count."""
    C = probs.shape[1]
    loss = 0.0
    for c in range(C):
        t = (target == c).float()
        sc = probs[:, c]
        if gate is not None:
            m = gate[:, c].bool()
            sc, t = sc[m], t[m]
        loss = loss + dice_term(sc, t) * (1.0 if weight is None else float(weight[c]))
    return loss / C


def binary_gated_dice(x: torch.Tensor, target: torch.Tensor, gate: Optional[torch.Tensor] = None, uce: bool = True,
                      sigmoid: bool = True) -> torch.Tensor:
    """EDiceLoss_full2.forward (loss_partial.py:150-170): Dice between sigmoid(x) (or x) and a soft target over the
    gated voxels, plus BCE-with-logits when ``uce``."""
    p = torch.sigmoid(x) if sigmoid else x
    if gate is None:
        gate = torch.ones_like(target).unsqueeze(0)
    t = target.float()
    m = gate.bool()
    sc = p[m]
    tt = t[gate.squeeze(1).bool()] if gate.dim() == t.dim() + 1 else t[m]
    dice = dice_term(sc, tt)
    if uce:
        return dice + F.binary_cross_entropy_with_logits(x.float().squeeze(0), t)
    return dice


def partial_label_loss_sums(logits: np.ndarray, target: np.ndarray) -> Dict[str, np.ndarray]:
    """float64 numpy restatement of the four per-class sums the fused kernel produces:
    I = sum p t, Z = sum p^2, Y = sum t, E = sum BCE terms (log clamped at -100)."""
    z = logits.astype(np.float64)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    p = e / e.sum(axis=1, keepdims=True)
    C = logits.shape[1]
    out = {k: np.zeros(C) for k in "IZYE"}
    for c in range(C):
        t = (target == c).astype(np.float64)
        pc = p[:, c]
        out["I"][c] = (pc * t).sum()
        out["Z"][c] = (pc * pc).sum()
        out["Y"][c] = t.sum()
        with np.errstate(divide="ignore"):
            lp = np.maximum(np.log(pc), -100.0)
            l1p = np.maximum(np.log1p(-pc), -100.0)
        out["E"][c] = -(t * lp + (1 - t) * l1p).sum()
    return out


def remap_unsupervised(labels: torch.Tensor, sup16: Sequence[float]) -> torch.Tensor:
    """cmask construction (train_amos_atlas_final.py:252-255): voxels whose organ id is not supervised for this
    volume are relabelled background.  ``sup16[l]`` is the supervision bit of class l (index 0 = background)."""
    out = labels.clone()
    for l in range(1, len(sup16)):
        if not sup16[l]:
            out[out == l] = 0
    return out


def read_supervise_mask(path: str, w_bg: float = 1.0) -> Dict[str, List[float]]:
    """supervise_mask.csv adapter: skip the ``name,mask`` header, strip ``.nii.gz``, and prepend
    the background slot the train loop expects (train_amos_atlas_final.py:177-183, :215-219)."""
    table: Dict[str, List[float]] = {}
    with open(path, "r") as f:
        for name, mask in csv.reader(f):
            if name == "name":
                continue
            bits = [float(v) for v in mask.strip("[] ").split(",")]
            table[name.replace(".nii.gz", "")] = [float(w_bg)] + bits
    return table


# ----------------------------------------------------------------------------------------------------------------
# Sliding-window inference + Dice                                                         evaluate_amos.py:92-279
# ----------------------------------------------------------------------------------------------------------------


def gaussian_importance(tile: Sequence[int], sigma_scale: float = 1.0 / 8) -> np.ndarray:
    """_get_gaussian (evaluate_amos.py:184-197): scipy gaussian_filter of a centred delta (sigma = tile/8,
    mode constant), divided by its max, cast to fp32, zeros replaced by the smallest non-zero value."""
    from scipy.ndimage import gaussian_filter

    tmp = np.zeros(tile)
    tmp[tuple(i // 2 for i in tile)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in tile], 0, mode="constant", cval=0)
    g = (g / np.max(g) * 1).astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def tile_starts(size: int, tile: int, stride: int) -> List[int]:
    """Window starts along one axis (evaluate_amos.py:218-239): ceil((S-t)/stride)+1 windows, each clamped so it
    ends inside the volume (the last ones slide back to the border)."""
    n = int(math.ceil((size - tile) / stride) + 1)
    out = []
    for i in range(n):
        a = int(i * stride)
        b = min(a + tile, size)
        out.append(max(int(b - tile), 0))
    return out


def tile_grid(image_size: Sequence[int], tile: Sequence[int]) -> List[Tuple[int, int, int]]:
    """All (d1, y1, x1) window origins in the reference's dep->row->col order (evaluate_amos.py:215-239).
    Note the H/W stride is derived from tile[1] only (:217)."""
    overlap = 1 / 4
    stride_hw = int(math.ceil(tile[1] * (1 - overlap)))
    stride_d = int(math.ceil(tile[0] * (1 - overlap)))
    ds = tile_starts(image_size[0], tile[0], stride_d)
    ys = tile_starts(image_size[1], tile[1], stride_hw)
    xs = tile_starts(image_size[2], tile[2], stride_hw)
    return [(d, y, x) for d in ds for y in ys for x in xs]


_TTA_FLIPS = ([2], [3], [4], [2, 3], [2, 4], [3, 4], [2, 3, 4])


def predict_sliding(net, image: np.ndarray, tile: Sequence[int], classes: int, tta: bool = False) -> torch.Tensor:
    """predict_sliding (evaluate_amos.py:211-279): Gaussian-weighted logit accumulation in float64, normalised by
    the accumulated weights.  ``net(img)`` maps a [B,1,d,h,w] fp32 tensor to logits.  ``tta``: mean over the identity
    and the seven axis-flip combinations, each prediction flipped back (:247-255)."""
    g = torch.from_numpy(gaussian_importance(tile))
    B, _, D, H, W = image.shape
    full = torch.zeros((B, classes, D, H, W), dtype=torch.float64)
    count = torch.zeros((B, classes, D, H, W), dtype=torch.float64)
    for d1, y1, x1 in tile_grid((D, H, W), tile):
        d2, y2, x2 = d1 + tile[0], y1 + tile[1], x1 + tile[2]
        img = torch.from_numpy(image[:, :, d1:d2, y1:y2, x1:x2])
        pred = net(img).float().cpu()
        if tta:
            for dims in _TTA_FLIPS:
                pred = pred + torch.flip(net(torch.flip(img, dims)).float().cpu(), dims)
            pred = pred / 8.
        pred = pred * g
        count[:, :, d1:d2, y1:y2, x1:x2] += g
        full[:, :, d1:d2, y1:y2, x1:x2] += pred
    return full / count


def get_dice(preds: torch.Tensor, labels: torch.Tensor, num_class: int = 13):
    """get_dice without atlas (evaluate_amos.py:128-141) on dice/senc/spec_score (:92-126): argmax of the softmax,
    then per class l=1..num_class 2|P&T| / (|P|+|T|+1), |P&T|/(|T|+1), |P&T|/(|P|+1), each averaged over batch."""
    pr = F.softmax(preds, dim=1)
    am = torch.argmax(pr, dim=1)
    lab = labels.reshape(labels.shape[0], -1)
    dices, senc, spec = [], [], []
    for l in range(1, num_class + 1):
        p = (am == l).reshape(am.shape[0], -1).double()
        t = (lab == l).double()
        num = (p * t).sum(1)
        dices.append((2 * num / (p.sum(1) + t.sum(1) + 1)).mean())
        senc.append((num / (t.sum(1) + 1)).mean())
        spec.append((num / (p.sum(1) + 1)).mean())
    return dices, senc, spec, am


# ----------------------------------------------------------------------------------------------------------------
# Optimiser                                                          train_amos_atlas_final.py:132-135, utils.py:53-60
# ----------------------------------------------------------------------------------------------------------------


def sgd_step(p: torch.Tensor, g: torch.Tensor, buf: Optional[torch.Tensor], lr: float, momentum: float = 0.9,
             weight_decay: float = 1e-4) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.optim.SGD(momentum=0.9, weight_decay=1e-4, dampening=0, nesterov=False): d = g + wd*p;
    buf = d on the first step else momentum*buf + d; p -= lr*buf."""
    d = g + weight_decay * p
    buf = d.clone() if buf is None else momentum * buf + d
    return p - lr * buf, buf


def lr_poly(base_lr: float, it: int, max_iter: int, power: float = 0.9) -> float:
    """utils.py:53-54."""
    return base_lr * ((1 - float(it) / max_iter) ** power)


# ----------------------------------------------------------------------------------------------------------------
# Synthetic inputs
# ----------------------------------------------------------------------------------------------------------------


def synth_patch(shape: Sequence[int], seed: int, modality: str = "ct") -> torch.Tensor:
    """CT: clip(N(0, .5), -1, 1) (mimics the +-325 HU window, MOTSDataset.py:171-183); MRI: z-scored N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(tuple(shape), generator=g)
    return (0.5 * x).clamp_(-1, 1) if modality == "ct" else x


def synth_labels(shape: Sequence[int], seed: int, num_classes: int = 16, n_seeds: int = 48) -> torch.Tensor:
    """Piecewise-constant blobs: nearest-seed Voronoi over [B, D, H, W], about half of the seeds background, every
    class present when n_seeds >= 2*num_classes.  Returned as float class ids [B,1,D,H,W] like the reference's
    label tensors (train_amos_atlas_final.py:214)."""
    B, D, H, W = shape
    rng = np.random.RandomState(seed)
    out = np.zeros((B, D, H, W), dtype=np.float32)
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    for b in range(B):
        pts = rng.rand(n_seeds, 3) * np.array([D, H, W])
        cls = np.where(np.arange(n_seeds) % 2 == 0, 0, (np.arange(n_seeds) // 2) % (num_classes - 1) + 1)
        best = np.full((D, H, W), np.inf)
        lab = np.zeros((D, H, W), dtype=np.float32)
        for (pz, py, px), c in zip(pts, cls):
            d2 = ((zz - pz) * 3.0) ** 2 + (yy - py) ** 2 + (xx - px) ** 2
            m = d2 < best
            best[m] = d2[m]
            lab[m] = c
        out[b] = lab
    return torch.from_numpy(out).unsqueeze(1)


# ----------------------------------------------------------------------------------------------------------------
# Input pipeline                                             MOTSDataset.py:33-52, :171-186, :269-297, :299-395
# ----------------------------------------------------------------------------------------------------------------
# The reference's dataset needs SimpleITK, nibabel and batchgenerators (none installed here), so these are
# restatements: numpy statements copied in meaning from __getitem__ for the deterministic part, and the PUBLISHED algorithms
# of batchgenerators' transforms (MIC-DKFZ/batchgenerators, the version the reference imports is unpinned -- no
# requirements file) for the augmentations, anchored on the call site get_train_transform (:33-52).  "parity unpinned" for
# the augmentations: the reference holds no fixture for them.


def pad_image_ref(img: np.ndarray, target) -> np.ndarray:
    """pad_image (MOTSDataset.py:269-282): zero-pad at the end of each axis up to ``target``."""
    miss = [max(int(math.ceil(t - s)), 0) for t, s in zip(target, img.shape[-3:])]
    pad = [(0, 0)] * (img.ndim - 3) + [(0, m) for m in miss]
    return np.pad(img, pad, "constant")


def truncate_ref(ct: np.ndarray, is_ct: bool) -> np.ndarray:
    """truncate (MOTSDataset.py:171-186): CT clipped to [-325, 325] HU and divided by 325; MRI z-scored."""
    ct = ct.copy()
    if is_ct:
        ct[np.where(ct <= -325)] = -325
        ct[np.where(ct >= 325)] = 325
        return (ct - 0) / 325.
    ct = ct - np.mean(ct)
    return ct / np.std(ct)


def prepare_patch_ref(image: np.ndarray, label: np.ndarray, atlas: Optional[np.ndarray], is_ct: bool, crop, origin):
    """AMOSDataSet_newatlas.__getitem__ from the atlas resize to the final cast (MOTSDataset.py:357-394) with the crop
    origin (b, c, a) given instead of drawn.  image / label [h, w, d]; -> image [1,D,H,W] f32, label [1,D,H,W] f32,
    catlas [K,D,H,W] f32 (or None)."""
    ch, cw, cd = crop
    b, c, a = origin
    catlas = None
    if atlas is not None:
        catlas = F.interpolate(torch.tensor(atlas).unsqueeze(0), image.shape).numpy()[0]                  # :357
    if image.shape != label.shape:                                                                       # :360-368
        fs = [min(image.shape[i], label.shape[i]) for i in range(3)]
        image, label = image[:fs[0], :fs[1], :fs[2]], label[:fs[0], :fs[1], :fs[2]]
    tgt = [ch + 5, cw + 5, cd + 5]
    image, label = pad_image_ref(image, tgt), pad_image_ref(label, tgt)                                   # :370-371
    if catlas is not None:
        catlas = pad_image_ref(catlas, tgt)                                                               # :372
    image = truncate_ref(image, is_ct)                                                                   # :374
    image = image[b:b + ch, c:c + cw, a:a + cd]                                                           # :381-383
    label = label[b:b + ch, c:c + cw, a:a + cd]
    image = image[np.newaxis].transpose((0, 3, 1, 2)).astype(np.float32)                                  # :386-392
    label = label[np.newaxis].transpose((0, 3, 1, 2)).astype(np.float32)
    if catlas is not None:
        catlas = catlas[:, b:b + ch, c:c + cw, a:a + cd].transpose((0, 3, 1, 2)).astype(np.float32)
    return image, label, catlas


def augment_ref(img: np.ndarray, params: dict, noise: Optional[np.ndarray] = None) -> np.ndarray:
    """get_train_transform (MOTSDataset.py:33-52) with the draws given: GaussianNoise (``noise`` = the unit normal field to
    scale by noise_std), GaussianBlur (scipy gaussian_filter, order 0), multiplicative and additive brightness, contrast
    with preserve_range.  img [1, D, H, W] float32."""
    from scipy.ndimage import gaussian_filter

    x = img.astype(np.float32).copy()
    if params.get("noise_std") and noise is not None:
        x = x + np.float32(params["noise_std"]) * noise.astype(np.float32)
    if params.get("blur_sigma"):
        x[0] = gaussian_filter(x[0], params["blur_sigma"], order=0)
    if params.get("mult") is not None:
        x = x * np.float32(params["mult"])
    if params.get("add") is not None:
        x = x + np.float32(params["add"])
    if params.get("contrast") is not None:
        mn, lo, hi = x.mean(), x.min(), x.max()
        x = (x - mn) * np.float32(params["contrast"]) + mn
        x[x < lo] = lo
        x[x > hi] = hi
    return x.astype(np.float32)


def synth_named_state(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic weights for ANY module, keyed by state_dict name (so the unmodified reference module and the B200
    module get identical values): 1-D "...weight" tensors (norm scales) ~ 1 + 0.1 N(0,1), biases ~ 0.1 N(0,1),
    everything else ~ N(0,1) / sqrt(fan_in)."""
    out = {}
    for i, (k, shp) in enumerate(sorted(shapes.items())):
        g = torch.Generator().manual_seed(seed * 100003 + 9000 + i)
        if len(shp) == 1 and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            t = torch.randn(shp, generator=g) / math.sqrt(int(np.prod(shp[1:])))
        out[k] = t.float()
    return out
