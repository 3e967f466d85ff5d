"""Drop-in for the reference's ``unet3D.py`` backbone, running on the B200 kernels of libmmpl_b200.so.

Same class names, constructor signatures, ``state_dict`` keys/shapes and call signatures as the reference
(TThuraya/multimodal-PL ``unet3D.py``): ``Conv3d`` (:16-27), ``conv3x3x3`` (:30-35), ``NoBottleneck`` (:40-73),
``unet3D_baseline`` (:584-718).  Checkpoints of the reference load unchanged.

Differences that are deliberate:
  * tensors between layers are logical NCDHW but stored channels-last (NDHWC) in the compute dtype
    (``ops.set_compute_dtype``: bf16 -> tcgen05 kernels, fp32 -> exact CUDA-core kernels);
  * GroupNorm+ReLU, the residual add, weight standardisation and up-sample+skip are fused kernels;
  * ``base`` (init width, 32 in the reference) is a keyword so the wide stress config (BASELINE configs[4]) can be
    built from the same recipe.
There is no CPU path: calling a module without a CUDA device raises.
"""
import torch
import torch.nn as nn

from . import ops

in_place = True
affine_par = True


def _triple(v):
    if isinstance(v, (tuple, list)):
        assert len(v) == 3 and v[0] == v[1] == v[2], f"anisotropic value {v} is not supported"
        return int(v[0])
    return int(v)


import os as _os

_COMPACT_DS = _os.environ.get("MMPL_COMPACT_DS", "1") != "0"     # compact second GroupNorm head for 1x1x1 s2 downsamples


class Conv3d(nn.Conv3d):
    """Weight-standardised convolution (reference unet3D.py:16-27).  Supports what the backbone uses: kernel 1 or 3,
    padding k//2, stride 1 or 2, dilation 1, groups 1, no bias."""

    _standardise = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1),
                 groups=1, bias=False):
        super(Conv3d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        k, s, p, d = _triple(self.kernel_size), _triple(self.stride), _triple(self.padding), _triple(self.dilation)
        if k not in (1, 3) or s not in (1, 2) or p != k // 2 or d != 1 or groups != 1 or self.bias is not None:
            raise NotImplementedError(
                f"B200 conv path covers k in {{1,3}}, stride in {{1,2}}, padding k//2, dilation 1, groups 1, bias=False; "
                f"got k={k} stride={s} padding={p} dilation={d} groups={groups} bias={self.bias is not None}")
        self._k, self._s = k, s

    def forward(self, x, residual=None, want_stats=False):
        if self.in_channels == 1:
            assert residual is None
            return ops.stem_conv(x, self.weight, self._standardise)
        return ops.ws_conv3d(x, self.weight, self._s, self._standardise, residual, want_stats)


class PlainConv3d(Conv3d):
    """nn.Conv3d without weight standardisation (what conv3x3x3(weight_std=False) returns, unet3D.py:35)."""

    _standardise = False


def conv3x3x3(in_planes, out_planes, kernel_size=(3, 3, 3), stride=(1, 1, 1), padding=1, dilation=1, bias=False,
              weight_std=False):
    "3x3x3 convolution with padding"
    cls = Conv3d if weight_std else PlainConv3d
    return cls(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation,
               bias=bias)


class GNReLUConv(nn.Sequential):
    """Sequential(GroupNorm, ReLU, conv) as used for ``downsample`` (:643-649) and ``fusionConv`` (:602-606); the
    children keep the reference's indices (0 = GroupNorm, 2 = conv) so state_dict keys are unchanged."""

    def forward(self, x, residual=None):
        gn, conv = self[0], self[2]
        a = ops.gn_relu(x, gn.weight, gn.bias, gn.num_groups, gn.eps)
        return conv(a, residual) if residual is not None else conv(a)


class GNReLUClassifier(nn.Sequential):
    """precls_conv = Sequential(GroupNorm(16, base), ReLU, nn.Conv3d(base, classes, 1)) (:629-633): plain 1x1x1
    convolution WITH bias and without weight standardisation, emitting fp32 NCDHW logits.  Also the deep-supervision
    heads deepout1..3 of unet3D_with_feam3 (:969-993; 128 / 64 / 32 channels)."""

    def forward(self, x, blend=None):
        gn, conv = self[0], self[2]
        a = ops.gn_relu(x, gn.weight, gn.bias, gn.num_groups, gn.eps)
        if blend is not None:       # sliding-window inference: accumulate g * logits straight into the volume
            ops.classifier_blend(a, conv.weight, conv.bias, blend)
            return None
        return ops.classifier(a, conv.weight, conv.bias)

    def partial_loss(self, x, target, class_weight, lut=None, uce=True, per_sample=False):
        """EDiceLoss_partial of this head's logits (loss_partial.py:71-99) without writing them: GroupNorm+ReLU, then the
        fused classifier + loss kernels (ops.classifier_partial_loss)."""
        gn, conv = self[0], self[2]
        a = ops.gn_relu(x, gn.weight, gn.bias, gn.num_groups, gn.eps)
        return ops.classifier_partial_loss(a, conv.weight, conv.bias, target, class_weight, lut, uce, per_sample)


class NoBottleneck(nn.Module):
    """Pre-activation residual block (reference unet3D.py:40-73):
    out = conv2(relu(gn2(conv1(relu(gn1(x)))))) + (downsample(x) if downsample is not None else x)."""

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, fist_dilation=1, multi_grid=1,
                 weight_std=False, group=16):
        super(NoBottleneck, self).__init__()
        self.weight_std = weight_std
        self.gn1 = nn.GroupNorm(group, inplanes)
        self.conv1 = conv3x3x3(inplanes, planes, kernel_size=(3, 3, 3), stride=stride, padding=(1, 1, 1),
                               dilation=dilation * multi_grid, bias=False, weight_std=self.weight_std)
        self.relu = nn.ReLU(inplace=in_place)
        self.gn2 = nn.GroupNorm(group, planes)
        self.conv2 = conv3x3x3(planes, planes, kernel_size=(3, 3, 3), stride=1, padding=(1, 1, 1),
                               dilation=dilation * multi_grid, bias=False, weight_std=self.weight_std)
        self.downsample = downsample
        self.dilation = dilation
        self.stride = stride

    def forward(self, x, want_alias=False):
        """``want_alias=True`` -> (out, x_alias): x_alias is the block input routed through gn1's autograd node, so a
        second consumer of the block input (the encoder skip, unet3D.py:669-677) has its gradient added inside the
        GroupNorm backward kernel instead of by a separate elementwise add."""
        ds = self.downsample
        fused_ds = isinstance(ds, GNReLUConv) and ds[0].num_groups == self.gn1.num_groups and ds[0].eps == self.gn1.eps
        x_alias = None
        if fused_ds:
            # gn1 and downsample.0 normalise the same tensor: one statistics pass, one read, two affine heads
            # a 1x1x1 stride-2 downsample reads the even voxels of its input only: that head is produced compact and the
            # convolution becomes a stride-1 one on it (same arithmetic, 7/8 of the head's traffic gone, both directions)
            dconv = ds[2]
            compact = isinstance(dconv, Conv3d) and dconv._k == 1 and dconv._s == 2 and _COMPACT_DS
            c1 = self.conv1
            psplit = (isinstance(c1, Conv3d) and x.is_cuda
                      and ops.psplit_consumer(c1.in_channels, c1.out_channels, c1._k, c1._s))
            outs = ops.gn_relu_dual(x, self.gn1.weight, self.gn1.bias, ds[0].weight, ds[0].bias,
                                    self.gn1.num_groups, self.gn1.eps, alias=want_alias, compact2=compact, psplit1=psplit)
            a1, ads = outs[0], outs[1]
            if want_alias:
                x_alias = outs[2]
            residual = ops.ws_conv3d(ads, dconv.weight, 1, dconv._standardise) if compact else dconv(ads)
        elif ds is None:
            # identity residual: its gradient is folded into gn1's backward through the alias output
            a1, x_alias = ops.gn_relu(x, self.gn1.weight, self.gn1.bias, self.gn1.num_groups, self.gn1.eps, alias=True)
            residual = x_alias
        else:
            a1 = ops.gn_relu(x, self.gn1.weight, self.gn1.bias, self.gn1.num_groups, self.gn1.eps)
            residual = ds(x)
        fuse = self.gn2.num_groups == 16
        out = self.conv1(a1, want_stats=fuse)    # GroupNorm statistics of conv1's output come from its epilogue
        a2 = ops.gn_relu(out, self.gn2.weight, self.gn2.bias, self.gn2.num_groups, self.gn2.eps)
        # residual add fused into conv2's epilogue; the block output usually feeds the next block's GroupNorm
        out = self.conv2(a2, residual, want_stats=fuse)
        if want_alias:
            return out, (x_alias if x_alias is not None else x)
        return out


class _Upsample2xAdd(nn.Upsample):
    """nn.Upsample(scale_factor=2, mode='trilinear'); ``forward(x, skip)`` fuses the additive skip (:686-687)."""

    def forward(self, x, skip=None):
        if skip is None:
            skip = torch.zeros((x.shape[0], x.shape[1], 2 * x.shape[2], 2 * x.shape[3], 2 * x.shape[4]),
                               dtype=x.dtype, device=x.device)
        return ops.upsample2x_add(x, skip)


class unet3D_baseline(nn.Module):
    """Reference unet3D.py:584-718.  forward(input, mask=None) -> (logits, [], []) in train mode, logits in eval."""

    def __init__(self, layers, num_classes=12, weight_std=False, ema=False, use_cm=[True, True, True], deep_up=False,
                 base=32):
        self.inplanes = 128
        self.weight_std = weight_std
        self.num_classes = num_classes
        self.use_cm = use_cm
        self.alpha = 0.01
        self.deep_up = deep_up
        super(unet3D_baseline, self).__init__()
        b = base
        self.conv1 = conv3x3x3(1, b, stride=[1, 1, 1], weight_std=self.weight_std)
        self.layer0 = self._make_layer(NoBottleneck, b, b, layers[0], stride=(1, 1, 1))
        self.layer1 = self._make_layer(NoBottleneck, b, 2 * b, layers[1], stride=(2, 2, 2))
        self.layer2 = self._make_layer(NoBottleneck, 2 * b, 4 * b, layers[2], stride=(2, 2, 2))
        self.layer3 = self._make_layer(NoBottleneck, 4 * b, 8 * b, layers[3], stride=(2, 2, 2))
        self.layer4 = self._make_layer(NoBottleneck, 8 * b, 8 * b, layers[4], stride=(2, 2, 2))
        self.fusionConv = GNReLUConv(
            nn.GroupNorm(16, 8 * b),
            nn.ReLU(inplace=in_place),
            conv3x3x3(8 * b, 8 * b, kernel_size=(1, 1, 1), padding=(0, 0, 0), weight_std=self.weight_std))
        self.upsamplex2 = _Upsample2xAdd(scale_factor=2, mode='trilinear')
        self.x8_resb = self._make_layer(NoBottleneck, 8 * b, 4 * b, 1, stride=(1, 1, 1))
        self.x4_resb = self._make_layer(NoBottleneck, 4 * b, 2 * b, 1, stride=(1, 1, 1))
        self.x2_resb = self._make_layer(NoBottleneck, 2 * b, b, 1, stride=(1, 1, 1))
        self.x1_resb = self._make_layer(NoBottleneck, b, b, 1, stride=(1, 1, 1))
        self.precls_conv = GNReLUClassifier(
            nn.GroupNorm(16, b),
            nn.ReLU(inplace=in_place),
            nn.Conv3d(b, num_classes, kernel_size=1))
        if ema:
            for param in self.parameters():
                param.detach_()

    def _make_layer(self, block, inplanes, planes, blocks, stride=(1, 1, 1), dilation=1, multi_grid=1):
        downsample = None
        if stride[0] != 1 or stride[1] != 1 or stride[2] != 1 or inplanes != planes:
            downsample = GNReLUConv(
                nn.GroupNorm(16, inplanes),
                nn.ReLU(inplace=in_place),
                conv3x3x3(inplanes, planes, kernel_size=(1, 1, 1), stride=stride, padding=0,
                          weight_std=self.weight_std))
        layers = [block(inplanes, planes, stride, dilation=dilation, downsample=downsample, multi_grid=1,
                        weight_std=self.weight_std)]
        for _ in range(1, blocks):
            layers.append(block(planes, planes, dilation=dilation, multi_grid=1, weight_std=self.weight_std))
        return nn.Sequential(*layers)

    @staticmethod
    def _stage_with_skip(layer, x):
        """Run an encoder stage; also return its input as routed through the first block (see NoBottleneck.forward)."""
        blocks = list(layer)
        if not blocks or not isinstance(blocks[0], NoBottleneck):
            return layer(x), x
        out, skip = blocks[0](x, want_alias=True)
        for blk in blocks[1:]:
            out = blk(out)
        return out, skip

    def _ws_convs(self):
        convs = getattr(self, "_ws_conv_list", None)
        if convs is None:
            convs = [m for m in self.modules() if isinstance(m, Conv3d)]
            object.__setattr__(self, "_ws_conv_list", convs)
        return [(m.weight, m._standardise, m.in_channels == 1) for m in convs]

    def blend_supported(self):
        """True if ``blend_tile`` can run (bf16 compute dtype, classifier width the fused kernel covers)."""
        conv = self.precls_conv[2]
        return ops.classifier_blend_supported(conv.in_channels, conv.out_channels, ops.get_compute_dtype())

    def _features(self, input):
        """Everything up to (not including) precls_conv: reference :666-709."""
        # one launch standardises + packs the weights of every convolution (the reference does it per Conv3d.forward,
        # unet3D.py:22-26); one zero-fill serves all GroupNorm statistics accumulators
        ops.prepare_ws(self._ws_convs())
        ops.begin_forward(input.device)
        x = self.conv1(input)
        x = self.layer0(x)
        x, skip0 = self._stage_with_skip(self.layer1, x)     # skipN = the input of stage N+1 (reference :669-677)
        x, skip1 = self._stage_with_skip(self.layer2, x)
        x, skip2 = self._stage_with_skip(self.layer3, x)
        x, skip3 = self._stage_with_skip(self.layer4, x)
        x = self.fusionConv(x)
        x = self.x8_resb(self.upsamplex2(x, skip3))
        x = self.x4_resb(self.upsamplex2(x, skip2))
        x = self.x2_resb(self.upsamplex2(x, skip1))
        return self.x1_resb(self.upsamplex2(x, skip0))

    def forward(self, input, mask=None):
        logits = self.precls_conv(self._features(input))
        if self.training:
            return logits, [], []
        return logits

    def forward_partial_loss(self, input, target, mask=None, lut=None, per_sample=False, uce=True):
        """``EDiceLoss_partial(C)(self(input)[0], target, mask, lut=..., per_sample=...)`` in one call (not part of the
        reference surface): the training step of train_amos_atlas_final.py:226-262 when only the loss of the main head
        is needed.  The fp32 logits and their gradient (2 x 64 B/voxel, four passes over them) are never written: the
        classifier runs inside the loss kernels, forward and backward.  Same value and gradients as the two-step form."""
        from .loss_functions.loss_partial import _class_weight

        conv = self.precls_conv[2]
        w = _class_weight(mask, conv.out_channels, input.device, per_sample)
        return self.precls_conv.partial_loss(self._features(input), target, w, lut, uce, per_sample)

    @torch.no_grad()
    def blend_tile(self, input, sink):
        """One sliding-window step (not part of the reference surface; used by evaluate.predict_sliding_dice): the
        logits of the tile are not returned but Gaussian-weighted and accumulated into the volume accumulator of
        ``sink`` (an ``ops.BlendSink``) by the classifier kernel itself (predict_sliding, evaluate_amos.py:244-276)."""
        self.precls_conv(self._features(input), blend=sink)

    @torch.no_grad()
    def blend_features(self, input):
        """First half of ``blend_tile``: the classifier's input a = relu(gn(features)) of a batch of tiles."""
        gn = self.precls_conv[0]
        return ops.gn_relu(self._features(input), gn.weight, gn.bias, gn.num_groups, gn.eps)

    @torch.no_grad()
    def blend_accumulate(self, a, sink):
        """Second half of ``blend_tile``: classifier + Gaussian-weighted accumulation of ``a`` into ``sink`` (so that the
        network of the next tile batch can run on another stream while this one is accumulated, in tile order)."""
        conv = self.precls_conv[2]
        ops.classifier_blend(a, conv.weight, conv.bias, sink)


class EAM(nn.Module):
    """Class-token cross attention (reference unet3D.py:142-212): same constructor, parameters (``kv``, ``q``, ``proj``,
    ``norm2``, ``norm3``) and ``forward(x, modality_token) -> (tokens_out, attn)`` contract, ``attn`` being the UNSCALED
    per-head q.k^T logits [B, heads, Nt, N].

    The model only consumes ``attn.mean(1)`` (:1133-1137).  ``attention_map`` computes exactly that on the device kernels
    without ever forming keys, values or per-head logits: the head mean of per-head dot products is one dot product over
    all channels, so with M = q(norm3(token)) @ Wk the map is a 15-row "classifier" (M * gamma2 / heads, bias M beta2 /
    heads) over the LayerNorm-ed voxel rows -- ``ops.layer_norm_rows`` + ``ops.classifier`` (csrc/eam.cu)."""

    def __init__(self, dim, input_resolution, num_heads, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, upsample=None, use_checkpoint=False):
        super(EAM, self).__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.kv = nn.Linear(dim, dim * 2, bias=False)
        self.q = nn.Linear(dim, dim, bias=False)
        self.softmax = nn.Softmax(dim=-1)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = norm_layer(dim)
        self.norm3 = norm_layer(dim)

    def _folded(self, token):
        """-> (W [Nt, C], b [Nt]) with  attn.mean(1)[t, v] = W[t] . xhat[v] + b[t],  xhat = LayerNorm2 without affine."""
        qt = self.q(self.norm3(token.float()))                         # [Nt, C]
        m = qt @ self.kv.weight[: self.dim].float()                    # keys are the first C rows of kv (:198-199)
        return m * self.norm2.weight.float() / self.num_heads, (m @ self.norm2.bias.float()) / self.num_heads

    def attention_map(self, x, token):
        """x [B, C, D, H, W] feature volume, token [Nt, C] -> head-mean attention logits [B, Nt, D, H, W] (fp32)."""
        w, b = self._folded(token)
        xhat = ops.layer_norm_rows(x, self.norm2.eps)
        return ops.classifier(xhat, w.reshape(w.shape[0], w.shape[1], 1, 1, 1), b)

    def forward(self, x, modality_token):
        """The reference's full contract, for callers that want the updated tokens or per-head logits; plain tensor
        algebra (the model's hot path is ``attention_map``)."""
        batch, n_vox, c = x.shape
        n_tok, heads, hd = modality_token.shape[1], self.num_heads, c // self.num_heads
        feats = self.norm2(x.float())
        keys, values = self.kv(feats).view(batch, n_vox, 2, heads, hd).unbind(2)           # each [B, N, heads, hd]
        queries = self.q(self.norm3(modality_token.float())).view(-1, n_tok, heads, hd)    # token batch is 1 in the model
        attn = torch.einsum("bthd,bnhd->bhtn", queries.expand(batch, -1, -1, -1), keys)
        mixed = torch.einsum("bhtn,bnhd->bthd", self.softmax(attn * self.scale), values).reshape(batch, n_tok, c)
        return self.proj(self.norm2(mixed)) + mixed, attn


class unet3D_with_feam3(unet3D_baseline):
    """Reference unet3D.py:938-1190: the baseline backbone plus three deep-supervision heads (GN-ReLU-1x1x1, :969-993),
    three class-token attention modules (EAM) and EMA class tokens.  ``forward(input, mask=None)`` returns
    ``(logits, atten_map[3], deep_map[3], feature_stored[3])`` in train mode and ``logits`` in eval mode;
    ``renew_token(features, mask)`` is the EMA update of :1051-1068.  Same ``state_dict`` keys as the reference."""

    def __init__(self, layers, num_classes=12, weight_std=False, ema=False, use_cm=[True, True, True], deep_up=False):
        super(unet3D_with_feam3, self).__init__(layers, num_classes=num_classes, weight_std=weight_std, ema=False,
                                                use_cm=use_cm, deep_up=deep_up, base=32)
        self.upsamplex3 = nn.Upsample(scale_factor=4, mode='trilinear')
        self.upsamplex4 = nn.Upsample(scale_factor=8, mode='trilinear')
        self.deepout1 = GNReLUClassifier(nn.GroupNorm(16, 128), nn.ReLU(inplace=in_place),
                                         nn.Conv3d(128, num_classes, kernel_size=1))
        self.eam84 = EAM(128, input_resolution=None, num_heads=4)
        self.deepout2 = GNReLUClassifier(nn.GroupNorm(16, 64), nn.ReLU(inplace=in_place),
                                         nn.Conv3d(64, num_classes, kernel_size=1))
        self.eam42 = EAM(64, input_resolution=None, num_heads=4)
        self.deepout3 = GNReLUClassifier(nn.GroupNorm(16, 32), nn.ReLU(inplace=in_place),
                                         nn.Conv3d(32, num_classes, kernel_size=1))
        self.eam21 = EAM(32, input_resolution=None, num_heads=4)
        # class tokens are plain tensors in the reference (not parameters, not in the state_dict), :1006-1011
        self.class_token1 = torch.randn(num_classes - 1, 128)
        self.class_token2 = torch.randn(num_classes - 1, 64)
        self.class_token3 = torch.randn(num_classes - 1, 32)
        if ema:
            for param in self.parameters():
                param.detach_()

    def _attend(self, eam, token, x, up):
        """attention map of one decoder scale (:1131-1137): mean over heads of the q.k^T logits, as a volume"""
        amap = eam.attention_map(x, token.detach())
        return up(amap) if self.deep_up else amap

    def forward(self, input, mask=None):
        dev = input.device
        self.class_token1 = self.class_token1.to(dev)
        self.class_token2 = self.class_token2.to(dev)
        self.class_token3 = self.class_token3.to(dev)
        atten_map, deep_map, feature_stored = [], [], []
        ops.prepare_ws(self._ws_convs())
        ops.begin_forward(dev)
        x = self.conv1(input)
        x = self.layer0(x)
        x, skip0 = self._stage_with_skip(self.layer1, x)
        x, skip1 = self._stage_with_skip(self.layer2, x)
        x, skip2 = self._stage_with_skip(self.layer3, x)
        x, skip3 = self._stage_with_skip(self.layer4, x)
        x = self.fusionConv(x)
        stages = [(self.x8_resb, skip3, self.deepout1, self.eam84, "class_token1", self.upsamplex4, 0),
                  (self.x4_resb, skip2, self.deepout2, self.eam42, "class_token2", self.upsamplex3, 1),
                  (self.x2_resb, skip1, self.deepout3, self.eam21, "class_token3", nn.Upsample(scale_factor=2, mode='trilinear'), 2)]
        for resb, skip, deepout, eam, tok, up, i in stages:
            x = resb(self.upsamplex2(x, skip))
            deep_map.append(deepout(x))
            feature_stored.append(x.detach().clone())
            if self.use_cm[i]:
                atten_map.append(self._attend(eam, getattr(self, tok), x, up))
        x = self.x1_resb(self.upsamplex2(x, skip0))
        logits = self.precls_conv(x)
        if self.training:
            return logits, atten_map, deep_map, feature_stored
        return logits

    @torch.no_grad()
    def renew_token(self, features, mask):
        """EMA of the per-class mean feature into the class tokens (:1051-1068) on the device (ops.renew_tokens: one
        segmented-sum launch + one update launch per level, no host synchronisation): classes absent from ``mask`` (or
        from its nearest-neighbour down-sampling) keep their token.  Per-channel means pool over the batch (the
        reference's reshape is only meaningful for batch 1)."""
        names = ["class_token1", "class_token2", "class_token3"]
        for index, x in enumerate(features):
            name = names[min(index, 2)]
            tok = getattr(self, name).to(x.device, torch.float32).contiguous()
            ops.renew_tokens(tok, x, mask, self.alpha)
            setattr(self, name, tok)
