"""The other two networks of the reference's train loop on the B200 kernels: the refiner ``unet3D_g``
(reference unet3D.py:1507-1623, built at train_amos_atlas_final.py:120 as ``unet3D_g([1,1,1,1,1], num_classes=2,
init_filter=24, in_channel=2)``) and the discriminator ``norm_style_discriminator_output`` (:1907-1947, :124).  Same class
names, constructor signatures, ``forward`` signatures and ``state_dict`` keys / shapes as the reference.

How their layers reach the tensor-core kernels (csrc/aux_nets.cu has the kernel-side notes):

* refiner widths 24 / 48 / 96 / 192 are not tcgen05 tile widths.  Activations are carried zero-PADDED to 32 / 64 / 128 / 256
  channels with every GroupNorm group padded in place (group g of a 24-channel tensor occupies padded channels
  8g .. 8g+5), so the block GroupNorms (``group=4``, :1540-1559) normalise the right channels; the kernels divide by the
  real element count (``real_cpg``).  Weights keep the reference shapes: each forward standardises them with a few tiny
  tensor ops (what the reference does per Conv3d.forward, :22-26), scatters them into the padded layout and hands them to
  the convolution kernels as plain weights -- gradients flow back through the scatter to the reference-shaped parameters.
  Padded channels stay exactly zero through the whole network (zero weights, zero GroupNorm affine).
* the discriminator's ``nn.Conv3d(k=4, stride=2, padding=1)`` layers run as 3x3x3 stride-1 convolutions over a
  space-to-depth(2) copy of their input (8C channels), the 4^3 taps scattered into the 3^3 x 8C filter; bias + LeakyReLU
  is one fused kernel.

Neither network is on the north-star path; they are here so that the whole loop body of train_amos_atlas_final.py:258-378
can run on the device library.  bf16 compute dtype = tensor cores; fp32 = the exact CUDA-core kernels (parity tests).
"""
import torch
import torch.nn as nn

from . import ops
from .unet3D import NoBottleneck, conv3x3x3, in_place

_PAD = ((32, 32), (64, 64), (128, 128), (256, 256), (512, 512))


def _pad_width(c):
    for limit, p in _PAD:
        if c <= limit:
            return p
    raise NotImplementedError(f"{c} channels")


class _Layout:
    """Where the real channels of a C-channel tensor sit inside its zero-padded tensor-core width.

    ``groups`` = GroupNorm group count of the consumers (0: none).  If the padded width splits into ``groups`` equal parts
    every group is padded in place (GroupNorm over ``groups`` groups of the padded tensor with ``real_cpg`` real channels
    each); otherwise the real channels come first and the padded tensor is normalised in groups of C/groups channels
    (the trailing all-zero groups normalise to zero)."""

    def __init__(self, c, groups=0, width=0):
        self.c, self.p = c, max(_pad_width(c), width)
        if groups and c % groups == 0 and self.p % groups == 0:
            cpg, cpgp = c // groups, self.p // groups
            self.index = torch.tensor([(i // cpg) * cpgp + i % cpg for i in range(c)], dtype=torch.long)
            self.gn_groups, self.real_cpg = groups, (cpg if cpg != cpgp else 0)
        else:
            self.index = torch.arange(c, dtype=torch.long)
            cpg = c // groups if groups else 1
            self.gn_groups, self.real_cpg = (self.p // cpg if groups else 0), 0
        self._dev = {}

    def idx(self, device):
        if device not in self._dev:
            self._dev[device] = self.index.to(device)
        return self._dev[device]

    def same(self, other):
        return self.c == other.c and torch.equal(self.index, other.index)


def _standardised(w):
    """Conv3d.forward of the reference (unet3D.py:22-26) as tensor ops (differentiable)."""
    m = w.mean(dim=(1, 2, 3, 4), keepdim=True)
    c = w - m
    std = torch.sqrt(torch.var(c.reshape(c.shape[0], -1), dim=1) + 1e-12).view(-1, 1, 1, 1, 1)
    return c / std


def _padded_weight(conv, lay_in, lay_out, standardise):
    w = conv.weight.float()
    if standardise:
        w = _standardised(w)
    dev = w.device
    wp = w.new_zeros((lay_out.p, lay_in.p) + tuple(w.shape[2:]))
    return wp.index_put((lay_out.idx(dev)[:, None], lay_in.idx(dev)[None, :]), w)


def _padded_affine(gn, lay):
    dev = gn.weight.device
    g = gn.weight.new_zeros(lay.p).index_put((lay.idx(dev),), gn.weight)
    b = gn.bias.new_zeros(lay.p).index_put((lay.idx(dev),), gn.bias)
    return g, b


def _conv(conv, x, lay_in, lay_out, residual=None):
    standardise = getattr(conv, "_standardise", False)
    return ops.ws_conv3d(x, _padded_weight(conv, lay_in, lay_out, standardise), conv._s, False, residual)


def _relayout(x, lay_from, lay_to):
    """Move the real channels of a padded channels-last tensor from one layout to another (one gather along channels)."""
    if lay_from.same(lay_to):
        return x
    dev = x.device
    src = torch.full((lay_to.p,), -1, dtype=torch.long, device=dev)
    src[lay_to.idx(dev)] = lay_from.idx(dev)
    rows = x.permute(0, 2, 3, 4, 1)
    out = rows.index_select(-1, src.clamp_min(0)) * (src >= 0).to(rows.dtype)
    return out.permute(0, 4, 1, 2, 3)


def _block(blk, x, lay_in, lay_out):
    """NoBottleneck.forward (unet3D.py:56-73) on padded tensors."""
    g1, b1 = _padded_affine(blk.gn1, lay_in)
    eps = blk.gn1.eps
    if blk.downsample is not None:
        g2, b2 = _padded_affine(blk.downsample[0], lay_in)
        a1, ads = ops.gn_relu_dual(x, g1, b1, g2, b2, lay_in.gn_groups, eps, real_cpg=lay_in.real_cpg)
        residual = _conv(blk.downsample[2], ads, lay_in, lay_out)
    else:
        a1, residual = ops.gn_relu(x, g1, b1, lay_in.gn_groups, eps, alias=True, real_cpg=lay_in.real_cpg)
    out = _conv(blk.conv1, a1, lay_in, lay_out)
    g, b = _padded_affine(blk.gn2, lay_out)
    a2 = ops.gn_relu(out, g, b, lay_out.gn_groups, blk.gn2.eps, real_cpg=lay_out.real_cpg)
    return _conv(blk.conv2, a2, lay_out, lay_out, residual)


class unet3D_g(nn.Module):
    """Reference unet3D.py:1507-1623: the light-weight refiner.  ``forward(input, _=None) -> logits`` at the input
    resolution (stride-2 stem, the backbone of unet3D_baseline at half resolution with GroupNorm(4) blocks, final
    trilinear x2 up-sampling of the logits)."""

    def __init__(self, layers, num_classes=3, weight_std=False, in_channel=2, init_filter=32):
        self.inplanes = 128
        self.weight_std = weight_std
        self.init_filter = init_filter
        super(unet3D_g, self).__init__()
        f = init_filter
        self.conv0 = conv3x3x3(in_channel, f, stride=[2, 2, 2], weight_std=self.weight_std)
        self.conv1 = conv3x3x3(f, f, stride=[1, 1, 1], weight_std=self.weight_std)
        self.layer0 = self._make_layer(NoBottleneck, f, f, layers[0], stride=(1, 1, 1))
        self.layer1 = self._make_layer(NoBottleneck, f, f * 2, layers[1], stride=(2, 2, 2))
        self.layer2 = self._make_layer(NoBottleneck, f * 2, f * 4, layers[2], stride=(2, 2, 2))
        self.layer3 = self._make_layer(NoBottleneck, f * 4, f * 8, layers[3], stride=(2, 2, 2))
        self.layer4 = self._make_layer(NoBottleneck, f * 8, f * 8, layers[4], stride=(2, 2, 2))
        self.fusionConv = nn.Sequential(
            nn.GroupNorm(f // 2, f * 8),
            nn.ReLU(inplace=in_place),
            conv3x3x3(f * 8, f * 8, kernel_size=(1, 1, 1), padding=(0, 0, 0), weight_std=self.weight_std))
        self.upsamplex2 = nn.Upsample(scale_factor=2, mode='trilinear')
        self.x8_resb = self._make_layer(NoBottleneck, f * 8, f * 4, 1, stride=(1, 1, 1))
        self.x4_resb = self._make_layer(NoBottleneck, f * 4, f * 2, 1, stride=(1, 1, 1))
        self.x2_resb = self._make_layer(NoBottleneck, f * 2, f, 1, stride=(1, 1, 1))
        self.x1_resb = self._make_layer(NoBottleneck, f, f, 1, stride=(1, 1, 1))
        self.precls_conv = nn.Sequential(
            nn.GroupNorm(f // 4, f),
            nn.ReLU(inplace=in_place),
            nn.Conv3d(f, num_classes, kernel_size=1))
        self.in_channel = in_channel

    def _make_layer(self, block, inplanes, planes, blocks, stride=(1, 1, 1), dilation=1, multi_grid=1):
        downsample = None
        if stride[0] != 1 or stride[1] != 1 or stride[2] != 1 or inplanes != planes:
            downsample = nn.Sequential(
                nn.GroupNorm(4, inplanes),
                nn.ReLU(inplace=in_place),
                conv3x3x3(inplanes, planes, kernel_size=(1, 1, 1), stride=stride, padding=0, weight_std=self.weight_std))
        layers = [block(inplanes, planes, stride, dilation=dilation, downsample=downsample, multi_grid=1,
                        weight_std=self.weight_std, group=4)]
        for _ in range(1, blocks):
            layers.append(block(planes, planes, dilation=dilation, multi_grid=1, weight_std=self.weight_std, group=4))
        return nn.Sequential(*layers)

    def _stage(self, layer, x, lay_in, planes):
        lay = lay_in
        for blk in layer:
            lay_out = _Layout(planes, 4)
            x = _block(blk, x, lay, lay_out)
            lay = lay_out
        return x, lay

    def forward(self, input, _=None):
        dt = ops.get_compute_dtype()                 # bf16: tcgen05 kernels; fp32: the exact CUDA-core kernels (parity tests)
        f = self.init_filter
        ops.begin_forward(input.device)
        lay_img = _Layout(self.in_channel)
        x = input.to(dt)
        pad = x.new_zeros((x.shape[0], lay_img.p - self.in_channel) + tuple(x.shape[2:]))
        x = ops.to_cl(torch.cat([x, pad], 1), dt)
        # conv0 is a stride-2 3x3x3 convolution: its tcgen05 weight-gradient kernel wants a multiple of 64 output channels,
        # so its output (consumed by conv1 directly, no GroupNorm in between, :1574-1576) is carried 64 wide
        lay0 = _Layout(f, 0, width=64)
        lay = _Layout(f, 4)
        x = _conv(self.conv0, x, lay_img, lay0)
        x = _conv(self.conv1, x, lay0, lay)
        x, lay = self._stage(self.layer0, x, lay, f)
        skips = [(x, lay)]
        for layer, planes in ((self.layer1, 2 * f), (self.layer2, 4 * f), (self.layer3, 8 * f)):
            x, lay = self._stage(layer, x, lay, planes)
            skips.append((x, lay))
        x, lay = self._stage(self.layer4, x, lay, 8 * f)
        # fusionConv: GroupNorm(f // 2 groups) -> ReLU -> 1x1x1 conv (:1525-1529); its groups are not the blocks' groups
        gn, conv = self.fusionConv[0], self.fusionConv[2]
        lay_f = _Layout(8 * f, gn.num_groups)
        x = _relayout(x, lay, lay_f)
        g, b = _padded_affine(gn, lay_f)
        a = ops.gn_relu(x, g, b, lay_f.gn_groups, gn.eps, real_cpg=lay_f.real_cpg)
        lay = skips[3][1]
        x = _conv(conv, a, lay_f, lay)
        for resb, (skip, lay_s), planes in ((self.x8_resb, skips[3], 4 * f), (self.x4_resb, skips[2], 2 * f),
                                            (self.x2_resb, skips[1], f), (self.x1_resb, skips[0], f)):
            x = ops.upsample2x_add(x, skip)                      # nn.Upsample(trilinear x2) + skip (:1595-1617)
            x, lay = self._stage(resb, x, lay_s, planes)
        gn, conv = self.precls_conv[0], self.precls_conv[2]
        lay_c = _Layout(f, gn.num_groups)
        x = _relayout(x, lay, lay_c)
        g, b = _padded_affine(gn, lay_c)
        a = ops.gn_relu(x, g, b, lay_c.gn_groups, gn.eps, real_cpg=lay_c.real_cpg)
        dev = a.device
        wc = conv.weight.new_zeros((conv.out_channels, lay_c.p, 1, 1, 1)).index_put(
            (torch.arange(conv.out_channels, device=dev)[:, None], lay_c.idx(dev)[None, :]), conv.weight)
        logits = ops.classifier(a, wc, conv.bias)
        return ops.upsample2x_ncdhw(logits)                      # self.upsamplex2(logits), :1621


class Reshape(nn.Module):
    def forward(self, x):
        return x.view(x.shape[0], -1)


_S2D_TAP = ((0, 1), (1, 0), (1, 1), (2, 0))     # tap t of a 4-wide stride-2 window -> (3-tap index, parity)


_S2D_COLS = {}


def _s2d_cols(c, dev):
    key = (c, str(dev))
    if key not in _S2D_COLS:
        t = torch.arange(4)
        k3 = torch.tensor([k for k, _ in _S2D_TAP])[t]          # 3-tap index of tap t
        par = torch.tensor([p for _, p in _S2D_TAP])[t]         # parity of tap t
        kidx = (k3[:, None, None] * 9 + k3[None, :, None] * 3 + k3[None, None, :]).reshape(-1)          # [64]
        pidx = (par[:, None, None] * 4 + par[None, :, None] * 2 + par[None, None, :]).reshape(-1)        # [64]
        ci = torch.arange(c)[:, None]
        _S2D_COLS[key] = ((pidx[None, :] * c + ci) * 27 + kidx[None, :]).reshape(-1).to(dev)            # [c * 64]
    return _S2D_COLS[key]


def _s2d_weight(w, cp):
    """[Cout, C, 4, 4, 4] -> [Cout, cp, 3, 3, 3] for the space-to-depth input (channel = parity * C + c)."""
    cout, c = w.shape[0], w.shape[1]
    dev = w.device
    cols = _s2d_cols(c, dev)
    flat = w.new_zeros((cout, cp * 27)).index_put((torch.arange(cout, device=dev)[:, None], cols[None, :]),
                                                  w.reshape(cout, c * 64).float())
    return flat.view(cout, cp, 3, 3, 3)


def _conv4s2(conv, x):
    """nn.Conv3d(C, Cout, kernel_size=4, stride=2, padding=1) + LeakyReLU(0.2) on the tensor-core kernels."""
    c = conv.in_channels
    cp = 32 if 8 * c <= 32 else (8 * c + 63) // 64 * 64
    xs = ops.space_to_depth2(x, cp)
    y = ops.ws_conv3d(xs, _s2d_weight(conv.weight, cp), 1, False)
    return ops.bias_leaky_relu(y, conv.bias, 0.2)


class norm_style_discriminator_output(nn.Module):
    """Reference unet3D.py:1907-1947: six Conv3d(k=4, s=2, p=1) + LeakyReLU(0.2) stages, global average pooling and a
    linear head.  ``forward(x_in) -> [B, 2]``.  Input extents must be even at every stage (64x192x192 patches are)."""

    def __init__(self, num_classes, ndf=32):
        super(norm_style_discriminator_output, self).__init__()
        self.ndf = ndf

        def stage(cin, cout):
            return [nn.Conv3d(cin, cout, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(negative_slope=0.2, inplace=True)]

        self.block1 = nn.Sequential(*stage(num_classes, ndf))
        self.block2 = nn.Sequential(*stage(ndf, ndf * 2))
        self.block3 = nn.Sequential(*stage(ndf * 2, ndf * 4))
        self.block4 = nn.Sequential(*(stage(ndf * 4, ndf * 8) + stage(ndf * 8, ndf * 8) + stage(ndf * 8, ndf * 8) +
                                      [nn.AdaptiveAvgPool3d(1), Reshape(), nn.Linear(ndf * 8, 2)]))

    def forward(self, x_in):
        dt = ops.get_compute_dtype()
        x = ops.to_cl(x_in.to(dt), dt)
        convs = [self.block1[0], self.block2[0], self.block3[0], self.block4[0], self.block4[2], self.block4[4]]
        for conv in convs:
            x = _conv4s2(conv, x)
        pooled = x.float().mean(dim=(2, 3, 4))                   # AdaptiveAvgPool3d(1) + Reshape on <= 27 voxels
        lin = self.block4[8]
        return pooled @ lin.weight.t().float() + lin.bias.float()
