"""Drop-in for the reference's ``loss_functions/loss_partial.py`` (DiceLoss :10-57, EDiceLoss_partial :59-99,
EDiceLoss_full2 :137-170) with the same class names and ``forward`` signatures.

The training hot path -- ``EDiceLoss_partial(C)(logits, target, mask=[w...], soft_max=True)`` as called from
``losses.get_loss`` (losses.py:113) -- runs as two fused CUDA kernels (one forward pass, one backward pass) with no
one-hot tensor and no host synchronisation.  The remaining call forms (sigmoid inputs, explicit voxel gates,
``DiceLoss`` on probabilities) are evaluated with device-side tensor ops, also without ``.item()`` syncs.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


_weight_cache = {}


def _class_weight(mask, n_classes, device, per_sample=False):
    """``mask[0]`` of the reference (loss_partial.py:87,92): the first sample's class-weight vector, as a device
    tensor -- or, with ``per_sample``, all of them stacked [B, C].  Host vectors are uploaded once per distinct value
    (a pageable H2D copy per step would serialise the host with the GPU stream)."""
    if mask is None:
        assert not per_sample, "per-sample class weights need one weight vector per sample"
        key = ("ones", n_classes, str(device))
        if key not in _weight_cache:
            _weight_cache[key] = torch.ones(n_classes, dtype=torch.float32, device=device)
        return _weight_cache[key]
    rows = list(mask) if per_sample else [mask[0]]
    if len(rows) == 1 and torch.is_tensor(rows[0]) and rows[0].device == torch.device(device) and rows[0].dtype == torch.float32:
        return rows[0]
    vals = tuple(tuple(float(v) for v in (w.tolist() if torch.is_tensor(w) else w)) for w in rows)
    key = (vals, str(device))
    if key not in _weight_cache:
        if len(_weight_cache) > 4096:
            _weight_cache.clear()
        t = torch.tensor(vals, dtype=torch.float32, device=device)
        _weight_cache[key] = t if per_sample else t[0]
    return _weight_cache[key]


class DiceLoss(nn.Module):
    def __init__(self, n_classes):
        super(DiceLoss, self).__init__()
        self.n_classes = n_classes

    def _one_hot_encoder(self, input_tensor):
        idx = torch.arange(self.n_classes, device=input_tensor.device, dtype=input_tensor.dtype)
        shape = (1, self.n_classes) + (1,) * (input_tensor.dim() - 1)
        return (input_tensor.unsqueeze(1) == idx.view(shape)).float()

    def _dice_loss(self, score, target, mask):
        """Binary Dice over the voxels where ``mask`` is true (loss_partial.py:24-36): one fused kernel on the device
        (ops.masked_dice); plain torch ops only for inputs whose shapes do not line up element by element."""
        if score.is_cuda and score.numel() == target.numel() == mask.numel():
            return ops.masked_dice(score, target, mask, sigmoid=False, uce=False)
        target = target.float()
        m = mask.bool()
        score = score[m]
        target = target[mask.squeeze(1).bool()] if mask.dim() == target.dim() + 1 else target[m]
        smooth = 1e-5
        intersect = torch.sum(score * target)
        y_sum = torch.sum(target * target)
        z_sum = torch.sum(score * score)
        return 1 - (2 * intersect + smooth) / (z_sum + y_sum + smooth)

    def forward(self, inputs, target, weight=None, softmax=True, mask=None):
        """``inputs`` are probabilities [B,C,...] (the ``softmax`` flag is ignored by the reference too, :38-57)."""
        onehot = self._one_hot_encoder(target)
        assert inputs.size() == onehot.size(), 'predict {} & target {} shape do not match'.format(inputs.size(), onehot.size())
        if weight is None:
            weight = torch.ones(self.n_classes, device=inputs.device)
        elif not torch.is_tensor(weight):
            weight = torch.tensor(list(weight), dtype=torch.float32)
        weight = weight.to(inputs.device, torch.float32)
        dims = [0] + list(range(2, inputs.dim()))
        gate = 1.0 if mask is None else mask.to(inputs.dtype)
        smooth = 1e-5
        inter = torch.sum(inputs * onehot * gate, dim=dims)
        y_sum = torch.sum(onehot * onehot * gate, dim=dims)
        z_sum = torch.sum(inputs * inputs * gate, dim=dims)
        dice = 1 - (2 * inter + smooth) / (z_sum + y_sum + smooth)
        return torch.sum(dice * weight) / self.n_classes


class EDiceLoss_partial(nn.Module):
    """Dice + class-gated BCE on softmax probabilities (reference loss_partial.py:59-99)."""

    def __init__(self, n_classes):
        super(EDiceLoss_partial, self).__init__()
        self.device = "cuda"
        self.n_classes = n_classes
        self.diceloss = DiceLoss(n_classes=n_classes)
        self.bce = nn.BCELoss()

    def forward(self, inputs, target, mask=None, soft_max=True, uce=True, lut=None, per_sample=False):
        """inputs [B,C,D,H,W] logits, target [B,D,H,W] class ids (float like the reference, or uint8), mask = list of
        per-sample weight vectors.  As in the reference only mask[0] is used and the Dice sums pool over the batch
        (loss_partial.py:87,92) -- unless ``per_sample=True`` (not part of the reference signature): then
        sample b is scored with mask[b] exactly as the reference would score it alone, and the batch is averaged, which
        is what a mixed CT/MRI partial-label batch needs.  ``lut`` ([C], or [B,C] with per_sample) optionally folds the
        cmask remap of train_amos_atlas_final.py:252-255 into the kernel."""
        w = _class_weight(mask, inputs.shape[1], inputs.device, per_sample and soft_max)
        if soft_max:
            return ops.partial_label_loss(inputs, target, w, lut=lut, uce=uce, per_sample=per_sample)
        assert not per_sample, "per_sample is implemented for the soft_max=True training form"
        # sigmoid variant (not used by the train loop): same formula on independent sigmoids
        p = torch.sigmoid(inputs)
        dice = self.diceloss(p, target, softmax=False, weight=w)
        if not uce:
            return dice
        onehot = self.diceloss._one_hot_encoder(target)
        dims = [0] + list(range(2, inputs.dim()))
        ce = F.binary_cross_entropy(p.float(), onehot, reduction='none').mean(dim=dims)
        return dice + torch.sum(ce * w)


class EDiceLoss_full2(nn.Module):
    """Binary Dice gated by a voxel confidence mask (+ BCE-with-logits), reference loss_partial.py:137-170."""

    def __init__(self, n_classes):
        super(EDiceLoss_full2, self).__init__()
        self.device = "cuda"
        self.n_classes = n_classes
        self.diceloss = DiceLoss(n_classes=n_classes)
        self.bce = nn.BCEWithLogitsLoss()

    def forward(self, inputs, target, uce=True, mask=None, sigmoid=True):
        if inputs.is_cuda and inputs.numel() == target.numel() and (mask is None or mask.numel() == target.numel()):
            # fused: sigmoid, gated Dice sums and BCE-with-logits (always on the raw inputs, :168) in one pass
            return ops.masked_dice(inputs, target, mask, sigmoid=sigmoid, uce=uce)
        p = torch.sigmoid(inputs) if sigmoid else inputs
        if mask is None:
            mask = torch.ones_like(target).unsqueeze(0)
        dice = self.diceloss._dice_loss(p, target, mask)
        if uce:
            return dice + self.bce(inputs.float().squeeze(0), target.float())
        return dice
