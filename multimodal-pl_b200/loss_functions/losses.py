"""Drop-in for ``get_loss`` of the reference's ``loss_functions/losses.py`` (:107-182) -- the call the train loop
makes (train_amos_atlas_final.py:303-312).  The base partial-label term and the deep-supervision terms run on the fused
partial-label kernel (ops.partial_label_loss); the pseudo-label terms of the refiner branch (:131-178) run on the fused
gated-Dice kernel (ops.masked_dice) -- one launch per (scale, unsupervised organ) instead of ~10 ATen kernels, a
boolean-index gather and a ``.cpu().numpy()`` synchronisation each (:171)."""
import torch
from torch import nn

from .loss_partial import EDiceLoss_full2, EDiceLoss_partial


def get_loss(output, cm, deep_out, target, mask=None, catlas=None, attns=None, refine_output=None, label_t=None,
             discard=0.05, confi_=0.10, aux_weight=1, weight_feature=0.1):
    edice = EDiceLoss_partial(output.shape[1])
    num_classes = output.shape[1] - 1
    dice_loss = edice(output, target.squeeze(1), soft_max=True, mask=mask)            # losses.py:113
    aux_loss = 0.0
    weights = [0.125, 0.25, 0.5, 1]
    if len(deep_out) != 0:                                                             # losses.py:119-129
        for idx, l in enumerate(deep_out):
            ctarget = nn.functional.interpolate(target, l.shape[2:], mode='nearest').float()
            aux_loss = aux_loss + edice(l, ctarget.squeeze(1), soft_max=True, mask=mask, uce=False) * weights[idx]
    if refine_output is None:
        # the reference returns dice_loss alone here (losses.py:179-182: aux terms are only added with a refiner)
        return dice_loss, confi_

    # ---- pseudo labels from the refiner (losses.py:131-178)
    refine_output_p = torch.softmax(refine_output, 1)                                  # [organs, 2, D, H, W]
    confi_mask = torch.logical_or(refine_output_p > (1 - confi_), refine_output_p < confi_).float()   # :141
    confi_ = 0.10                                                                      # :143
    supcount = sum(1 for l in range(refine_output_p.shape[0]) if label_t[l])           # :149-153
    # (the reference also assembles ``refine_label`` here, :134,:149-159; nothing reads it afterwards)
    cedice = EDiceLoss_full2(2)
    attns = list(attns) + [torch.softmax(output, 1)[:, 1:]]                            # :161 (popped again at :174)
    for idx, l in enumerate(attns):
        for gan in range(num_classes):
            if label_t[gan]:
                continue
            # the last entry is already a probability map (:167-168), the attention maps are logits (:169-170)
            cdice = cedice(l[:, gan:gan + 1], refine_output_p[gan:gan + 1, 1], uce=False, sigmoid=(idx != 3),
                           mask=confi_mask[gan:gan + 1, 1:])
            aux_loss = aux_loss + cdice / (num_classes - supcount) * weights[idx] * weight_feature
    return dice_loss + aux_loss * aux_weight, confi_
