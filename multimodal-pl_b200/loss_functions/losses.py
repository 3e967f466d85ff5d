"""Drop-in for ``get_loss`` of the reference's ``loss_functions/losses.py`` (:107-182) -- the call the train loop
makes (train_amos_atlas_final.py:303-312).  The base partial-label term and the deep-supervision terms run on the
fused kernel; the refiner pseudo-label branch (losses.py:131-178) is outside the hot path."""
import torch
from torch import nn

from .loss_partial import EDiceLoss_partial


def get_loss(output, cm, deep_out, target, mask=None, catlas=None, attns=None, refine_output=None, label_t=None,
             discard=0.05, confi_=0.10, aux_weight=1, weight_feature=0.1):
    edice = EDiceLoss_partial(output.shape[1])
    dice_loss = edice(output, target.squeeze(1), soft_max=True, mask=mask)            # losses.py:113
    aux_loss = 0.0
    weights = [0.125, 0.25, 0.5, 1]
    if len(deep_out) != 0:                                                             # losses.py:119-129
        for idx, l in enumerate(deep_out):
            ctarget = nn.functional.interpolate(target, l.shape[2:], mode='nearest').float()
            aux_loss = aux_loss + edice(l, ctarget.squeeze(1), soft_max=True, mask=mask, uce=False) * weights[idx]
    if refine_output is not None:
        raise NotImplementedError("get_loss: the refiner pseudo-label branch (reference losses.py:131-178) is not "
                                  "part of the B200 hot path")
    if torch.is_tensor(aux_loss):
        # the reference returns dice_loss alone here (losses.py:179-182: aux terms are only added with a refiner)
        pass
    return dice_loss, confi_
