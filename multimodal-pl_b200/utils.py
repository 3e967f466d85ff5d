"""Host-side helpers of the reference's ``utils.py`` that the train loop uses around the hot path (kept verbatim in
behaviour): poly learning-rate schedule (utils.py:53-60) and the scalar all-reduce stub (:72-74)."""
import torch


def lr_poly(base_lr, iter, max_iter, power):
    return base_lr * ((1 - float(iter) / max_iter) ** (power))


def adjust_learning_rate(optimizer, i_iter, lr, num_stemps, power):
    """Sets param_groups[0]['lr'] to the poly-decayed value (FusedSGD mirrors it into its device scalar)."""
    lr = lr_poly(lr, i_iter, num_stemps, power)
    optimizer.param_groups[0]['lr'] = lr
    return lr


def all_reduce_tensor(tensor, norm=True):
    return torch.mean(tensor)
