"""Fixture for the pseudo-label (refiner) branch of get_loss (reference loss_functions/losses.py:131-178), written by
running the UNMODIFIED reference function on seeded inputs.  Authoring container only:
    python oracle/make_golden_get_loss.py      ->  tests/golden/get_loss_refine.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refload import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def inputs(seed=0, organs=3, dhw=(6, 10, 12)):
    g = torch.Generator().manual_seed(seed)
    C = organs + 1
    output = 2 * torch.randn((1, C) + dhw, generator=g)
    target = torch.randint(0, C, (1, 1) + dhw, generator=g).float()
    attns = [torch.randn((1, organs) + dhw, generator=g) for _ in range(3)]
    refine = 3 * torch.randn((organs, 2) + dhw, generator=g)
    deep = [torch.randn((1, C) + tuple(max(1, s // 2) for s in dhw), generator=g)]
    return output, target, attns, refine, deep


def main():
    load_reference()
    from loss_functions import loss_partial as ref_lp
    from loss_functions import losses as ref_losses

    ref_lp.autocast = lambda enabled=False: torch.autocast("cpu", enabled=enabled)
    fix = {}
    for tag, label_t, wmask in [("mixed", [True, False, False], [1.0, 1.0, 0.0, 0.0]),
                                ("none_supervised", [False, False, False], [1.0, 0.0, 0.0, 0.0])]:
        output, target, attns, refine, deep = inputs()
        leaves = [output.requires_grad_(True)] + [a.requires_grad_(True) for a in attns] + [refine.requires_grad_(True)]
        loss, confi = ref_losses.get_loss(output, 0, deep, target, mask=[torch.tensor(wmask)], attns=list(attns),
                                          refine_output=refine, label_t=label_t, aux_weight=0.7, weight_feature=0.3)
        loss.backward()
        fix[tag + ":loss"] = np.float64(loss.item())
        fix[tag + ":confi"] = np.float64(confi)
        fix[tag + ":label_t"] = np.array(label_t)
        fix[tag + ":wmask"] = np.array(wmask, dtype=np.float32)
        for name, t in zip(["output", "attn0", "attn1", "attn2", "refine"], leaves):
            fix[tag + ":grad:" + name] = t.grad.numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "get_loss_refine.npz"), **fix)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in fix.items()})


if __name__ == "__main__":
    main()
