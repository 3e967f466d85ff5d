"""Fixture for unet3D_with_feam3 (reference unet3D.py:938-1190), written by running the UNMODIFIED reference model on
seeded weights / inputs on the CPU.  Authoring container only:
    python oracle/make_golden_feam3.py      ->  tests/golden/feam3.npz
Stored: all four train-mode outputs, the eval-mode logits flag, gradients of a few tensors under a fixed scalar
objective, and the class tokens after one renew_token() call."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mmpl_oracle as O  # noqa: E402
from _refload import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SHAPE, SEED, CLASSES = (1, 1, 16, 32, 32), 3, 16
GRAD_KEYS = ["conv1.weight", "x2_resb.0.conv1.weight", "deepout1.2.weight", "deepout3.2.bias", "eam84.kv.weight",
             "eam42.q.weight", "eam21.norm2.weight", "eam21.norm3.bias", "precls_conv.2.weight"]


def objective(logits, attn, deep, dys):
    """fixed scalar: sum_i <out_i, dy_i> with seeded dy (shared by the generator and the test)"""
    outs = [logits] + list(attn) + list(deep)
    return sum((o.float() * d.to(o.device)).sum() for o, d in zip(outs, dys))


def seeded_dys(shapes):
    return [torch.randn(s, generator=torch.Generator().manual_seed(900 + i)) / float(np.prod(s)) ** 0.5
            for i, s in enumerate(shapes)]


def main():
    ref_unet, _, _ = load_reference()
    torch.set_num_threads(os.cpu_count())
    sd, tokens = O.synth_feam3_state_dict(CLASSES, SEED)
    model = ref_unet.unet3D_with_feam3([1, 2, 2, 2, 2], num_classes=CLASSES, weight_std=True)
    model.load_state_dict(sd)
    model.class_token1, model.class_token2, model.class_token3 = [t.clone() for t in tokens]
    model.train()
    x = O.synth_patch(SHAPE, 1000 + SEED, "ct")
    lab = O.synth_labels((SHAPE[0],) + SHAPE[2:], 2000 + SEED, CLASSES, 32)
    logits, attn, deep, feats = model(x, lab)
    fix = {"logits": logits.detach().numpy()}
    for i in range(3):
        fix[f"attn{i}"] = attn[i].detach().numpy()
        fix[f"deep{i}"] = deep[i].detach().numpy()
        fix[f"feat{i}"] = feats[i].detach().numpy()
    dys = seeded_dys([tuple(logits.shape)] + [tuple(a.shape) for a in attn] + [tuple(d.shape) for d in deep])
    objective(logits, attn, deep, dys).backward()
    params = dict(model.named_parameters())
    for k in GRAD_KEYS:
        fix["grad:" + k] = params[k].grad.numpy()
    fix["unused_grad_is_none"] = np.array([params["eam84.proj.weight"].grad is None])
    model.renew_token(feats, lab)
    for i, t in enumerate([model.class_token1, model.class_token2, model.class_token3]):
        fix[f"token{i}"] = t.detach().numpy()
    model.eval()
    with torch.no_grad():
        fix["eval_equals_train_logits"] = np.array([torch.equal(model(x), logits.detach())])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "feam3.npz"), **fix)
    print({k: v.shape for k, v in fix.items()})


if __name__ == "__main__":
    main()
