"""Logits error of the bf16 path per stem mode (golden b1/b2)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O
import multimodal_pl_b200 as mm
from multimodal_pl_b200.unet3D import unet3D_baseline

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

for tag in ["b1", "b2"]:
    g = np.load(os.path.join(ROOT, "tests/golden", f"unet_{tag}.npz"))
    shape, seed = tuple(int(v) for v in g["shape"]), int(g["seed"])
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
    model.load_state_dict(O.synth_state_dict(32, 16, seed)); model.eval()
    x = O.synth_patch(shape, 1000 + seed, "ct" if seed == 0 else "mri").cuda()
    ref = torch.from_numpy(g["logits"])
    for mode in ["direct", "fp32fwd", "tc32", "split"]:
        mm.set_stem_mode(mode)
        with torch.no_grad():
            lg = model(x)
            st = model.conv1(x)
        print(tag, mode, f"logits rel {rel(lg, ref):.4e}", flush=True)
