"""Partial-label loss forward + backward and the classifier forward + backward at cfg2 size, for ncu captures.
Usage: python tools/prof_loss.py [reps=3]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_pl_b200 as mm
from multimodal_pl_b200 import ops, _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
_lib.require_device()
mm.set_compute_dtype(torch.bfloat16)
dev = torch.device("cuda")
N, D, H, W = 2, 64, 192, 192
a = torch.randn((N, D, H, W, 32), device=dev).bfloat16().permute(0, 4, 1, 2, 3).requires_grad_(True)
wc = torch.randn(16, 32, 1, 1, 1, device=dev, requires_grad=True)
bc = torch.zeros(16, device=dev, requires_grad=True)
tgt = torch.randint(0, 16, (N, D, H, W), device=dev).to(torch.uint8)
cw = torch.tensor([[1.0, 0, 0, 0, 1.0] + [0.0] * 11, [1.0] + [0.0] * 15], device=dev)
for _ in range(reps):
    logits = ops.classifier(a, wc, bc)
    loss = ops.partial_label_loss(logits, tgt, cw, per_sample=True)
    loss.backward()
    torch.cuda.synchronize()
print("ok", float(loss))
