"""Stem conv (1 -> 32, fused tcgen05 kernels) forward + weight gradient at cfg2 size, for ncu captures.
Usage: python tools/prof_stem.py [reps=3]"""
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
img = torch.randn(N, 1, D, H, W, device=dev)
w = torch.randn(32, 1, 3, 3, 3, device=dev, requires_grad=True)
dy = torch.randn((N, D, H, W, 32), device=dev).bfloat16().permute(0, 4, 1, 2, 3)
for _ in range(reps):
    ops.begin_forward(dev)
    y = ops.stem_conv(img, w)
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    y2 = ops.stem_conv(img, w)
    b.record()
    y2.backward(dy)
    c.record()
    torch.cuda.synchronize()
    print("stem fwd ms", a.elapsed_time(b), "bwd ms", b.elapsed_time(c))
