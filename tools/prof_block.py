"""One GN+ReLU -> 3x3x3 conv (C->C, cfg2 resolution of the given level) forward+backward, for ncu captures of the
fprop / dgrad / wgrad kernels in isolation.  Usage: python tools/prof_block.py [C=32] [level=0] [reps=1]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_pl_b200 as mm
from multimodal_pl_b200 import ops, _lib

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lvl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
_lib.require_device()
dev = torch.device("cuda")
N, D, H, W = 2, 64 >> lvl, 192 >> lvl, 192 >> lvl
x = torch.randn((N, D, H, W, C), device=dev).bfloat16().permute(0, 4, 1, 2, 3).requires_grad_(True)
g, b = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
w = torch.randn(C, C, 3, 3, 3, device=dev, requires_grad=True)
dy = torch.randn((N, D, H, W, C), device=dev).bfloat16().permute(0, 4, 1, 2, 3)
for _ in range(reps):
    ops.begin_forward(dev)
    y = ops.ws_conv3d(ops.gn_relu(x, g, b), w, 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    y.backward(dy)
    ev[1].record()
    torch.cuda.synchronize()
print("backward ms", ev[0].elapsed_time(ev[1]))
