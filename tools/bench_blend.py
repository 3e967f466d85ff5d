"""Micro-benchmark of mmpl_cls_blend (classifier + Gaussian accumulation of one 64x192x192 tile into a 300x512x512
depth-major accumulator): CUDA-event time per launch over 24 launches at the first 24 tile origins of the volume.
MMPL_LIB selects the library build.  Usage: python tools/bench_blend.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_pl_b200 import _lib, ops  # noqa: E402
from multimodal_pl_b200.evaluate import _gaussian_device, tile_origins  # noqa: E402

dev = torch.device("cuda", 0)
tile, vol, C = (64, 192, 192), (300, 512, 512), 16
torch.manual_seed(0)
a = torch.randn((1, 32) + tile, device=dev).to(torch.bfloat16)
a = ops.to_cl(a, torch.bfloat16)
w = torch.randn(C, 32, 1, 1, 1, device=dev) * 0.1
b = torch.randn(C, device=dev) * 0.1
acc = torch.zeros((1, vol[0], C) + vol[1:], dtype=torch.float32, device=dev)
origins = torch.tensor(tile_origins((1, 1) + vol, tile)[:24], dtype=torch.int32, device=dev)
g = _gaussian_device(tile, dev)
sinks = [ops.BlendSink(acc, g, origins[i].contiguous(), tile, d_outer=True) for i in range(24)]
for s in sinks[:4]:
    ops.classifier_blend(a, w, b, s)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(25)]
e[0].record()
for i, s in enumerate(sinks):
    ops.classifier_blend(a, w, b, s)
    e[i + 1].record()
torch.cuda.synchronize()
t = sorted(e[i].elapsed_time(e[i + 1]) for i in range(24))
byt = 2.36e6 * (32 * 2 + 2 * 16 * 4)
print(f"{os.environ.get('MMPL_LIB', 'default lib')}: cls_blend median {t[12] * 1e3:.1f} us, min {t[0] * 1e3:.1f} us per tile "
      f"({byt / t[12] / 1e6:.0f} GB/s of {byt / 1e6:.0f} MB algorithmic), checksum {acc.double().sum().item():.6e}")
