"""1x1x1 convolution Cin -> Cout on a cfg2 level-`lvl` tensor, forward + backward, timed back to back (CUDA events) and
for ncu captures.  Usage: python tools/prof_1x1.py [Cin=32] [Cout=64] [level=1] [reps=20]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_pl_b200 as mm
from multimodal_pl_b200 import ops, _lib

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lvl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
_lib.require_device()
mm.set_compute_dtype(torch.bfloat16)
dev = torch.device("cuda")
N, D, H, W = 2, 64 >> lvl, 192 >> lvl, 192 >> lvl
x = torch.randn((N, D, H, W, cin), device=dev).bfloat16().permute(0, 4, 1, 2, 3).requires_grad_(True)
w = torch.randn(cout, cin, 1, 1, 1, device=dev, requires_grad=True)
dy = torch.randn((N, D, H, W, cout), device=dev).bfloat16().permute(0, 4, 1, 2, 3)
ops.begin_forward(dev)
y = ops.ws_conv3d(x, w, 1)
y.backward(dy)
torch.cuda.synchronize()
with torch.no_grad():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ops.begin_forward(dev)
    ops.ws_conv3d(x, w, 1)
    torch.cuda._sleep(int(0.02 * 1.9e9))
    e[0].record()
    for _ in range(reps):
        ops.ws_conv3d(x, w, 1)
    e[1].record()
    torch.cuda.synchronize()
vox = N * D * H * W
t = e[0].elapsed_time(e[1]) / reps
print(f"1x1 {cin}->{cout} on {vox} voxels: fprop {t * 1e3:.1f} us per launch (incl. weight standardisation), "
      f"{vox * (cin + cout) * 2 / t / 1e6:.0f} GB/s algorithmic")
