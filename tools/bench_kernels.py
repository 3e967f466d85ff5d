"""Time the bandwidth-bound kernels one by one at cfg2 full-resolution sizes (CUDA events, after warm-up) and report
achieved GB/s against their algorithmic bytes.  Usage: python tools/bench_kernels.py [filter]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_pl_b200 as mm
from multimodal_pl_b200 import ops, _lib

_lib.require_device()
mm.set_compute_dtype(torch.bfloat16)
dev = torch.device("cuda")
flt = sys.argv[1] if len(sys.argv) > 1 else ""
N, D, H, W = 2, 64, 192, 192
V = N * D * H * W
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


WARM, REPS = int(os.environ.get("MMPL_BK_WARM", "3")), int(os.environ.get("MMPL_BK_REPS", "10"))


def timeit(name, fn, nbytes, reps=None):
    reps = reps or REPS
    if flt and flt not in name:
        return
    for _ in range(WARM):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"{name:44s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:8.0f} GB/s  ({nbytes / 1e6:.0f} MB)", flush=True)


def cl(c, n=N, d=D, h=H, w=W, dtype=torch.bfloat16):
    return torch.randn((n, d, h, w, c), device=dev, dtype=torch.float32).to(dtype).permute(0, 4, 1, 2, 3)


A = 2
for C, lvl in [(32, 0), (64, 1)]:
    d, h, w = D >> lvl, H >> lvl, W >> lvl
    v = N * d * h * w
    x = cl(C, N, d, h, w).requires_grad_(True)
    g1, b1 = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
    g2, b2 = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
    dy = cl(C, N, d, h, w)
    L = _lib.lib()
    stats = torch.zeros(N * 32, dtype=torch.float64, device=dev)
    timeit(f"gn_stats C={C} L{lvl}", lambda: L.mmpl_gn_stats(x.data_ptr(), stats.data_ptr(), N, d * h * w, C, 16, 1, _lib.stream_ptr()), v * C * A)
    y = ops.gn_relu(x, g1, b1)
    timeit(f"gn_relu fwd C={C} L{lvl}", lambda: ops.gn_relu(x, g1, b1), 3 * v * C * A)
    timeit(f"gn_relu bwd C={C} L{lvl}", lambda: torch.autograd.grad(y, x, dy, retain_graph=True), 5 * v * C * A)
    y1, y2 = ops.gn_relu_dual(x, g1, b1, g2, b2)
    timeit(f"gn_relu_dual fwd C={C} L{lvl}", lambda: ops.gn_relu_dual(x, g1, b1, g2, b2), 4 * v * C * A)
    timeit(f"gn_relu_dual bwd C={C} L{lvl}", lambda: torch.autograd.grad([y1, y2], x, [dy, dy], retain_graph=True), 7 * v * C * A)
    ya, xa = ops.gn_relu(x, g1, b1, alias=True)
    timeit(f"gn_relu bwd+addend C={C} L{lvl}", lambda: torch.autograd.grad([ya, xa], x, [dy, dy], retain_graph=True), 6 * v * C * A)

# upsample L1 -> L0
xlo = cl(32, N, D // 2, H // 2, W // 2).requires_grad_(True)
skip = cl(32)
dy = cl(32)
yu = ops.upsample2x_add(xlo, skip)
timeit("upsample2x_add fwd C=32 ->L0", lambda: ops.upsample2x_add(xlo, skip), V * 32 * A * (2 + 1 / 8))
timeit("upsample2x bwd C=32 ->L0", lambda: torch.autograd.grad(yu, xlo, dy, retain_graph=True), V * 32 * A * (1 + 1 / 8))

# classifier + loss
a = cl(32).requires_grad_(True)
wc = torch.randn(16, 32, 1, 1, 1, device=dev, requires_grad=True)
bc = torch.zeros(16, device=dev, requires_grad=True)
logits = ops.classifier(a, wc, bc)
dl = torch.randn_like(logits)
timeit("cls fwd", lambda: ops.classifier(a, wc, bc), V * (32 * A + 64))
timeit("cls bwd", lambda: torch.autograd.grad(logits, a, dl, retain_graph=True), V * (64 + 32 * A + 32 * A))
z = torch.randn(N, 16, D, H, W, device=dev).requires_grad_(True)
tgt = torch.randint(0, 16, (N, D, H, W), device=dev).float()
cw = torch.tensor([1.0, 0, 0, 0, 1.0] + [0.0] * 11, device=dev)
loss = ops.partial_label_loss(z, tgt, cw)
timeit("partial_loss fwd", lambda: ops.partial_label_loss(z, tgt, cw), V * 68)
timeit("partial_loss bwd", lambda: torch.autograd.grad(loss, z, retain_graph=True), V * 132)
cw1 = torch.ones(16, device=dev)
loss1 = ops.partial_label_loss(z, tgt, cw1)
timeit("partial_loss fwd (all classes)", lambda: ops.partial_label_loss(z, tgt, cw1), V * 68)
timeit("partial_loss bwd (all classes)", lambda: torch.autograd.grad(loss1, z, retain_graph=True), V * 132)

# classifier + loss fused (csrc/cls_loss.cu): per-sample CT/MRI weights + LUT and uint8 labels like bench.py
tgt8 = tgt.to(torch.uint8)
cw2 = torch.tensor([[1.0, 0, 0, 0, 1.0] + [0.0] * 11, [1.0] + [0.0] * 15], device=dev)
lut2 = torch.stack([torch.tensor([float(l) if (l == 0 or w[l]) else 0.0 for l in range(16)]) for w in cw2.tolist()]).to(dev)
lossf = ops.classifier_partial_loss(a, wc, bc, tgt8, cw2, lut2, True, True)
timeit("head fused fwd (cls+loss)", lambda: ops.classifier_partial_loss(a, wc, bc, tgt8, cw2, lut2, True, True), V * (32 * A + 1))
timeit("head fused bwd (cls+loss)", lambda: torch.autograd.grad(lossf, a, retain_graph=True), V * (64 * A + 1))
loss2 = ops.partial_label_loss(ops.classifier(a, wc, bc), tgt8, cw2, lut2, True, True)
timeit("head two-step fwd (cls, loss)", lambda: ops.partial_label_loss(ops.classifier(a, wc, bc), tgt8, cw2, lut2, True, True), V * (32 * A + 1))
timeit("head two-step bwd (loss, cls)", lambda: torch.autograd.grad(loss2, a, retain_graph=True), V * (64 * A + 1))
lossa = ops.classifier_partial_loss(a, wc, bc, tgt8, torch.ones(2, 16, device=dev), None, True, True)
timeit("head fused fwd (all classes)", lambda: ops.classifier_partial_loss(a, wc, bc, tgt8, torch.ones(2, 16, device=dev), None, True, True), V * (32 * A + 1))
timeit("head fused bwd (all classes)", lambda: torch.autograd.grad(lossa, a, retain_graph=True), V * (64 * A + 1))

# stem
img = torch.randn(N, 1, D, H, W, device=dev)
ws = torch.randn(32, 1, 3, 3, 3, device=dev, requires_grad=True)
ys = ops.stem_conv(img, ws)
dys = cl(32)
timeit("stem fwd (im2col + conv)", lambda: ops.stem_conv(img, ws), V * (4 + 32 * A))
timeit("stem bwd (wgrad)", lambda: torch.autograd.grad(ys, ws, dys, retain_graph=True), V * (4 + 32 * A))
