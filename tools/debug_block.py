import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O
import multimodal_pl_b200 as mm
from multimodal_pl_b200.unet3D import NoBottleneck
ops = mm.ops
mm.set_compute_dtype(torch.float32); mm.set_conv_algo("direct")
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def R(shape, seed, scale=1.0): return scale * torch.randn(shape, generator=torch.Generator().manual_seed(seed))
blk = NoBottleneck(32, 32, weight_std=True).cuda()
sd = {}
for i, (k, v) in enumerate(blk.state_dict().items()):
    t = R(tuple(v.shape), 10 + i)
    if k.endswith("weight") and v.dim() == 1: t = 1 + 0.1 * t
    elif v.dim() == 1: t = 0.1 * t
    sd[k] = t
blk.load_state_dict(sd)
for shape, mean in [((1,32,16,32,32), 0.0), ((1,32,16,32,32), 5.0)]:
    x = R(shape, 1) + mean; dy = R(shape, 2)
    def run(dt):
        p = {"blk." + k: v.to(dt).clone().requires_grad_(True) for k, v in sd.items()}
        xx = x.to(dt).clone().requires_grad_(True)
        feats = {}
        out = O.gn_relu(xx, p["blk.gn1.weight"], p["blk.gn1.bias"]); feats["a1"] = out
        out = O.ws_conv3d(out, p["blk.conv1.weight"], 1, 1); feats["c1"] = out
        out = O.gn_relu(out, p["blk.gn2.weight"], p["blk.gn2.bias"]); feats["a2"] = out
        out = O.ws_conv3d(out, p["blk.conv2.weight"], 1, 1) + xx
        for t in feats.values(): t.retain_grad()
        (out * dy.to(dt)).sum().backward()
        return out, xx.grad, p, feats
    o64, g64, p64, f64 = run(torch.float64)
    o32, g32, p32, f32 = run(torch.float32)
    xd = x.cuda().requires_grad_(True)
    feats = {}
    def keep(n, t): t.retain_grad(); feats[n] = t; return t
    a1 = keep("a1", ops.gn_relu(xd, blk.gn1.weight, blk.gn1.bias))
    c1 = keep("c1", blk.conv1(a1))
    a2 = keep("a2", ops.gn_relu(c1, blk.gn2.weight, blk.gn2.bias))
    out = blk.conv2(a2, xd)
    blk.zero_grad(); out.backward(dy.cuda())
    print(f"block mean={mean}: out mine {rel(out,o64):.2e} t32 {rel(o32,o64):.2e} | dx mine {rel(xd.grad,g64):.2e} t32 {rel(g32,g64):.2e}")
    for n in ["a2", "c1", "a1"]:
        print(f"    grad {n}: mine {rel(feats[n].grad, f64[n].grad):.2e} t32 {rel(f32[n].grad, f64[n].grad):.2e}   fwd mine {rel(feats[n], f64[n]):.2e}")
    for k, prm in blk.named_parameters():
        print(f"    {k:14s} mine {rel(prm.grad, p64['blk.'+k].grad):.2e} t32 {rel(p32['blk.'+k].grad, p64['blk.'+k].grad):.2e}")
    for n in ["a1", "a2"]:
        m_mine = (feats[n].detach().cpu() > 0); m64 = (f64[n].detach() > 0); m32 = (f32[n].detach() > 0)
        print(f"    gate flips {n}: mine-vs-f64 {(m_mine != m64).sum().item()}  t32-vs-f64 {(m32 != m64).sum().item()}  of {m64.numel()}")
