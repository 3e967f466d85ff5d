import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O
import multimodal_pl_b200 as mm
from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
from multimodal_pl_b200.unet3D import unet3D_baseline

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

for dtype, algo in [(torch.float32, "direct")]:
    mm.set_compute_dtype(dtype); mm.set_conv_algo(algo)
    tag = "b1"
    g = np.load(os.path.join(ROOT, "tests/golden", f"unet_{tag}.npz"))
    shape, seed = tuple(int(v) for v in g["shape"]), int(g["seed"])
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
    model.load_state_dict(O.synth_state_dict(32, 16, seed)); model.train()
    x = O.synth_patch(shape, 1000 + seed, "ct")
    lab = O.synth_labels((shape[0],) + shape[2:], 2000 + seed, 16, 32)
    w16 = g["w16"].tolist(); cmask = O.remap_unsupervised(lab, w16)
    feats = {}
    def fwd(self, input, mask=None):
        def keep(n, t):
            t.retain_grad(); feats[n] = t; return t
        x = keep("stem", self.conv1(input)); x = keep("layer0", self.layer0(x)); skip0 = x
        x = keep("layer1", self.layer1(x)); skip1 = x
        x = keep("layer2", self.layer2(x)); skip2 = x
        x = keep("layer3", self.layer3(x)); skip3 = x
        x = keep("layer4", self.layer4(x)); x = keep("fusion", self.fusionConv(x))
        x = keep("x8", self.x8_resb(keep("up8", self.upsamplex2(x, skip3))))
        x = keep("x4", self.x4_resb(keep("up4", self.upsamplex2(x, skip2))))
        x = keep("x2", self.x2_resb(keep("up2", self.upsamplex2(x, skip1))))
        x = keep("x1", self.x1_resb(keep("up1", self.upsamplex2(x, skip0))))
        return self.precls_conv(x), [], []
    logits, _, _ = fwd(model, x.cuda(), cmask.cuda())
    L = EDiceLoss_partial(16)(logits, cmask.squeeze(1).cuda(), mask=[torch.tensor(w16)], soft_max=True)
    L.backward()
    print(f"== {dtype} {algo}: logits rel {rel(logits, torch.from_numpy(g['logits'])):.3e} loss {L.item():.6f} ref {float(g['loss']):.6f}")
    # oracle grads on CPU for full comparison
    sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(32, 16, seed).items()}
    lo = O.unet3d_forward(sd, x); Lo = O.partial_label_loss(lo, cmask.squeeze(1), w16); Lo.backward()
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in O.synth_state_dict(32, 16, seed).items()}
    f64 = {}
    l64 = O.unet3d_forward(sd64, x.double(), feats=f64)
    for t in f64.values(): t.retain_grad()
    L64 = O.partial_label_loss(l64, cmask.squeeze(1).double(), w16); L64.backward()
    print(f"   logits: mine-vs-f64 {rel(logits, l64):.3e}  torchfp32-vs-f64 {rel(lo, l64):.3e}")
    for n in f64:
        print(f"   feat {n:8s} fwd {rel(feats[n], f64[n]):.2e}  grad {rel(feats[n].grad, f64[n].grad):.2e}  |grad| {f64[n].grad.norm().item():.3e}")
    for k, p in model.named_parameters():
        r = rel(p.grad, sd64[k].grad); r32 = rel(sd[k].grad, sd64[k].grad)
        if r > (1e-4 if dtype == torch.float32 else 3e-2):
            print(f"   {k:40s} mine-vs-f64 {r:.3e}  torchfp32-vs-f64 {r32:.3e}")
