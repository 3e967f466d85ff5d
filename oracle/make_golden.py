"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the authoring container only:   python oracle/make_golden.py
The fixtures pin oracle/mmpl_oracle.py (tests/test_oracle_golden.py) and, through it and directly, the CUDA path
(tests/test_gpu_*.py).  Every fixture stores the seeds/shapes needed to regenerate its inputs with
oracle.mmpl_oracle's synthetic generators plus the reference's outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mmpl_oracle as O  # noqa: E402
from _refload import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def grad_summary(t: torch.Tensor):
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, 8).long()
    return np.concatenate([[f.norm().item(), f.sum().item()], f[idx].numpy()])


def main():
    torch.set_num_threads(os.cpu_count())
    ref_unet, ref_lp, ref_eval = load_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- 1. weight standardisation (unet3D.py:21-27)
    ws = {}
    for name, shp, seed in [("stem", (32, 1, 3, 3, 3), 1), ("c3", (64, 32, 3, 3, 3), 2), ("c1", (128, 64, 1, 1, 1), 3)]:
        conv = ref_unet.Conv3d(shp[1], shp[0], kernel_size=shp[2], padding=shp[2] // 2)
        w = torch.randn(shp, generator=torch.Generator().manual_seed(seed))
        conv.weight.data.copy_(w)
        x = torch.randn((1, shp[1], 4, 6, 6), generator=torch.Generator().manual_seed(seed + 10))
        ws[name + "_w"] = w.numpy()
        ws[name + "_x"] = x.numpy()
        ws[name + "_y"] = conv(x).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "ws_conv.npz"), **ws)

    # ---------------------------------------------------------------- 2. partial-label loss (loss_partial.py:71-99)
    loss_fix = {}
    cases = [
        ("ct_one_organ", (2, 16, 6, 10, 12), [1, 0, 0, 0, 1] + [0] * 11, 0),
        ("mri_bg_only", (1, 16, 4, 8, 8), [1] + [0] * 15, 1),
        ("all_ones", (2, 16, 4, 8, 8), None, 2),
        ("all_zero", (1, 16, 4, 8, 8), [0] * 16, 3),
        ("c4_frac", (3, 4, 5, 7, 9), [0.5, 1.0, 0.0, 2.0], 4),
    ]
    for name, shp, w, seed in cases:
        g = torch.Generator().manual_seed(100 + seed)
        z = (3 * torch.randn(shp, generator=g)).requires_grad_(True)
        tgt = torch.randint(0, shp[1], (shp[0],) + shp[2:], generator=g).float()
        mask = None if w is None else [torch.tensor(w, dtype=torch.float32)] * shp[0]
        for uce in (True, False):
            crit = ref_lp.EDiceLoss_partial(shp[1])
            L = crit(z, tgt, mask=mask, soft_max=True, uce=uce)
            gz, = torch.autograd.grad(L, z)
            tag = f"{name}_uce{int(uce)}"
            loss_fix[tag + "_loss"] = np.float64(L.item())
            loss_fix[tag + "_grad"] = gz.numpy()
        loss_fix[name + "_z"] = z.detach().numpy()
        loss_fix[name + "_t"] = tgt.numpy()
        loss_fix[name + "_w"] = np.array([1.0] * shp[1] if w is None else w, dtype=np.float32)
    # saturated logits: exercises the -100 log clamp of nn.BCELoss
    z = torch.zeros((1, 4, 2, 2, 2))
    z[:, 0] = 200.0
    z.requires_grad_(True)
    tgt = torch.ones((1, 2, 2, 2))
    L = ref_lp.EDiceLoss_partial(4)(z, tgt, mask=None)
    loss_fix["saturated_loss"] = np.float64(L.item())
    loss_fix["saturated_grad"] = torch.autograd.grad(L, z)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "partial_loss.npz"), **loss_fix)

    # ---------------------------------------------------------------- 3. unet3D_baseline fwd + loss + bwd
    for tag, shape, seed in [("b1", (1, 1, 16, 32, 32), 0), ("b2", (2, 1, 16, 16, 32), 1)]:
        sd = O.synth_state_dict(32, 16, seed)
        model = ref_unet.unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True)
        model.load_state_dict(sd)
        model.train()
        x = O.synth_patch(shape, 1000 + seed, "ct" if seed == 0 else "mri")
        lab = O.synth_labels((shape[0],) + shape[2:], 2000 + seed, 16, 32)
        w16 = [1, 0, 0, 0, 1] + [0] * 11 if seed == 0 else [1] + [0] * 7 + [1] + [0] * 7
        cmask = O.remap_unsupervised(lab, w16)
        logits = model(x, cmask)[0]
        crit = ref_lp.EDiceLoss_partial(16)
        L = crit(logits, cmask.squeeze(1), mask=[torch.tensor(w16, dtype=torch.float32)] * shape[0], soft_max=True)
        L.backward()
        fix = {
            "shape": np.array(shape), "seed": np.array(seed), "w16": np.array(w16, dtype=np.float32),
            "logits": logits.detach().numpy().astype(np.float32),
            "loss": np.float64(L.item()),
        }
        for k, p in model.named_parameters():
            fix["grad:" + k] = grad_summary(p.grad)
        fix["gradfull:conv1.weight"] = model.conv1.weight.grad.numpy()
        fix["gradfull:layer0.0.conv1.weight"] = model.layer0[0].conv1.weight.grad.numpy()
        fix["gradfull:layer1.0.downsample.2.weight"] = model.layer1[0].downsample[2].weight.grad.numpy()
        fix["gradfull:precls_conv.2.weight"] = model.precls_conv[2].weight.grad.numpy()
        fix["gradfull:layer0.0.gn1.weight"] = model.layer0[0].gn1.weight.grad.numpy()
        model.eval()
        with torch.no_grad():
            assert torch.equal(model(x), logits.detach())
        np.savez_compressed(os.path.join(OUT, f"unet_{tag}.npz"), **fix)
        print(tag, "loss", L.item())

    # ---------------------------------------------------------------- 4. sliding window + dice (evaluate_amos.py)
    sw = {}
    g = ref_eval._get_gaussian((8, 16, 16))
    sw["gauss_8_16_16"] = g
    gbig = ref_eval._get_gaussian((64, 192, 192))
    sw["gauss_big_stats"] = np.array([gbig.max(), gbig.min(), gbig.sum(dtype=np.float64), float((gbig == 0).sum())])
    sw["gauss_big_line"] = gbig[32, 96, :].copy()

    class TinyNet(torch.nn.Module):  # a cheap deterministic stand-in so the blend itself is pinned
        def __init__(self):
            super().__init__()
            self.c = torch.nn.Conv3d(1, 5, 3, padding=1)
            torch.manual_seed(7)
            torch.nn.init.normal_(self.c.weight, std=0.5)
            torch.nn.init.normal_(self.c.bias, std=0.5)

        def forward(self, img, task_id=None):
            return self.c(img)

    net = TinyNet().eval()
    vol = O.synth_patch((1, 1, 19, 37, 41), 5, "ct").numpy()
    with torch.no_grad():
        full = ref_eval.predict_sliding(None, [net], vol, (8, 16, 16), 5, None)
    sw["sw_vol_shape"] = np.array(vol.shape)
    sw["sw_out"] = full.numpy()
    sw["sw_w"] = net.c.weight.detach().numpy()
    sw["sw_b"] = net.c.bias.detach().numpy()
    lab = torch.randint(0, 5, (1, 1, 19, 37, 41), generator=torch.Generator().manual_seed(9)).float()
    dices, senc, spec, preds = ref_eval.get_dice(full, lab, None, num_class=4)
    sw["dice_labels"] = lab.numpy()
    sw["dice"] = np.array([float(d) for d in dices])
    sw["senc"] = np.array([float(d) for d in senc])
    sw["spec"] = np.array([float(d) for d in spec])
    sw["argmax"] = preds.numpy().astype(np.uint8)
    # tile-grid bookkeeping for the cfg4 volume, reproduced from predict_sliding's loop bounds
    from math import ceil

    image_size, tile = (300, 512, 512), (64, 192, 192)
    sHW, sD = ceil(tile[1] * 0.75), ceil(tile[0] * 0.75)
    n = [int(ceil((image_size[0] - tile[0]) / sD) + 1), int(ceil((image_size[1] - tile[1]) / sHW) + 1),
         int(ceil((image_size[2] - tile[2]) / sHW) + 1)]
    sw["cfg4_tiles"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "sliding_window.npz"), **sw)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
