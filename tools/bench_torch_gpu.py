"""The reference's own ATen/cuDNN path on the B200 (the 'library kernel to beat'): the oracle restatement
(identical ops to unet3D.py + loss_partial.py) run on the GPU in eager PyTorch, fp32 (TF32 off) and bf16 autocast +
channels_last_3d, cfg2 shapes.  Prints patches/s.  Test/bench infrastructure only (imports oracle/)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B, dhw, base, classes = 2, (64, 192, 192), 32, 16
sd = {k: v.to(dev).requires_grad_(True) for k, v in O.synth_state_dict(base, classes, 0).items()}
x = torch.cat([O.synth_patch((1, 1) + dhw, 100 + b, "ct") for b in range(B)]).to(dev)
lab = torch.randint(0, classes, (B,) + dhw, device=dev).float()
w = [1.0, 0, 0, 0, 1.0] + [0.0] * (classes - 5)
opt = torch.optim.SGD(list(sd.values()), lr=1e-2, momentum=0.9, weight_decay=1e-4)


def step(autocast, cl):
    opt.zero_grad(set_to_none=True)
    xi = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = O.unet3d_forward(sd, xi, base)
    loss = O.partial_label_loss(logits.float(), lab, w)
    loss.backward()
    opt.step()
    return loss


for name, ac, cl in [("fp32 (TF32 off)", False, False), ("bf16 autocast + channels_last_3d", True, True)]:
    try:
        for _ in range(3):
            step(ac, cl)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            step(ac, cl)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"torch eager {name}: {dt * 1e3:.1f} ms/step, {B / dt:.1f} patches/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"torch eager {name}: failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
