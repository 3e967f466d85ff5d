"""Host-side enqueue time per train step vs device time (are we launch-bound?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O
import multimodal_pl_b200 as mm
from multimodal_pl_b200.engine import DataParallelModel, FusedSGD
from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
from multimodal_pl_b200.unet3D import unet3D_baseline
dhw = tuple(int(v) for v in (sys.argv[1:4] or (64, 192, 192)))
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 2
model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda().train()
dp = DataParallelModel(model, 1); opt = FusedSGD(dp.parameters(), lr=1e-2, flat_grad=dp.flat_grad)
crit = EDiceLoss_partial(16)
x = torch.randn((batch, 1) + dhw, device="cuda").clamp_(-1, 1)
lab = torch.randint(0, 16, (batch, 1) + dhw, device="cuda").float()
w = [torch.ones(16)] * batch
def step():
    opt.zero_grad(); loss = crit(dp(x, lab)[0], lab.squeeze(1), mask=w); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"patch {dhw} batch {batch}: host enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
if os.environ.get("PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): step()
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
if os.environ.get("TPROF"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    busy = sum(e.time_range.elapsed_us() for e in ev)
    span = ev[-1].time_range.end - ev[0].time_range.start
    print(f"3 steps: GPU busy {busy/3e3:.2f} ms/step, span {span/3e3:.2f} ms/step, kernels {len(ev)/3:.0f}/step")
    gaps = []
    for a, b in zip(ev[:-1], ev[1:]):
        g = b.time_range.start - a.time_range.end
        if g > 50: gaps.append((g, a.name[:50], b.name[:50]))
    gaps.sort(reverse=True)
    print("largest idle gaps (us):")
    for g in gaps[:15]: print("  %8.0f  after %-50s before %s" % g)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=60))
