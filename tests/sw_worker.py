"""torchrun worker (tests/test_gpu_train.py::test_sharded_sliding_window): the sharded sliding window must give the same
argmax / Dice as a single rank -- generic path (tiles split over the ranks, fp64 accumulators all-reduced) and production
path (bf16 unet3D_baseline, fused classifier+blend, exchange of the touched accumulator planes along depth, local
finalize, all-gather of the mask)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O  # noqa: E402
from multimodal_pl_b200.evaluate import predict_sliding_dice  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
net = torch.nn.Conv3d(1, 6, 3, padding=1).cuda().eval()
torch.manual_seed(3)
torch.nn.init.normal_(net.weight, std=0.5)
torch.nn.init.normal_(net.bias, std=0.5)
for p in net.parameters():
    dist.broadcast(p.data, src=0)
vol = O.synth_patch((1, 1, 40, 70, 90), 4, "ct").numpy()
lab = torch.randint(0, 6, (1, 1, 40, 70, 90), generator=torch.Generator().manual_seed(5)).float()
f = [lambda im, tid: net(im)]
d_sh, _, _, am_sh = predict_sliding_dice(None, f, vol, (16, 32, 32), 6, None, label=lab, num_class=5, sharded=True)
d_1, _, _, am_1 = predict_sliding_dice(None, f, vol, (16, 32, 32), 6, None, label=lab, num_class=5, sharded=False)
# fp64 accumulation: the only difference is the summation order of the tile contributions per voxel
mism = (am_sh != am_1).sum().item()
assert mism <= 2, mism
assert max(abs(float(a) - float(b)) for a, b in zip(d_sh, d_1)) < 1e-5
# ---- production path: plane exchange along depth (40 planes / 2 ranks; 41 planes exercises the padded slab)
import multimodal_pl_b200 as mm  # noqa: E402
from multimodal_pl_b200.engine import GraphedSlidingWindow  # noqa: E402
from multimodal_pl_b200.unet3D import unet3D_baseline  # noqa: E402

mm.set_compute_dtype(torch.bfloat16)
model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda().eval()
model.load_state_dict(O.synth_state_dict(32, 16, 2))
for depth in (40, 41):
    vol = O.synth_patch((1, 1, depth, 72, 88), 41, "ct")
    lab = O.synth_labels((1, depth, 72, 88), 42, 16, 32).to(torch.uint8)
    kw = dict(label=lab, acc_dtype=torch.float32, num_class=15)
    one = predict_sliding_dice(None, [model], vol, (16, 32, 32), 16, None, sharded=False, **kw)
    two = predict_sliding_dice(None, [model], vol, (16, 32, 32), 16, None, sharded=True, **kw)
    eng = GraphedSlidingWindow(model, (depth, 72, 88), (16, 32, 32), 16, world_size=world)
    thr = predict_sliding_dice(None, [eng], vol, (16, 32, 32), 16, None, sharded=True, **kw)
    # a second volume through the same object: only the planes a rank writes or owns are re-zeroed between volumes
    vol2 = O.synth_patch((1, 1, depth, 72, 88), 43, "ct")
    predict_sliding_dice(None, [eng], vol2, (16, 32, 32), 16, None, sharded=True, **kw)
    again = predict_sliding_dice(None, [eng], vol, (16, 32, 32), 16, None, sharded=True, **kw)
    assert torch.equal(again[3], thr[3])
    for got in (two, thr):
        # fp32 sums of <= 8 tile contributions in a different order: only exact near-ties may flip
        bad = (got[3] != one[3]).sum().item()
        assert bad <= 5, (depth, bad)
        assert max(abs(float(a) - float(b)) for a, b in zip(got[0], one[0])) < 1e-4
dist.barrier()
if rank == 0:
    print("SW_OK", mism)
dist.destroy_process_group()
