"""torchrun worker: sliding-window tiles dealt round-robin over the ranks + one all-reduce of the accumulators must give
the same argmax / Dice as a single rank (tests/test_gpu_train.py::test_sharded_sliding_window)."""
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
dist.barrier()
if rank == 0:
    print("SW_OK", mism)
dist.destroy_process_group()
