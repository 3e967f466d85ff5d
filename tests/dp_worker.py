"""torchrun worker for tests/test_gpu_train.py::test_nccl_data_parallel_equals_single_process (fp32 exact path):
2 ranks x batch 1 with the gradient all-reduce must reproduce the gradient of the mean of the two per-sample losses."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmpl_oracle as O  # noqa: E402
import multimodal_pl_b200 as mm  # noqa: E402
from multimodal_pl_b200.engine import DataParallelModel  # noqa: E402
from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial  # noqa: E402
from multimodal_pl_b200.unet3D import unet3D_baseline  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
mm.set_compute_dtype(torch.float32)
mm.set_conv_algo("direct")
sd = O.synth_state_dict(32, 16, 0)
crit = EDiceLoss_partial(16)
xs = [O.synth_patch((1, 1, 16, 16, 32), 50 + r, "ct").cuda() for r in range(world)]
ls = [O.synth_labels((1, 16, 16, 32), 60 + r, 16, 32).cuda() for r in range(world)]
model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
model.load_state_dict(sd)
dp = DataParallelModel(model, world, bucket_mb=8)
dp.zero_grad()
crit(dp(xs[rank])[0], ls[rank].squeeze(1), mask=[torch.ones(16)]).backward()
torch.cuda.synchronize()
ref = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
ref.load_state_dict(sd)
total = sum(crit(ref(xs[r])[0], ls[r].squeeze(1), mask=[torch.ones(16)]) for r in range(world)) / world
total.backward()
flat_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
err = ((dp.flat_grad - flat_ref).norm() / flat_ref.norm()).item()
assert err < 1e-4, err
dist.barrier()
if rank == 0:
    print("DP_OK", err)
dist.destroy_process_group()
