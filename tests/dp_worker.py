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

# ---- the benchmarked path: CUDA-graph step with the NCCL exchange captured inside the graph, per-bucket SGD.
# After 3 optimisation steps the weights must equal those of a single process that averages the per-rank gradients.
from multimodal_pl_b200.engine import FusedSGD, GraphedTrainStep  # noqa: E402


def loss_fn(logits, lab):
    return crit(logits, lab.squeeze(1), mask=[torch.ones(16)])


m2 = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
m2.load_state_dict(sd)
dp2 = DataParallelModel(m2, world, bucket_mb=2, average=False)
opt2 = FusedSGD(dp2.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4, flat_grad=dp2.flat_grad)
step = GraphedTrainStep(dp2, loss_fn, opt2, xs[rank], ls[rank], warmup=1)      # 1 eager step inside
assert step.comm_in_graph, "NCCL capture failed: the exchange fell back outside the graph"
for _ in range(2):
    step(xs[rank], ls[rank])
torch.cuda.synchronize()
m3 = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
m3.load_state_dict(sd)
opt3 = torch.optim.SGD(m3.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
for _ in range(3):
    opt3.zero_grad()
    (sum(loss_fn(m3(xs[r])[0], ls[r]) for r in range(world)) / world).backward()
    opt3.step()
worst = max(((a - b).norm() / b.norm().clamp_min(1e-12)).item() for a, b in zip(m2.parameters(), m3.parameters()))
assert worst < 2e-3, worst      # 3 steps of fp32 atomics + ReLU-gate noise, as in test_fused_sgd_matches_torch_sgd_on_model
step.close()                    # a live CUDA graph that holds NCCL kernels must go before the process group does
dist.barrier()
if rank == 0:
    print("DP_OK", err, worst)
dist.destroy_process_group()
