"""Train-step level checks on the device: fused SGD == torch.optim.SGD on the model, loss decreases over a few steps
of the bf16 tcgen05 path, and (when >= 2 GPUs are visible) the NCCL data-parallel step equals the single-process
step on the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

import mmpl_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_sgd_matches_torch_sgd_on_model():
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import DataParallelModel, FusedSGD
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.float32)
    mm.set_conv_algo("direct")
    try:
        sd = O.synth_state_dict(32, 16, 0)
        x = O.synth_patch((1, 1, 16, 16, 32), 5, "ct").cuda()
        lab = O.synth_labels((1, 16, 16, 32), 6, 16, 32).cuda()
        w = [torch.ones(16)]
        crit = EDiceLoss_partial(16)
        a = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
        b = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
        a.load_state_dict(sd)
        b.load_state_dict(sd)
        opt_a = torch.optim.SGD(a.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        dp = DataParallelModel(b, 1)
        opt_b = FusedSGD(dp.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, flat_grad=dp.flat_grad)
        for _ in range(3):
            opt_a.zero_grad()
            crit(a(x)[0], lab.squeeze(1), mask=w).backward()
            opt_a.step()
            opt_b.zero_grad()
            crit(dp(x)[0], lab.squeeze(1), mask=w).backward()
            opt_b.step()
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert torch.allclose(pa, pb, rtol=5e-3, atol=2e-5), k   # 3 steps of fp32 atomics + ReLU-gate noise; the kernel itself is exact in test_sgd_step
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


def test_fused_sgd_skips_parameters_without_gradient_like_torch_sgd():
    """torch.optim.SGD leaves a parameter whose .grad is None untouched (no weight decay, no momentum): in
    unet3D_with_feam3 the eam*.proj parameters never receive a gradient.  FusedSGD steps only the runs of the flat buffer
    that have one, including a run that starts off a 16-byte boundary (the 14-element bias)."""
    from multimodal_pl_b200.engine import FusedSGD

    torch.manual_seed(0)
    shapes = [(6,), (14,), (3, 5), (9,), (32,)]           # element offsets 0, 6, 20, 35, 44
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    opt_a = torch.optim.SGD(pa, lr=0.05, momentum=0.9, weight_decay=1e-2)
    opt_b = FusedSGD(pb, lr=0.05, momentum=0.9, weight_decay=1e-2)
    for step in range(3):
        opt_a.zero_grad(set_to_none=True)
        opt_b.zero_grad()
        live = [0, 1, 3] if step < 2 else [1, 3, 4]        # parameter 2 never gets a gradient, 4 only at the end
        for i in live:
            g = torch.randn(shapes[i], device="cuda", generator=torch.Generator("cuda").manual_seed(10 * step + i))
            pa[i].grad = g.clone()
            pb[i].grad = g.clone()
        opt_a.step()
        opt_b.step()
    for i, (a, b) in enumerate(zip(pa, pb)):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), i


def test_backward_refuses_standardised_weights_refreshed_from_changed_weights():
    """The standardised-weight buffers a convolution saved for backward are per-parameter and rewritten in place by the
    next forward.  Unchanged weights -> same values -> a second forward before the backward is fine (GAN-style loops do
    that); weights changed in between -> backward must raise like stock PyTorch, not silently use the new weights."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.unet3D import Conv3d

    mm.set_compute_dtype(torch.bfloat16)
    conv = Conv3d(32, 32, 3, padding=1).cuda()
    x = torch.randn(1, 32, 4, 8, 8, device="cuda").requires_grad_(True)
    y1 = conv(x)
    y2 = conv(x)                       # same weights: harmless refresh
    (y1.float().sum() + y2.float().sum()).backward()
    assert torch.isfinite(conv.weight.grad).all()
    y1 = conv(x)
    with torch.no_grad():
        conv.weight.mul_(1.5)          # an optimizer step between forward and backward ...
    conv(x)                            # ... and a forward that re-standardises into the same buffers
    with pytest.raises(RuntimeError, match="between its forward and its backward"):
        y1.float().sum().backward()


def test_parameter_gradients_land_in_the_flat_buffer_without_copies():
    """Backward kernels write every parameter gradient straight into its slot of the flat buffer and autograd adopts
    the view (no per-parameter add/copy launches); a second backward before zero_grad accumulates (p.grad += new)."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import DataParallelModel, FusedSGD
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
    dp = DataParallelModel(model, 1)
    opt = FusedSGD(dp.parameters(), lr=0.0, momentum=0.0, weight_decay=0.0, flat_grad=dp.flat_grad)
    x = O.synth_patch((1, 1, 16, 32, 32), 5, "ct").cuda()
    lab = O.synth_labels((1, 16, 32, 32), 6, 16, 32).cuda()
    crit = EDiceLoss_partial(16)
    opt.zero_grad()
    assert all(p.grad is None for p in model.parameters())
    crit(dp(x, lab)[0], lab.squeeze(1), mask=[torch.ones(16)]).backward()
    base = dp.flat_grad.data_ptr()
    for k, p in model.named_parameters():
        _, off, n = p._mmpl_grad_slot
        assert p.grad is not None and p.grad.data_ptr() == base + 4 * off, k
    once = dp.flat_grad.clone()
    assert once.abs().max().item() > 0
    crit(dp(x, lab)[0], lab.squeeze(1), mask=[torch.ones(16)]).backward()      # accumulate
    assert torch.allclose(dp.flat_grad, 2 * once, rtol=1e-3, atol=1e-6 * once.abs().max().item())
    opt.zero_grad()
    assert dp.flat_grad.abs().max().item() == 0 and all(p.grad is None for p in model.parameters())


def test_graphed_train_step_matches_eager_steps():
    """engine.GraphedTrainStep (what bench.py times): constructor warm-up step + 2 replays == 3 eager steps -- same losses
    and parameters up to the summation-order noise of the fp32/fp64 atomics (the captured graph contains the side-stream
    weight-gradient branch, the batched weight standardisation and the fused SGD)."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import DataParallelModel, FusedSGD, GraphedTrainStep
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.bfloat16)
    sd = O.synth_state_dict(32, 16, 0)
    x = O.synth_patch((2, 1, 16, 32, 32), 5, "ct").cuda()
    lab = O.synth_labels((2, 16, 32, 32), 6, 16, 32).cuda()
    crit = EDiceLoss_partial(16)
    w = [torch.ones(16)] * 2

    def loss_fn(logits, l):
        return crit(logits, l.squeeze(1), mask=w)

    def make():
        m = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
        m.load_state_dict(sd)
        m.train()
        dp = DataParallelModel(m, 1)
        return m, dp, FusedSGD(dp.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, flat_grad=dp.flat_grad)

    ma, dpa, opta = make()
    eager_losses = []
    for _ in range(3):
        opta.zero_grad()
        loss = loss_fn(dpa(x, lab)[0], lab)
        loss.backward()
        opta.step()
        eager_losses.append(loss.item())
    mb, dpb, optb = make()
    step = GraphedTrainStep(dpb, loss_fn, optb, x, lab, warmup=1)
    graph_losses = [step(x, lab).item() for _ in range(2)]
    assert abs(graph_losses[0] - eager_losses[1]) < 2e-3 * abs(eager_losses[1])
    assert abs(graph_losses[1] - eager_losses[2]) < 2e-3 * abs(eager_losses[2])
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert torch.allclose(pa, pb, rtol=2e-2, atol=2e-4), k


def test_graphed_inference_matches_eager_forward():
    """engine.GraphedInference (sliding-window tiles): the replayed forward with frozen standardised weights equals the
    eager eval-mode forward; other input shapes fall back to the eager module."""
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import GraphedInference
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.bfloat16)
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
    model.load_state_dict(O.synth_state_dict(32, 16, 1))
    model.eval()
    x1 = O.synth_patch((1, 1, 16, 32, 32), 7, "ct").cuda()
    x2 = O.synth_patch((1, 1, 16, 32, 32), 8, "mri").cuda()
    with torch.no_grad():
        r1, r2 = model(x1).clone(), model(x2).clone()
    net = GraphedInference(model, x1)
    for xin, ref in ((x1, r1), (x2, r2), (x1, r1)):
        out = net(xin, None)
        assert ((out - ref).norm() / ref.norm()).item() < 1e-3
    small = O.synth_patch((1, 1, 16, 16, 32), 9, "ct").cuda()
    with torch.no_grad():
        ref_small = model(small)
    assert ((net(small, None) - ref_small).norm() / ref_small.norm()).item() < 1e-3


def test_bf16_training_reduces_loss():
    import multimodal_pl_b200 as mm
    from multimodal_pl_b200.engine import DataParallelModel, FusedSGD
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).cuda()
    dp = DataParallelModel(model, 1)
    opt = FusedSGD(dp.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, flat_grad=dp.flat_grad)
    x = O.synth_patch((2, 1, 16, 32, 32), 5, "ct").cuda()
    lab = O.synth_labels((2, 16, 32, 32), 6, 16, 32).cuda()
    crit = EDiceLoss_partial(16)
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = crit(dp(x, lab)[0], lab.squeeze(1), mask=[torch.ones(16)] * 2)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < 0.8 * losses[0], losses


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_data_parallel_equals_single_process():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tests", "dp_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DP_OK" in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_sliding_window():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29612",
                          os.path.join(ROOT, "tests", "sw_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SW_OK" in out.stdout
