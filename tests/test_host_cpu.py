"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol declared in include/mmpl_b200.h, the
drop-in modules keep the reference's state_dict / signatures, host-side logic (tile grid, mask table, tile sharding,
gradient bucketing over a 2-process gloo group) behaves like the oracle.  No compute kernels are called."""
import inspect
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mmpl_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from multimodal_pl_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "mmpl_b200.h")).read()
    declared = set(re.findall(r"\b(mmpl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    l = _lib.lib()
    for name in sorted(declared):
        assert hasattr(l, name), f"{name} declared in include/mmpl_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert l.mmpl_version() >= 100
    assert _lib.launch_count() == 0            # nothing may have launched on a CPU-only host


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from multimodal_pl_b200.unet3D import unet3D_baseline

    net = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 16, 32, 32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal-pl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "mmpl_oracle" not in src and "import oracle" not in src, f


def test_state_dict_contract_and_signatures():
    from multimodal_pl_b200 import unet3D
    from multimodal_pl_b200.loss_functions import loss_partial, losses

    net = unet3D.unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True)
    sd, shapes = net.state_dict(), O.state_dict_shapes(32, 16)
    assert len(sd) == 107 and set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert sum(p.numel() for p in net.parameters()) == 17286512
    net.load_state_dict(O.synth_state_dict(32, 16, 0))          # reference-shaped checkpoints load unchanged
    sig = inspect.signature
    assert list(sig(unet3D.unet3D_baseline.__init__).parameters)[1:7] == \
        ["layers", "num_classes", "weight_std", "ema", "use_cm", "deep_up"]
    assert list(sig(unet3D.unet3D_baseline.forward).parameters) == ["self", "input", "mask"]
    assert list(sig(unet3D.NoBottleneck.__init__).parameters)[1:] == \
        ["inplanes", "planes", "stride", "dilation", "downsample", "fist_dilation", "multi_grid", "weight_std", "group"]
    assert list(sig(unet3D.Conv3d.__init__).parameters)[1:] == \
        ["in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups", "bias"]
    assert list(sig(loss_partial.EDiceLoss_partial.forward).parameters)[:6] == \
        ["self", "inputs", "target", "mask", "soft_max", "uce"]
    assert list(sig(losses.get_loss).parameters)[:5] == ["output", "cm", "deep_out", "target", "mask"]
    with pytest.raises(NotImplementedError):
        unet3D.Conv3d(32, 32, kernel_size=(5, 5, 5), padding=(2, 2, 2))


def test_feam3_contract_eam_and_tokens_on_cpu(golden_dir):
    """unet3D_with_feam3: reference state_dict keys/shapes (143 tensors), constructor signature, and the parts that are
    plain tensor ops and therefore run on the CPU -- the EAM module against its formula, and the folding of its head-mean
    attention logits into one 15-row matrix."""
    from multimodal_pl_b200 import unet3D

    net = unet3D.unet3D_with_feam3([1, 2, 2, 2, 2], num_classes=16, weight_std=True)
    sd_ref, tokens = O.synth_feam3_state_dict(16, 3)
    sd = net.state_dict()
    assert len(sd) == 143 and set(sd) == set(sd_ref)
    assert all(tuple(sd[k].shape) == tuple(sd_ref[k].shape) for k in sd_ref)
    net.load_state_dict(sd_ref)
    assert list(inspect.signature(unet3D.unet3D_with_feam3.__init__).parameters)[1:7] == \
        ["layers", "num_classes", "weight_std", "ema", "use_cm", "deep_up"]
    # EAM: attn = q k^T (unscaled), per head; returned x = proj(norm2(softmax(attn*scale) v)) + softmax(...) v
    eam = net.eam21
    g = torch.Generator().manual_seed(1)
    x = torch.randn((1, 40, 32), generator=g)
    tok = torch.randn((1, 15, 32), generator=g)
    out, attn = eam(x, tok)
    xn = torch.nn.functional.layer_norm(x, (32,), eam.norm2.weight, eam.norm2.bias)
    tn = torch.nn.functional.layer_norm(tok, (32,), eam.norm3.weight, eam.norm3.bias)
    kv = xn @ eam.kv.weight.t()
    k, v = kv[..., :32].reshape(1, 40, 4, 8), kv[..., 32:].reshape(1, 40, 4, 8)
    q = (tn @ eam.q.weight.t()).reshape(1, 15, 4, 8)
    ref_attn = torch.einsum("bthd,bnhd->bhtn", q, k)
    assert tuple(attn.shape) == (1, 4, 15, 40) and torch.allclose(attn, ref_attn, atol=1e-5)
    av = torch.einsum("bhtn,bnhd->bthd", torch.softmax(ref_attn * eam.scale, -1), v).reshape(1, 15, 32)
    ref_out = torch.nn.functional.layer_norm(av, (32,), eam.norm2.weight, eam.norm2.bias) @ eam.proj.weight.t() \
        + eam.proj.bias + av
    assert torch.allclose(out, ref_out, atol=1e-5)
    # the head mean of the per-head logits -- all the model consumes (:1133-1137) -- as ONE 15-row matrix over the
    # LayerNorm-ed rows: what EAM.attention_map feeds the classifier kernel with (csrc/eam.cu)
    w, b = eam._folded(tok[0])
    xhat = torch.nn.functional.layer_norm(x, (32,), eps=eam.norm2.eps)
    assert torch.allclose(xhat[0] @ w.t() + b, ref_attn.mean(1)[0].t(), atol=1e-5)
    # (renew_token runs on the device: tests/test_gpu_more.py checks it against the tokens the reference produced)


def test_aux_nets_state_dict_contract():
    """unet3D_g / norm_style_discriminator_output: the reference's state_dict keys and shapes, recorded in
    tests/golden/aux_nets.npz by oracle/make_golden_aux.py (one gradient-norm entry per reference parameter)."""
    from multimodal_pl_b200.aux_nets import norm_style_discriminator_output, unet3D_g

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_nets.npz"))
    ref_keys = {k.split("norm:")[1] for k in g.files if k.startswith("refiner/norm:")}
    net = unet3D_g([1, 1, 1, 1, 1], num_classes=2, weight_std=True, init_filter=24, in_channel=2)
    assert {k for k, _ in net.named_parameters()} == ref_keys
    assert tuple(net.conv0.weight.shape) == (24, 2, 3, 3, 3) and net.fusionConv[0].num_groups == 12
    assert net.layer1[0].gn1.num_groups == 4 and net.precls_conv[0].num_groups == 6
    assert list(inspect.signature(unet3D_g.__init__).parameters)[1:] == \
        ["layers", "num_classes", "weight_std", "in_channel", "init_filter"]
    dis = norm_style_discriminator_output(num_classes=2)
    assert {k for k, _ in dis.named_parameters()} == {k.split("norm:")[1] for k in g.files if k.startswith("disc/norm:")}
    assert tuple(dis.block1[0].weight.shape) == (32, 2, 4, 4, 4) and tuple(dis.block4[8].weight.shape) == (2, 256)


def test_flat_gradient_adoption_on_cpu():
    """engine.DataParallelModel keeps gradients in one flat buffer: gradients that autograd produced elsewhere are
    folded into their slots (engine._adopt_grad), zero_grad drops the views and clears the buffer."""
    from multimodal_pl_b200.engine import DataParallelModel

    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    dp = DataParallelModel(m, 1)
    assert all(p.grad is None for p in m.parameters())
    x = torch.randn(4, 5)
    dp(x).square().sum().backward()
    ref = torch.cat([p.grad.reshape(-1).clone() for p in m.parameters()])
    dp.all_reduce_flat()                                   # world 1: only adopts
    assert torch.allclose(dp.flat_grad, ref)
    for p in m.parameters():
        _, off, n = p._mmpl_grad_slot
        assert p.grad.data_ptr() == dp.flat_grad.data_ptr() + 4 * off
    dp(x).square().sum().backward()                        # accumulates in place, inside the flat buffer
    assert torch.allclose(dp.flat_grad, 2 * ref)
    dp.zero_grad()
    assert dp.flat_grad.abs().max().item() == 0 and all(p.grad is None for p in m.parameters())


def test_tile_grid_matches_oracle():
    from multimodal_pl_b200.evaluate import _get_gaussian, tile_origins

    for vol, tile in [((300, 512, 512), (64, 192, 192)), ((19, 37, 41), (8, 16, 16)), ((64, 192, 192), (64, 192, 192)),
                      ((70, 200, 193), (64, 192, 192))]:
        assert tile_origins((1, 1) + vol, tile) == O.tile_grid(vol, tile)
    assert len(tile_origins((1, 1, 300, 512, 512), (64, 192, 192))) == 96
    assert np.array_equal(_get_gaussian((8, 16, 16)), O.gaussian_importance((8, 16, 16)))
    # round-robin sharding over 8 ranks covers every tile exactly once, 12 each
    tiles = list(range(96))
    owned = [[t for t in tiles if t % 8 == r] for r in range(8)]
    assert sorted(sum(owned, [])) == tiles and all(len(o) == 12 for o in owned)


def test_supervise_mask_adapter(tmp_path):
    from multimodal_pl_b200.supervise_mask import cmask_lut, read_supervise_mask, remap_unsupervised

    p = tmp_path / "m.csv"
    p.write_text('name,mask\namos_0001.nii.gz,"[0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]"\n'
                 'amos_0507.nii.gz,"[0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]"\n')
    t = read_supervise_mask(str(p))
    assert t == O.read_supervise_mask(str(p))
    lab = torch.randint(0, 16, (2, 1, 3, 4, 5)).float()
    for key in t:
        assert torch.equal(remap_unsupervised(lab, t[key]), O.remap_unsupervised(lab, t[key]))
        assert cmask_lut(t[key])[0] == 0 and len(cmask_lut(t[key])) == 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_pl_b200.engine import DataParallelModel

    torch.manual_seed(100 + rank)                       # different init per rank: the wrapper must broadcast rank 0's
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    dp = DataParallelModel(model, world, bucket_mb=0.0002)        # tiny buckets -> several collectives
    assert len(dp._buckets) >= 2
    g = torch.Generator().manual_seed(7)
    xs, ys = torch.randn(world, 5, 8, generator=g), torch.randn(world, 5, 4, generator=g)
    for step in range(2):
        dp.zero_grad()
        ((dp(xs[rank]) - ys[rank]) ** 2).mean().backward()
    torch.save({"flat": dp.flat_grad.clone(), "w0": model[0].weight.detach().clone()}, os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo(tmp_path):
    """2-process gloo: bucketed, overlapped all-reduce == gradient of the mean loss over the concatenated batch."""
    world, port = 2, _free_port()
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    assert torch.equal(r[0]["w0"], r[1]["w0"])                          # parameters were broadcast
    assert torch.allclose(r[0]["flat"], r[1]["flat"], atol=1e-7)        # same averaged gradient everywhere
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    g = torch.Generator().manual_seed(7)
    xs, ys = torch.randn(world, 5, 8, generator=g), torch.randn(world, 5, 4, generator=g)
    loss = sum(((ref(xs[i]) - ys[i]) ** 2).mean() for i in range(world)) / world
    loss.backward()
    flat_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    assert torch.allclose(r[0]["flat"], flat_ref, atol=1e-6)


def test_engine_surface():
    import argparse
    import sys

    from multimodal_pl_b200.engine import Engine

    argv, sys.argv = sys.argv, ["prog"]
    try:
        with Engine(custom_parser=argparse.ArgumentParser()) as e:
            assert e.world_size == 1 and e.local_rank == 0 and not e.distributed
            assert abs(e.all_reduce_tensor(torch.tensor([1.0, 3.0])).item() - 2.0) < 1e-7
            for name in ("data_parallel", "get_train_loader", "get_test_loader", "all_reduce_tensor"):
                assert callable(getattr(e, name))
    finally:
        sys.argv = argv


@pytest.mark.parametrize("volume,tile,world", [((300, 512, 512), (64, 192, 192), 8), ((300, 512, 512), (64, 192, 192), 4),
                                               ((40, 72, 88), (16, 32, 32), 2), ((41, 72, 88), (16, 32, 32), 3),
                                               ((64, 192, 192), (64, 192, 192), 8)])
def test_sliding_window_plane_exchange_plan_sums_every_contribution_once(volume, tile, world):
    """The sharded sliding window exchanges only the accumulator planes a rank touched (evaluate._exchange_plan).  Simulated
    on the host with one number per (rank, plane): after the exchange the owner of every plane holds the sum over all
    ranks that touched it, every send has a matching receive of the same extent, and far fewer planes move than a
    reduce-scatter of the whole accumulator would move."""
    from multimodal_pl_b200 import evaluate as E

    D = volume[0]
    tiles = E.tile_origins((1, 1) + tuple(volume), tile)
    runs = [E._my_tiles(tiles, r, world) for r in range(world)]
    assert sorted(t for m in runs for t in m) == sorted(tiles)
    ranges = [E._depth_range(m, tile[0]) for m in runs]
    dpad = (D + world - 1) // world * world
    slab = dpad // world
    rng = np.random.RandomState(0)
    acc = np.zeros((world, dpad))
    for r, (lo, hi) in enumerate(ranges):
        acc[r, lo:hi] = rng.rand(hi - lo) + 1.0
    plans = [E._exchange_plan(ranges, r, slab, world) for r in range(world)]
    out = acc.copy()
    moved = 0
    for r in range(world):
        for peer, lo, hi in plans[r][1]:
            assert (r, lo, hi) in plans[peer][0], "a receive without the matching send"
            assert r * slab <= lo < hi <= (r + 1) * slab
            out[r, lo:hi] += acc[peer, lo:hi]
            moved += hi - lo
        for peer, lo, hi in plans[r][0]:
            assert (r, lo, hi) in plans[peer][1], "a send without the matching receive"
    total = acc.sum(0)
    for r in range(world):
        np.testing.assert_allclose(out[r, r * slab:(r + 1) * slab], total[r * slab:(r + 1) * slab], rtol=1e-12)
    if world >= 4 and len(tiles) >= 4 * world:
        assert moved < 0.5 * (world - 1) * dpad, (moved, (world - 1) * dpad)


def test_sliding_window_tile_batch_plan():
    """engine.plan_tile_batches: tiles per forward and streams for a rank's run of tiles.  Every tile is covered exactly
    once by full batches plus one remainder batch, a batch never exceeds 8 tiles unless asked for, the automatic choice
    gives every stream the same number of batches, and there are never more streams than batches."""
    from multimodal_pl_b200 import evaluate as E
    from multimodal_pl_b200.engine import plan_tile_batches

    assert plan_tile_batches(96) == (8, 2)            # one GPU, 300x512x512 / 64x192x192
    assert plan_tile_batches(48) == (8, 2)            # 2 GPUs
    assert plan_tile_batches(24) == (6, 2)            # 4 GPUs: 4 x 6 instead of 3 x 8 (two batches per stream)
    assert plan_tile_batches(12) == (6, 2)            # 8 GPUs: one batch per stream instead of 8 + 4
    assert plan_tile_batches(36) == (6, 2)
    assert plan_tile_batches(1) == (1, 1)
    assert plan_tile_batches(5, lanes=2) == (3, 2)    # 3 + 2
    assert plan_tile_batches(36, lanes=1, tile_batch=8) == (8, 1)
    assert plan_tile_batches(3, lanes=2, tile_batch=8) == (3, 1)     # clamped to the run: one batch, one stream
    assert plan_tile_batches(96, lanes=3) == (8, 3)
    for world in (1, 2, 3, 4, 8):
        run = (len(E.tile_origins((1, 1, 300, 512, 512), (64, 192, 192))) + world - 1) // world
        for lanes in (1, 2, 3):
            tb, ln = plan_tile_batches(run, lanes)
            nb = (run + tb - 1) // tb
            assert 1 <= tb <= 8 and 1 <= ln <= min(lanes, nb)
            assert (run // tb) * tb + run % tb == run and run % tb < tb
