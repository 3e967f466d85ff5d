#!/usr/bin/env python
"""bench.py -- throughput of the multimodal-PL hot path on B200: 3-D patches/s of the unet3D train step (and of
sliding-window inference).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg5|cfg4]

A "step" is one pass of the hot path over one batch of synthetic patches: unet3D_baseline forward, partial-label loss,
backward, (N>1: bucketed NCCL gradient all-reduce captured inside the step's CUDA graph), fused SGD.  Default workload =
BASELINE.json configs[1] ("cfg2"): bf16, batch 2 per GPU, 1x64x192x192 patches, 16 classes, random init, one CT and one
MRI sample per GPU with per-sample partial-label masks.  Inputs are 2 x 2.36 M voxels per step and the activations touched
per step (> 4 GB) exceed the 126 MB L2, so no explicit L2 flush is needed (stated in config).

Rank 0 prints ONE JSON line (see README / the task contract):
  value            patches/s over all ranks, inputs resident in HBM, CUDA events, MAX over ranks
  e2e              the same metric through the public API from PINNED HOST buffers (H2D of every step's batch on a copy
                   stream, loss read back D2H every step)
  roofline         the dominant kernel (tcgen05 conv) against the measured bf16 peak -- the BURST peak when the clocks
                   sampled during the timed region show the chip was not power-throttled (>= 1.8 GHz), else the sustained
                   one; `all_tcgen05_convs` is the north-star figure (all conv launches together)
  infer            BASELINE configs[3] in the same record: sliding-window inference of a 300x512x512 volume (96 tiles)
  cpu_baseline     the reference's own modules (baseline/_ref; else the oracle port) on this box's host cores, bounded sample
  gpu_library_baseline  the reference's own modules on THIS GPU through eager PyTorch (cuDNN / ATen, bf16 autocast +
                   channels_last_3d) -- the library path the hand-written kernels replace
  dp_parity / sw_parity (N > 1)  DP step == single process, sharded sliding window == single rank, checked before timing

--impl reference times the reference's own CPU implementation of the path on all host threads: the unmodified
reference modules when baseline/_ref is staged (kind "reference"), else the oracle port (kind "port"); each step is a
bounded sample (ONE 64x192x192 patch of the batch of two: fwd + loss + bwd + SGD) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (batch per GPU, patch D,H,W, base width, classes)
    "cfg1": (1, (64, 128, 128), 32, 16),
    "cfg2": (2, (64, 192, 192), 32, 16),
    "cfg5": (4, (96, 224, 224), 64, 16),
}
# algorithmic conv FLOPs per patch (BASELINE.md section 3 / SURVEY 8d)
TRAIN_TFLOP_PER_PATCH = {"cfg1": 1.3662, "cfg2": 3.0740, "cfg5": 25.041}
FWD_TFLOP_PER_TILE = 1.0260          # one 64x192x192 tile, forward
CFG4 = {"volume": (1, 1, 300, 512, 512), "tile": (64, 192, 192), "base": 32, "classes": 16}
# one supervised organ (CT rows of supervise_mask.csv) / background only (MRI rows), w16 = [w_bg] + csv_row(15)
W16_CT = [1.0, 0, 0, 0, 1.0] + [0.0] * 11
W16_MRI = [1.0] + [0.0] * 15


def train_config(workload, world):
    """Identical for both arms (ours and --impl reference): the workload, not how it is executed (see `launch`)."""
    batch, dhw, base, classes = WORKLOADS[workload]
    return {"workload": workload, "patch": list(dhw), "batch_per_gpu": batch, "base": base, "classes": classes,
            "parallelism": f"dp{world}", "l2": "inputs+activations per step >> 126 MB L2, no flush needed",
            "optimizer": "SGD momentum 0.9 wd 1e-4", "loss": "EDiceLoss_partial (partial-label Dice + BCE)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "power_w": pw[len(pw) // 2] if pw else None}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def pick_peak(peaks, clocks):
    """Burst peak for a kernel timed while the chip runs at full clocks, sustained peak when the timed region itself
    was power-throttled (B200_PROFILING.md); the fallbacks are the recipe's."""
    sm = (clocks or {}).get("sm_mhz")
    burst = peaks.get("bf16_tflops") or 1670.0
    sustained = peaks.get("bf16_tflops_sustained") or 1400.0
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    if sm is not None and sm >= 1800.0:
        return burst, f"{src} bf16_tflops (burst): median SM clock {sm:.0f} MHz during the timed region, not power-throttled"
    return sustained, f"{src} bf16_tflops_sustained: median SM clock {sm} MHz during the timed region"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu capture of
    THIS workload (profiles/r02_ncu_dominant.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_dominant.json")) as f:
            rec = json.load(f).get(workload)
        return (rec or {}).get("traffic_bytes_per_launch"), (rec or {}).get("source")
    except (OSError, ValueError):
        return None, None


# ------------------------------------------------------------------------------------------------------------------
# reference / CPU legs (the only code here that touches oracle/ or baseline/_ref)
# ------------------------------------------------------------------------------------------------------------------
class ReferenceStep:
    """One train step of the reference path on ONE 64x192x192 patch (B = 1): forward, EDiceLoss_partial, backward,
    torch.optim.SGD -- through the unmodified reference modules when they are staged (baseline/_ref), else through the
    oracle port (same ATen ops in the same order).  Runs on ``device`` (cpu for the baseline, cuda for the library leg)."""

    def __init__(self, workload, device="cpu", autocast=False):
        import torch

        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import _refload
        import mmpl_oracle as O

        self.torch, self.O = torch, O
        batch, dhw, base, classes = WORKLOADS[workload]
        self.dhw, self.base, self.classes, self.device, self.autocast = dhw, base, classes, device, autocast
        sd = O.synth_state_dict(base, classes, 0)
        self.kind = "port"
        self.model = None
        root = _refload.reference_root() if base == 32 else None        # the reference class is hard-wired to base 32
        if root is not None:
            ref_unet, ref_lp, _ = _refload.load_reference(root, device_type="cuda" if device != "cpu" else "cpu",
                                                          with_eval=False)
            self.model = ref_unet.unet3D_baseline([1, 2, 2, 2, 2], num_classes=classes, weight_std=True)
            self.model.load_state_dict(sd)
            self.model.to(device).train()
            self.crit = ref_lp.EDiceLoss_partial(classes)
            self.crit.device = device
            params = list(self.model.parameters())
            self.kind = "reference"
        else:
            self.sd = {k: v.to(device).requires_grad_(True) for k, v in sd.items()}
            params = list(self.sd.values())
        self.opt = torch.optim.SGD(params, lr=1e-2, momentum=0.9, weight_decay=1e-4)
        self.x = O.synth_patch((1, 1) + tuple(dhw), 100, "ct").to(device)
        lab = torch.nn.functional.interpolate(O.synth_labels((1,) + tuple(max(1, s // 4) for s in dhw), 200, classes, 32),
                                              size=tuple(dhw), mode="nearest")
        self.lab = O.remap_unsupervised(lab, W16_CT).to(device)
        self.w = [torch.tensor(W16_CT, device=device)]
        if autocast:
            self.x = self.x.contiguous(memory_format=torch.channels_last_3d)

    def __call__(self):
        torch = self.torch
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            if self.model is not None:
                logits = self.model(self.x, self.lab)[0]
            else:
                logits = self.O.unet3d_forward(self.sd, self.x, self.base)
        if self.model is not None:
            loss = self.crit(logits.float(), self.lab.squeeze(1), mask=self.w, soft_max=True)
        else:
            loss = self.O.partial_label_loss(logits.float(), self.lab.squeeze(1), W16_CT)
        loss.backward()
        self.opt.step()
        return loss

    def sample_text(self):
        d = self.dhw
        what = "unmodified reference modules (baseline/_ref)" if self.kind == "reference" else "oracle port (same ATen ops/order)"
        return f"one {d[0]}x{d[1]}x{d[2]} patch (B=1) fwd + EDiceLoss_partial + bwd + SGD per step, fp32, {what}"


def run_reference(args):
    """Reference arm: the path's own CPU implementation on all host threads, same config as our arm; each step a
    bounded sample (one patch of the batch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    if args.workload == "cfg4":
        emit({"impl": "reference", "unavailable": "the reference arm times the train step; cfg4 is reported by our arm's `infer` record"})
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step = ReferenceStep(args.workload, "cpu")
    for _ in range(min(args.warmup, 2)):          # CPU: two warm-up steps page everything in
        step()
    times = []
    budget = 240.0                                # the whole run ends within a few minutes whatever K is
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget:
            break
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": "3D patches/sec (train fwd+bwd)", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(args.workload, max(args.gpus, 1)),
        "launch": "eager PyTorch CPU (oneDNN), all host threads, rank 0 only",
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": threads, "kind": step.kind,
                         "sample": step.sample_text() + (f"; stopped after {len(times)} of {args.steps} steps (time budget)"
                                                         if len(times) < args.steps else "")},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def _e2e_probe(step, image_h, label_h, steps):
    """Diagnostic (MMPL_BENCH_E2E_PROBE=1): where does a step of the host-fed path spend its time?  Prints to stderr: the
    H2D time of one batch alone, then three repetitions of the e2e loop with, per step, the host time of each call and
    the device time between consecutive step-end events."""
    import time

    import torch

    cs = step._copy_stream
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    for k in range(3):
        a.record(cs)
        step.stage(image_h, label_h)
        b.record(cs)
        torch.cuda.synchronize()
        step._stage_free.record()
        sys.stderr.write(f"[e2e probe] H2D of one batch alone: {a.elapsed_time(b):.3f} ms\n")
    loss_pin = torch.empty(2, dtype=torch.float32).pin_memory()
    for rep in range(3):
        loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        host = []
        torch.cuda.synchronize()
        ends[0].record()
        t_begin = time.perf_counter()
        step.stage(image_h, label_h)
        for i in range(steps):
            t0 = time.perf_counter()
            out = step.run_staged()
            t1 = time.perf_counter()
            loss_pin[i % 2:i % 2 + 1].copy_(out.detach().reshape(1), non_blocking=True)
            loss_ev[i % 2].record()
            ends[i + 1].record()
            t2 = time.perf_counter()
            if i + 1 < steps:
                step.stage(image_h, label_h)
            t3 = time.perf_counter()
            if i > 0:
                loss_ev[(i - 1) % 2].synchronize()
            t4 = time.perf_counter()
            host.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
        torch.cuda.synchronize()
        total = (time.perf_counter() - t_begin) * 1e3
        dev = [ends[i].elapsed_time(ends[i + 1]) for i in range(steps)]
        sys.stderr.write(f"[e2e probe] rep {rep}: {total / steps:.3f} ms/step wall; device ms between step ends: "
                         + " ".join(f"{d:.2f}" for d in dev) + "\n")
        sys.stderr.write("[e2e probe]   host ms (run_staged, loss copy+record, stage, wait prev): "
                         + " | ".join(" ".join(f"{x * 1e3:.2f}" for x in h) for h in host) + "\n")


def cpu_baseline(workload, seconds=14.0):
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step = ReferenceStep(workload, "cpu")
    step()
    n, tsum = 0, 0.0
    while tsum < seconds and n < 6:
        t0 = time.perf_counter()
        step()
        tsum += time.perf_counter() - t0
        n += 1
    return {"value": n / tsum, "unit": "patches/s", "cores": threads, "kind": step.kind,
            "sample": f"{n} x " + step.sample_text()}


def gpu_library_baseline(workload, dev):
    """The reference's modules on this GPU through eager PyTorch (cuDNN conv3d + ATen), bf16 autocast +
    channels_last_3d and plain fp32 (TF32 off): what a user gets from the reference as shipped on a B200."""
    import torch

    out = {}
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for name, ac in (("bf16_autocast_channels_last", True), ("fp32", False)):
        try:
            step = ReferenceStep(workload, dev, autocast=ac)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            a.record()
            for _ in range(n):
                step()
            b.record()
            torch.cuda.synchronize()
            out[name] = {"value": n / (a.elapsed_time(b) / 1e3), "unit": "patches/s", "ms_per_patch": a.elapsed_time(b) / n,
                         "kind": step.kind, "sample": "B=1 " + step.sample_text()}
            del step
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    return out


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    return rank, local_rank, world, dev


def dp_parity(world, rank, dev):
    """N > 1: the NCCL data-parallel step must equal a single process.  fp32 exact kernels, tiny patches: every rank
    computes the averaged gradient through DataParallelModel (bucketed all-reduce) and, on a second copy of the model,
    the mean of the per-rank losses serially.  Returns the max relative L2 error over the flat gradient (tolerance 1e-4)."""
    import torch
    import torch.distributed as dist

    import multimodal_pl_b200 as mm
    from multimodal_pl_b200 import synth
    from multimodal_pl_b200.engine import DataParallelModel
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.unet3D import unet3D_baseline

    mm.set_compute_dtype(torch.float32)
    mm.set_conv_algo("direct")
    try:
        torch.manual_seed(1)
        model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).to(dev)
        ref = unet3D_baseline([1, 2, 2, 2, 2], num_classes=16, weight_std=True).to(dev)
        dp = DataParallelModel(model, world, bucket_mb=2)            # broadcasts rank 0's weights
        ref.load_state_dict(model.state_dict())
        crit = EDiceLoss_partial(16)
        xs = [synth.synth_patch((1, 1, 16, 16, 32), 50 + r, "ct").to(dev) for r in range(world)]
        ls = [synth.synth_labels((1, 16, 16, 32), 60 + r, 16, 32).to(dev) for r in range(world)]
        w = [torch.ones(16)]
        dp.zero_grad()
        crit(dp(xs[rank])[0], ls[rank].squeeze(1), mask=w).backward()
        (sum(crit(ref(xs[r])[0], ls[r].squeeze(1), mask=w) for r in range(world)) / world).backward()
        flat_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        err = ((dp.flat_grad - flat_ref).norm() / flat_ref.norm()).reshape(1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        return float(err.item())
    finally:
        mm.set_conv_algo("auto")
        mm.set_compute_dtype(torch.bfloat16)


def run_train(args):
    import torch
    import torch.distributed as dist

    import multimodal_pl_b200 as mm
    from multimodal_pl_b200 import _lib, ops, synth
    from multimodal_pl_b200.engine import DataParallelModel, FusedSGD, GraphedTrainStep
    from multimodal_pl_b200.loss_functions.loss_partial import EDiceLoss_partial
    from multimodal_pl_b200.supervise_mask import cmask_lut
    from multimodal_pl_b200.unet3D import unet3D_baseline

    rank, local_rank, world, dev = _dist_setup()
    _lib.require_device()
    mm.set_compute_dtype(torch.bfloat16)
    parity = None
    if world > 1:
        parity = dp_parity(world, rank, dev)
        assert parity < 1e-4, f"data-parallel step differs from the single-process step: rel-L2 {parity:.3e}"

    batch, dhw, base, classes = WORKLOADS[args.workload]
    torch.manual_seed(0)
    model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=classes, weight_std=True, base=base).to(dev)
    model.train()
    dp = DataParallelModel(model, world, bucket_mb=args.bucket_mb, average=False)
    opt = FusedSGD(dp.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4, flat_grad=dp.flat_grad)
    crit = EDiceLoss_partial(classes)

    # synthetic batch: CT / MRI samples alternate, Voronoi labels (uint8), one supervised organ (CT) or background only
    # (MRI); the class weights and the cmask remap are PER SAMPLE, different seeds per rank
    imgs, labs, wts = [], [], []
    for b in range(batch):
        ct = b % 2 == 0
        imgs.append(synth.synth_patch((1, 1) + dhw, 100 + rank * 16 + b, "ct" if ct else "mri"))
        labs.append(synth.synth_labels_upsampled(1, dhw, 200 + rank * 16 + b, classes, 32, dtype=torch.uint8))
        wts.append(torch.tensor((W16_CT if ct else W16_MRI)[:classes]))
    image_h = torch.cat(imgs).pin_memory()
    label_h = torch.cat(labs).pin_memory()                                   # uint8 [B,1,D,H,W]
    luts = torch.stack([cmask_lut(w.tolist()) for w in wts]).to(dev)
    image_d, label_d = image_h.to(dev), label_h.to(dev)

    def loss_fn(logits, lab):
        return crit(logits, lab.squeeze(1), mask=wts, soft_max=True, lut=luts, per_sample=True)

    # default: the head of the step is unet3D_baseline.forward_partial_loss -- classifier + EDiceLoss_partial in one
    # forward and one backward launch, no logits tensor (csrc/cls_loss.cu); --two-step-head runs model(x) -> crit(logits)
    fused_head = not args.two_step_head

    def fused_loss(module, img, lab):
        return module.forward_partial_loss(img, lab, wts, lut=luts, per_sample=True)

    def eager_step(img, lab):
        opt.zero_grad()
        if fused_head:
            loss = fused_loss(dp.module, img, lab)
        else:
            logits, _, _ = dp(img, lab)
            loss = loss_fn(logits, lab)
        loss.backward()
        opt.step(grad_scale=1.0 / world)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.eager
    warm = max(args.warmup, 3)
    if use_graph:
        # the public API for a launch-overhead-free step: capture once, replay per batch (engine.GraphedTrainStep)
        step = GraphedTrainStep(dp, loss_fn, opt, image_d, label_d, warmup=warm,
                                fused_loss=fused_loss if fused_head else None)
    else:
        step = eager_step
    for _ in range(warm):
        step(image_d, label_d)
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(image_d, label_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    time.sleep(0.05)          # let the poller's teardown finish before the host-fed loop starts

    # ---- timed region 2: end to end through the public API with pinned-host inputs ------------------------------------
    # every step: H2D of that step's batch (copy stream, overlapping the previous step's compute) and D2H of its loss
    if use_graph:
        # warm-up of THIS path too (copy stream, staging buffers, first transfers out of the pinned batch): untimed
        for _ in range(warm):
            step.stage(image_h, label_h)
            step.run_staged()
    barrier()
    # no nvidia-smi polling in THIS region: the host feeds it step by step, and a 20 ms NVML poll holds driver locks long
    # enough to stall that loop (measured on B200: the identical loop 11.8-11.9 ms/step alone, 12.0-16.6 ms/step with the
    # sampler running; region 1 is immune because all its replays are queued within the first millisecond)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the pinned landing buffer of the losses exists before the clock starts: cudaHostAlloc is a millisecond-scale,
    # device-synchronising call (it used to sit inside the region and cost ~12 ms once, i.e. 1.2 ms/step at 10 steps)
    loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    torch.cuda.synchronize()
    f0.record()
    last = 0.0
    if use_graph:
        # the loss of every step is copied to pinned host memory right behind its step and READ one step later, like a
        # training loop that logs asynchronously: the host stays one step ahead, every step's result reaches the host inside
        # the timed region (the last one before the closing event)
        step.stage(image_h, label_h)
        for i in range(args.steps):
            out = step.run_staged()
            loss_pin[i % 2:i % 2 + 1].copy_(out.detach().reshape(1), non_blocking=True)
            loss_ev[i % 2].record()
            if i + 1 < args.steps:
                step.stage(image_h, label_h)           # next batch's PCIe transfer runs under this step's kernels
            if i > 0:
                loss_ev[(i - 1) % 2].synchronize()
                last = float(loss_pin[(i - 1) % 2])
        loss_ev[(args.steps - 1) % 2].synchronize()
        last = float(loss_pin[(args.steps - 1) % 2])
    else:
        for _ in range(args.steps):
            last = step(image_h.to(dev, non_blocking=True), label_h.to(dev, non_blocking=True)).item()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if os.environ.get("MMPL_BENCH_E2E_PROBE") and use_graph and rank == 0:
        _e2e_probe(step, image_h, label_h, args.steps)

    # ---- roofline region: the same step run eagerly with every tcgen05 conv launch bracketed by CUDA events on the
    # launching stream (a graph replay cannot be bracketed per kernel); identical kernels, shapes and data
    dp.sync_in_backward = True
    ops.enable_conv_profile(True)
    launches_eager0 = _lib.launch_count()
    prof_steps = 3
    for _ in range(prof_steps):
        # let the host run ahead of the device (about 30 ms of spin on the stream) so the bracketed launches execute
        # back to back and an event pair measures the kernel, not the Python launch latency in front of it
        torch.cuda._sleep(int(0.030 * 1.9e9))
        eager_step(image_d, label_d)
    launches_per_step = (_lib.launch_count() - launches_eager0) // prof_steps
    prof = ops.collect_conv_profile()
    if args.conv_table and rank == 0:
        sys.stderr.write("tcgen05 conv launches per kernel key (op, Cmin, Cmax, k, stride, voxels): launches/step, ms/step, TFLOP/s\n")
        for k, v in sorted(prof["per_key"].items(), key=lambda kv: -kv[1]["ms"]):
            sys.stderr.write(f"  {k:52s} {v['launches'] // prof_steps:3d} {v['ms'] / prof_steps:8.3f} "
                             f"{v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f}\n")
    ops.enable_conv_profile(False)
    if use_graph:
        launches = launches_per_step * args.steps      # kernels executed by the replayed graphs in the timed region
    barrier()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # the slowest rank defines the step time
    ms, ms_e2e = t.tolist()

    # ---- BASELINE configs[3] in the same record: sliding-window inference (all ranks take part) ----------------------
    infer = None
    comm_in_graph = bool(use_graph and getattr(step, "comm_in_graph", False))
    if use_graph:
        step.close()        # the graph pins NCCL kernels: release it before the process group is destroyed
    del step
    torch.cuda.empty_cache()
    if args.workload == "cfg2" and not args.no_infer:
        try:
            # a freshly initialised network (seed 0, broadcast from rank 0): 30 SGD steps on noise labels would leave one that
            # predicts background everywhere
            del model, dp, opt
            torch.cuda.empty_cache()
            infer = infer_record(None, dev, rank, world, steps=3, warmup=2, eager=args.eager)
        except Exception as e:  # noqa: BLE001
            infer = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    patches = batch * world * args.steps
    value = patches / (ms / 1e3)
    peaks = load_peaks()
    peak_tf, peak_src = pick_peak(peaks, clocks)
    # dominant kernel = the tcgen05 conv instantiation with the largest total time in the timed region
    # (cfg2: conv_tc_kernel<32,32,4> on the full-resolution 32->32 3x3x3 layers, fprop + dgrad launches)
    traffic, traffic_src = ncu_traffic(args.workload)
    dom = prof["dominant"]
    dom_ms, dom_flops, dom_n = dom["ms"], dom["flops"], max(dom["launches"], 1)
    ms_step = ms / args.steps
    achieved = dom_flops / (dom_ms / 1e3) / 1e12 if dom_ms > 0 else 0.0
    tc_ms, tc_flops = prof["ms"], prof["flops"]
    tc_tf = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    k3_tf = prof["k3"]["flops"] / (prof["k3"]["ms"] / 1e3) / 1e12 if prof["k3"]["ms"] > 0 else 0.0
    dom_name = "wgrad_tc_kernel (tcgen05 weight gradient) " if dom["key"] and dom["key"][0] == "wgrad_tc" else \
        "conv_tc_kernel (tcgen05 implicit-GEMM conv: fprop, fprop+residual, dgrad, dgrad+GN-backward launches) "
    dom_vox = dom["key"][5] if dom["key"] else 0
    dom_c = dom["key"][2] if dom["key"] else 0
    roofline = {"bound": "tensor", "kernel": dom_name + str(dom["key"]),
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": 2 * dom_vox * dom_c * 2,
                "peak_source": peak_src, "launches": dom["launches"],
                "flops_per_launch": dom_flops / dom_n, "avg_launch_ms": dom_ms / dom_n,
                "share_of_step": (dom_ms / prof_steps) / ms_step,
                "measured": f"{prof_steps} eager steps right after the timed region (host queued ahead of the device), one CUDA-event "
                            "pair per launch; in this pass the weight-gradient chain stays on the main stream, so the shares "
                            "are of a serialised step (the timed step overlaps wgrad with the critical path)",
                "all_tcgen05_convs": {"achieved": tc_tf, "frac": tc_tf / peak_tf, "unit": "TFLOP/s",
                                      "launches": prof["launches"], "ms_per_step": tc_ms / prof_steps,
                                      "share_of_step": (tc_ms / prof_steps) / ms_step,
                                      "note": "north_star target: >= 0.5 of the dense bf16 roofline over all conv launches"},
                "tensor_bound_convs": {"achieved": k3_tf, "frac": k3_tf / peak_tf, "unit": "TFLOP/s",
                                       "launches": prof["k3"]["launches"], "ms_per_step": prof["k3"]["ms"] / prof_steps,
                                       "note": "the 3x3x3 convolutions with >= 32 input channels only: the Cin=1 stem and the "
                                               "1x1x1 convolutions are HBM / issue bound (SURVEY 8d) and dilute the aggregate above"},
                "whole_step": {"achieved": TRAIN_TFLOP_PER_PATCH[args.workload] * value / world,
                               "frac": TRAIN_TFLOP_PER_PATCH[args.workload] * value / world / peak_tf, "unit": "TFLOP/s per GPU",
                               "note": "algorithmic conv FLOPs of the step / step time (includes every non-conv kernel)"}}
    cpu = gpu_lib = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            gpu_lib = gpu_library_baseline(args.workload, dev) if args.workload != "cfg5" else None
        except Exception as e:  # noqa: BLE001
            gpu_lib = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        cpu = cpu_baseline(args.workload) if args.workload != "cfg5" else None
    launch = "eager"
    if use_graph:
        launch = "cuda graph replay (engine.GraphedTrainStep"
        if world > 1:
            launch += (", NCCL bucket all-reduces + per-bucket SGD captured in the graph" if comm_in_graph
                       else ", one NCCL all-reduce of the flat gradient + SGD after the replay")
        launch += ")"
    line = {
        "metric": "3D patches/sec (train fwd+bwd)", "value": value, "unit": "patches/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": train_config(args.workload, world), "launch": launch,
        "batch": "per GPU: CT and MRI samples alternate, uint8 Voronoi labels, per-sample class weights + cmask LUT",
        "head": ("unet3D_baseline.forward_partial_loss: classifier + EDiceLoss_partial fused, no logits tensor (mmpl_cls_loss_fwd/_bwd)"
                 if fused_head else "model(x) -> EDiceLoss_partial(logits) (mmpl_cls_fwd, mmpl_partial_loss_fwd/_bwd, mmpl_cls_bwd)"),
        "clocks": clocks,
        "e2e": {"value": patches / (ms_e2e / 1e3), "unit": "patches/s",
                "h2d_bytes_per_step": image_h.numel() * image_h.element_size() + label_h.numel() * label_h.element_size(),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "last_loss": last,
                "how": "pinned host batch -> staging buffers on a copy stream (overlaps the previous step) -> D2D into the "
                       "graph's static inputs -> replay -> loss copied to pinned host memory behind every step and read one step later "
                       "(the host runs one step ahead); fp32 image + uint8 labels"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "infer": infer,
        "cpu_baseline": cpu,
        "gpu_library_baseline": gpu_lib,
    }
    if parity is not None:
        line["dp_parity"] = {"rel_l2": parity, "tolerance": 1e-4,
                             "what": "flat gradient of the NCCL data-parallel step vs the mean of the per-rank losses computed "
                                     "in one process, fp32 exact kernels, before the timed region"}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def infer_record(model, dev, rank, world, steps, warmup, eager=False):
    """BASELINE configs[3]: sliding-window inference of one synthetic 300x512x512 CT volume, tile 64x192x192 -> 96 tiles.
    Forward + classifier + Gaussian blending per tile in ONE CUDA graph replayed per tile, argmax + Dice on the device
    (evaluate.predict_sliding_dice); N > 1: contiguous runs of tiles per rank, exchange of the touched fp32 accumulator planes along
    depth, local argmax/Dice per slab, all-gather of the uint8 mask.  A step = one volume; value = tiles (3-D patches) per
    second over the whole job with the volume resident in HBM; e2e = the same call fed from pinned host memory (each rank
    uploads only the depth range it needs) with the uint8 mask and the Dice values read back.  Fixed total work: strong."""
    import torch
    import torch.distributed as dist

    from multimodal_pl_b200 import _lib, synth
    from multimodal_pl_b200.engine import GraphedSlidingWindow
    from multimodal_pl_b200.evaluate import predict_sliding_dice, tile_origins
    from multimodal_pl_b200.unet3D import unet3D_baseline

    vol_shape, tile, base, classes = CFG4["volume"], CFG4["tile"], CFG4["base"], CFG4["classes"]
    if model is None:
        torch.manual_seed(0)
        model = unet3D_baseline([1, 2, 2, 2, 2], num_classes=classes, weight_std=True, base=base).to(dev)
        if world > 1:
            for p_ in model.parameters():
                dist.broadcast(p_.data, src=0)
    model.eval()
    g = torch.Generator().manual_seed(7)
    vol_h = torch.randn(vol_shape, generator=g).mul_(0.5).clamp_(-1, 1).pin_memory()
    lab_h = synth.synth_labels_upsampled(1, vol_shape[2:], 11, classes, 48, dtype=torch.uint8).pin_memory()
    vol_d, lab_d = vol_h.to(dev), lab_h.to(dev)
    ntiles = len(tile_origins(vol_shape, tile))
    sw_par = None
    if world > 1:                       # sharded == single rank, on a small volume, before timing
        small = synth.synth_patch((1, 1, 40, 72, 88), 41, "ct").to(dev)
        slab = synth.synth_labels((1, 40, 72, 88), 42, classes, 32, dtype=torch.uint8).to(dev)
        kw = dict(label=slab, acc_dtype=torch.float32, num_class=classes - 1)
        one = predict_sliding_dice(None, [model], small, (16, 32, 32), classes, None, sharded=False, **kw)
        two = predict_sliding_dice(None, [model], small, (16, 32, 32), classes, None, sharded=True, **kw)
        mism = (one[3] != two[3]).sum().reshape(1).float()
        dist.all_reduce(mism, op=dist.ReduceOp.MAX)
        sw_par = {"argmax_mismatches": int(mism.item()), "voxels": int(one[3].numel()),
                  "max_dice_diff": max(abs(float(a) - float(b)) for a, b in zip(one[0], two[0])),
                  "what": "sharded (per-slab plane exchange along depth) vs single-rank sliding window on a 40x72x88 volume; fp32 sums "
                          "of <= 8 tile contributions in a different order, only exact near-ties may flip"}
        assert sw_par["argmax_mismatches"] <= 5 and sw_par["max_dice_diff"] < 1e-4, sw_par
    nets = [model] if eager else [GraphedSlidingWindow(model, vol_shape[2:], tile, classes, world_size=world)]

    def volume(v, l):
        return predict_sliding_dice(None, nets, v, tile, classes, None, label=l, acc_dtype=torch.float32,
                                    sharded=world > 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, warmup)):
        volume(vol_d, lab_d)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    r0 = 0 if eager else nets[0].launches_replayed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dices, _, _, amax = volume(vol_d, lab_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0      # C-ABI calls made directly (finalize, ...) in the timed region
    if not eager:                           # + the kernels executed by this rank's graph replays
        launches += nets[0].launches_replayed - r0
    clocks = sampler.stop() if rank == 0 else None
    # end to end: the volume and its labels start in pinned host memory (each rank uploads the depth range / label slab it
    # needs); the result -- the uint8 segmentation mask and the Dice values -- is read back into host memory on rank 0,
    # the process that would write the prediction (evaluate_amos.py:333-349); the other ranks read back the Dice values
    mask_h = torch.empty((1,) + tuple(vol_shape[2:]), dtype=torch.uint8).pin_memory() if rank == 0 else None
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        dices, _, _, amax = volume(vol_h, lab_h)
        if rank == 0:
            mask_h.copy_(amax, non_blocking=True)
        dice_h = torch.stack([d.reshape(()) for d in dices]).tolist()      # one D2H, synchronises the step
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    peaks = load_peaks()
    peak_tf, peak_src = pick_peak(peaks, clocks)
    value = ntiles * steps / (ms / 1e3)
    rec = {
        "metric": "3D patches/sec (infer, sliding window)", "value": value, "unit": "patches/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg4", "volume": list(vol_shape[2:]), "tile": list(tile), "tiles": ntiles, "base": base,
                   "classes": classes,
                   "parallelism": f"contiguous tile runs over {world} rank(s)" + (", touched accumulator planes sent to their depth-slab owner (grouped NCCL send/recv) + all-gather of the uint8 mask" if world > 1 else ""),
                   "blend": "classifier + Gaussian accumulation fused (mmpl_cls_blend), fp32 depth-major accumulator",
                   "launch": "eager" if eager else f"cuda graphs per batch of {nets[0].tile_batch} tiles, batches alternating between {nets[0].lanes} stream(s); tiles accumulated one launch each in the reference's order (engine.GraphedSlidingWindow)",
                   "l2": "volume + accumulators >> 126 MB L2, no flush needed"},
        "clocks": clocks,
        "e2e": {"value": ntiles * steps / (ms_e2e / 1e3), "unit": "patches/s",
                "h2d_bytes_per_step": vol_h.numel() * 4 + lab_h.numel(),
                "d2h_bytes_per_step": vol_shape[2] * vol_shape[3] * vol_shape[4] + 8 * len(dice_h), "ms_per_step": ms_e2e / steps,
                "mean_dice": sum(dice_h) / max(len(dice_h), 1)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "all conv kernels of the forward pass (whole-volume average)",
                     "achieved": value * FWD_TFLOP_PER_TILE / world, "peak": peak_tf, "unit": "TFLOP/s per GPU",
                     "frac": value * FWD_TFLOP_PER_TILE / peak_tf / world, "traffic": None, "peak_source": peak_src,
                     "frac_of_burst_peak": value * FWD_TFLOP_PER_TILE / (peaks.get("bf16_tflops") or 1670.0) / world},
    }
    if sw_par is not None:
        rec["sw_parity"] = sw_par
    return rec


def run_infer(args):
    """--workload cfg4: the sliding-window record as the top-level line."""
    import torch
    import torch.distributed as dist

    import multimodal_pl_b200 as mm
    from multimodal_pl_b200 import _lib

    rank, local_rank, world, dev = _dist_setup()
    _lib.require_device()
    mm.set_compute_dtype(torch.bfloat16)
    rec = infer_record(None, dev, rank, world, steps=args.steps, warmup=max(1, min(args.warmup, 3)), eager=args.eager)
    if rank == 0:
        rec["vs_baseline"] = None
        rec["cpu_baseline"] = None
        emit(rec)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _quiet_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr; return a file object bound
    to the real stdout so that the ONE JSON line is the only thing printed there."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4"],
                    help="cfg2 (default, train step + infer record), cfg1, cfg5 (train step) or cfg4 (sliding-window inference)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and gpu_library_baseline legs")
    ap.add_argument("--no-infer", action="store_true", help="skip the cfg4 inference record of the default run")
    ap.add_argument("--conv-table", action="store_true", help="print per-kernel-key tcgen05 conv timings to stderr")
    ap.add_argument("--eager", action="store_true", help="drive every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--bucket-mb", type=float, default=8.0, help="gradient bucket size of the in-graph NCCL all-reduces")
    ap.add_argument("--two-step-head", action="store_true",
                    help="model(x) -> EDiceLoss_partial(logits) instead of the fused classifier + loss kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg4":
        run_infer(args)
    else:
        run_train(args)


if __name__ == "__main__":
    main()
