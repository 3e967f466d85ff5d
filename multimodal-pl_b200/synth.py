"""Synthetic inputs of the measured workloads, product side: what bench.py and the examples feed the
hot path.  (The test oracle has its own copy of these generators; the product never imports ``oracle/``.)

CT patch: clip(N(0, 0.5), -1, 1) -- mimics the +-325 HU window of MOTSDataset.py:171-183; MRI patch: z-scored N(0, 1)
(:184-185).  Labels: nearest-seed Voronoi blobs, about half background, every class present; returned as float class ids
[B,1,D,H,W] like the reference's label tensors (train_amos_atlas_final.py:214) or as uint8.
"""
import numpy as np
import torch


def synth_patch(shape, seed: int, modality: str = "ct") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(tuple(shape), generator=g)
    return (0.5 * x).clamp_(-1, 1) if modality == "ct" else x


def synth_labels(shape, seed: int, num_classes: int = 16, n_seeds: int = 48, dtype=torch.float32) -> torch.Tensor:
    """``shape`` = (B, D, H, W) -> [B,1,D,H,W]."""
    B, D, H, W = shape
    rng = np.random.RandomState(seed)
    out = np.zeros((B, D, H, W), dtype=np.float32)
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    for b in range(B):
        pts = rng.rand(n_seeds, 3) * np.array([D, H, W])
        cls = np.where(np.arange(n_seeds) % 2 == 0, 0, (np.arange(n_seeds) // 2) % (num_classes - 1) + 1)
        best = np.full((D, H, W), np.inf)
        lab = np.zeros((D, H, W), dtype=np.float32)
        for (pz, py, px), c in zip(pts, cls):
            d2 = ((zz - pz) * 3.0) ** 2 + (yy - py) ** 2 + (xx - px) ** 2
            m = d2 < best
            best[m] = d2[m]
            lab[m] = c
        out[b] = lab
    return torch.from_numpy(out).unsqueeze(1).to(dtype)


def synth_labels_upsampled(batch: int, dhw, seed: int, num_classes: int = 16, n_seeds: int = 32, factor: int = 4,
                           dtype=torch.float32) -> torch.Tensor:
    """Blobs generated at 1/factor resolution and nearest-neighbour up-sampled to ``dhw`` (cheap for full-size patches)."""
    lo = synth_labels((batch,) + tuple(max(1, s // factor) for s in dhw), seed, num_classes, n_seeds)
    return torch.nn.functional.interpolate(lo, size=tuple(dhw), mode="nearest").contiguous().to(dtype)
