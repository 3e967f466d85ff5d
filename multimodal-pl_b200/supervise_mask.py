"""Adapter for the reference's ``supervise_mask.csv`` label-mask table and the ``cmask`` construction of the train
loop (train_amos_atlas_final.py:177-183, :215-219, :252-255).

The shipped CSV does not match the loop as written: it has a ``name,mask`` header, keys carry
``.nii.gz`` while the loop looks up ``"amos_" + name``, and rows are 15 long (organs 1..15, no background slot) while
the loop indexes ``mask[l]`` for organ ids l >= 1.  This module defines the adapter once:
``w16 = [w_bg] + row15`` (background weight is unspecified by the reference; default 1).
"""
import csv
from typing import Dict, List, Sequence

import torch


def read_supervise_mask(path: str, w_bg: float = 1.0) -> Dict[str, List[float]]:
    table: Dict[str, List[float]] = {}
    with open(path, "r") as f:
        for row in csv.reader(f):
            if len(row) != 2 or row[0] == "name":
                continue
            name, mask = row
            bits = [float(v) for v in mask.strip("[] ").split(",") if v.strip() != ""]
            key = name[:-7] if name.endswith(".nii.gz") else name
            table[key] = [float(w_bg)] + bits
    return table


def cmask_lut(w: Sequence[float]) -> torch.Tensor:
    """16-entry LUT equivalent to ``cmask[cmask == l] = 0 for unsupervised l`` (train:252-255): label l maps to l if
    class l is supervised (or l == 0), else to background."""
    return torch.tensor([float(l) if (l == 0 or w[l]) else 0.0 for l in range(len(w))], dtype=torch.float32)


def remap_unsupervised(labels: torch.Tensor, w: Sequence[float]) -> torch.Tensor:
    """Materialised cmask (what the train loop passes to the model and the loss): one device-side gather, no
    per-class Python loop."""
    lut = cmask_lut(w).to(labels.device)
    idx = labels.long().clamp_(0, len(w) - 1)
    out = lut[idx].to(labels.dtype)
    valid = (labels >= 0) & (labels < len(w))
    return torch.where(valid, out, labels)
