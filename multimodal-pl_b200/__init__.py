"""multimodal-pl_b200: the dense hot path of TThuraya/multimodal-PL (unet3D encoder-decoder fwd/bwd + partial-label
loss, data-parallel training, sliding-window inference) on hand-written sm_100a kernels behind a C ABI
(include/mmpl_b200.h, libmmpl_b200.so).  Import name: ``multimodal_pl_b200`` (alias module at the repo root).
"""
from . import _lib, ops  # noqa: F401
from .ops import get_compute_dtype, set_compute_dtype, set_conv_algo, set_stem_mode  # noqa: F401

__all__ = ["ops", "set_compute_dtype", "get_compute_dtype", "set_conv_algo", "set_stem_mode"]
