"""Loader for the UNMODIFIED reference modules from /root/reference (authoring container only).

Used by oracle/make_golden.py to generate the fixtures under tests/golden/.  /root/reference does not exist on
the GPU box, so nothing that runs there may import this file.  Shims follow SURVEY.md App. D.
"""
import sys
import types

REF = "/root/reference"


def load_reference():
    import torch

    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ["matplotlib", "matplotlib.pyplot", "tensorboardX", "SimpleITK", "nibabel"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.argv = [sys.argv[0]]
    import unet3D as ref_unet  # noqa
    from loss_functions import loss_partial as ref_lp  # noqa

    # F7: the reference uses autocast without importing it (loss_partial.py:4,90)
    ref_lp.autocast = lambda enabled=False: torch.autocast("cpu", enabled=enabled)
    import evaluate_amos as ref_eval  # noqa

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # evaluate_amos.py:242 on a CPU-only host
    return ref_unet, ref_lp, ref_eval
