"""Loader for the UNMODIFIED reference modules (test / baseline infrastructure only).

Two roots: ``/root/reference`` (authoring container; used by oracle/make_golden*.py to write tests/golden/) and
``baseline/_ref`` -- a git-ignored staging copy of the reference's hot-path files made by ``oracle/stage_reference.py``
(run from ``__graft_entry__.build()`` when /root/reference is present; the offline ``pip install --target baseline/_ref
/root/reference`` cannot work because the reference's setup.py is a data-preparation script, not a package, see DESIGN.md
section 2).  The staging copy travels to the GPU box, where ``bench.py --impl reference`` and the ``cpu_baseline`` /
``gpu_library_baseline`` legs time the reference's own modules.  Nothing in the product imports this file.
Shims follow SURVEY.md App. D; the reference files themselves are never edited.
"""
import os
import sys
import types

REF = "/root/reference"
STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_root():
    """The first root that holds the reference's hot-path files, or None."""
    for root in (REF, STAGED):
        if os.path.isfile(os.path.join(root, "unet3D.py")) and os.path.isfile(
                os.path.join(root, "loss_functions", "loss_partial.py")):
            return root
    return None


def load_reference(root=None, device_type="cpu", with_eval=True):
    import torch

    root = root or reference_root()
    if root is None:
        raise FileNotFoundError("no reference tree: neither /root/reference nor baseline/_ref is populated")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ["matplotlib", "matplotlib.pyplot", "tensorboardX", "SimpleITK", "nibabel"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    argv = sys.argv
    sys.argv = [sys.argv[0]]
    try:
        import unet3D as ref_unet  # noqa
        from loss_functions import loss_partial as ref_lp  # noqa

        # F7: the reference uses autocast without importing it (loss_partial.py:4,90)
        ref_lp.autocast = lambda enabled=False: torch.autocast(device_type, enabled=enabled)
        ref_eval = None
        if with_eval:
            import evaluate_amos as ref_eval  # noqa
    finally:
        sys.argv = argv
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # evaluate_amos.py:242 on a CPU-only host
    return ref_unet, ref_lp, ref_eval
