"""Stage the reference's hot-path files, unmodified, into baseline/_ref (git-ignored; travels to the GPU box with the
repo snapshot) so that ``bench.py --impl reference`` can time the reference's OWN modules there.  Run in the authoring
container, where /root/reference exists: ``python oracle/stage_reference.py`` (also called by __graft_entry__.build()).
The sanctioned ``pip install --target baseline/_ref /root/reference`` fails at metadata generation -- the reference's
setup.py is a script that imports three absent modules, not a setuptools package -- hence this plain copy."""
import os
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["unet3D.py", "engine.py", "utils.py", "evaluate_amos.py", "supervise_mask.csv", "LICENSE",
         "loss_functions/__init__.py", "loss_functions/loss_partial.py", "loss_functions/losses.py", "loss_functions/loss.py"]


def stage(dst=None) -> bool:
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(src, out)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here", file=sys.stderr)
