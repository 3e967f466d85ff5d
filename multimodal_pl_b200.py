"""Import alias: the package directory is named ``multimodal-pl_b200`` (not a Python identifier); this module makes
it importable as ``multimodal_pl_b200`` (``import multimodal_pl_b200.unet3D`` etc.)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "multimodal-pl_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
_init = _os.path.join(__path__[0], "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
