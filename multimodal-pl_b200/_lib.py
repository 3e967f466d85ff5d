"""ctypes binding of libmmpl_b200.so (C ABI declared in include/mmpl_b200.h).

The product path has NO CPU fallback: if the shared library is missing or the device is not sm_100-class the
import / first call raises.  PyTorch is only the owner of device memory and streams here.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMPL_LIB: an alternative build of the same ABI (kernel experiments); the product library sits next to this file
_LIB_PATH = os.environ.get("MMPL_LIB") or os.path.join(_HERE, "libmmpl_b200.so")

F32, BF16 = 0, 1
ALGO_DIRECT, ALGO_TCGEN05, ALGO_TCGEN05_PSPLIT = 0, 1, 2

_c_int, _c_i64, _c_f32, _ptr, _c_size = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> argument types (return type is int unless listed in _RESTYPES)
_SIGNATURES = {
    "mmpl_version": [],
    "mmpl_last_error": [],
    "mmpl_check_device": [],
    "mmpl_launch_count": [],
    "mmpl_ws_weight_fwd": [_ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr],
    "mmpl_ws_weight_fwd_batched": [_ptr, _c_int, _c_int, _c_int, _ptr],
    "mmpl_ws_weight_bwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "mmpl_parity_split": [_ptr, _ptr] + [_c_int] * 6 + [_ptr],
    "mmpl_conv3d_fprop": [_ptr, _ptr, _ptr, _ptr] + [_c_int] * 10 + [_ptr, _ptr],
    "mmpl_conv3d_dgrad": [_ptr, _ptr, _ptr, _ptr] + [_c_int] * 10 + [_ptr, _ptr, _ptr],
    "mmpl_conv3d_wgrad": [_ptr, _ptr, _ptr] + [_c_int] * 10 + [_ptr, _c_size, _ptr],
    "mmpl_conv3d_wgrad_workspace": [_c_int] * 9,
    "mmpl_stem_im2col": [_ptr, _ptr] + [_c_int] * 5 + [_ptr],
    "mmpl_stem_tc_fwd": [_ptr, _ptr, _ptr, _ptr] + [_c_int] * 5 + [_ptr],
    "mmpl_stem_tc_wgrad": [_ptr, _ptr, _ptr] + [_c_int] * 5 + [_ptr],
    "mmpl_stem_conv_fwd": [_ptr, _ptr, _ptr] + [_c_int] * 6 + [_ptr],
    "mmpl_stem_conv_wgrad": [_ptr, _ptr, _ptr] + [_c_int] * 6 + [_ptr, _c_size, _ptr],
    "mmpl_stem_conv_wgrad_workspace": [_c_int] * 4,
    "mmpl_cls_fwd": [_ptr, _ptr, _ptr, _ptr, _c_int, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_cls_bwd": [_ptr] * 8 + [_c_int, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_ln_rows_fwd": [_ptr, _ptr, _ptr, _c_i64, _c_int, _c_f32, _c_int, _ptr],
    "mmpl_ln_rows_bwd": [_ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _ptr],
    "mmpl_token_stats": [_ptr, _ptr, _c_int, _ptr, _ptr] + [_c_int] * 10 + [_ptr],
    "mmpl_token_ema": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_f32, _ptr],
    "mmpl_volume_moments": [_ptr, _c_int, _c_i64, _ptr, _ptr],
    "mmpl_prepare_patch": [_ptr, _c_int, _ptr, _c_int] + [_c_int] * 10 + [_ptr, _ptr],
    "mmpl_atlas_patch": [_ptr, _ptr] + [_c_int] * 13 + [_ptr],
    "mmpl_patch_stats": [_ptr, _c_i64, _ptr, _ptr],
    "mmpl_augment_patch": [_ptr, _c_i64, _c_f32, ctypes.c_uint64, _c_f32, _c_f32, _c_f32, _ptr, _ptr],
    "mmpl_blur_axis": [_ptr, _ptr] + [_c_int] * 4 + [_ptr, _c_int, _ptr],
    "mmpl_space_to_depth2": [_ptr, _ptr] + [_c_int] * 8 + [_ptr],
    "mmpl_bias_lrelu_fwd": [_ptr, _ptr, _ptr, _c_i64, _c_int, _c_f32, _c_int, _ptr],
    "mmpl_bias_lrelu_bwd": [_ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_f32, _c_int, _ptr],
    "mmpl_upsample2x_ncdhw_fwd": [_ptr, _ptr, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_upsample2x_ncdhw_bwd": [_ptr, _ptr, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_gn_stats": [_ptr, _ptr, _c_int, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_gn_relu_fwd": [_ptr] * 8 + [_c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f32, _c_int, _ptr],
    "mmpl_gn_relu_bwd": [_ptr] * 15 + [_c_int, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f32, _c_int, _ptr],
    "mmpl_upsample2x_add_fwd": [_ptr, _ptr, _ptr] + [_c_int] * 6 + [_ptr, _ptr],
    "mmpl_upsample2x_bwd": [_ptr, _ptr] + [_c_int] * 6 + [_ptr],
    "mmpl_partial_loss_fwd": [_ptr, _ptr, _c_int, _ptr, _ptr, _c_int, _ptr, _ptr, _c_int, _c_i64, _c_int, _c_int, _ptr],
    "mmpl_partial_loss_bwd": [_ptr, _ptr, _c_int, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _c_int, _c_i64, _c_int, _c_int,
                              _ptr],
    "mmpl_masked_dice_fwd": [_ptr] * 5 + [_c_i64, _c_int, _c_int, _ptr],
    "mmpl_masked_dice_bwd": [_ptr] * 7 + [_c_i64, _c_int, _c_int, _ptr],
    "mmpl_sgd_step": [_ptr, _ptr, _ptr, _c_i64, _ptr, _c_f32, _c_f32, _c_f32, _c_int, _ptr],
    "mmpl_sw_blend": [_ptr, _ptr, _ptr, _ptr] + [_c_int] * 12 + [_ptr],
    "mmpl_cls_blend": [_ptr] * 7 + [_c_int] * 9 + [_ptr],
    "mmpl_accumulate_f32": [_ptr, _ptr, _c_i64, _ptr],
    "mmpl_cls_loss_fwd": [_ptr] * 4 + [_c_int] + [_ptr] * 2 + [_c_int] + [_ptr] * 2 + [_c_int, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_cls_loss_bwd": [_ptr] * 4 + [_c_int] + [_ptr] * 2 + [_c_int] + [_ptr] * 7 + [_c_int, _c_i64, _c_int, _c_int, _c_int, _ptr],
    "mmpl_sw_finalize": [_ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _c_int, _c_i64, _c_i64, _c_int, _ptr],
}
_RESTYPES = {"mmpl_last_error": ctypes.c_char_p, "mmpl_launch_count": ctypes.c_uint64,
             "mmpl_conv3d_wgrad_workspace": ctypes.c_size_t, "mmpl_stem_conv_wgrad_workspace": ctypes.c_size_t}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class GnBwdFuse(ctypes.Structure):
    """mmpl_gn_bwd_fuse (include/mmpl_b200.h)."""
    _fields_ = [("a", _ptr), ("beta", _ptr), ("ws", _ptr), ("a_is_parity_split", _c_int), ("head", _c_int)]


def build(verbose: bool = False) -> str:
    """Compile libmmpl_b200.so for sm_100a with nvcc (cross-compiles on a CPU-only host)."""
    out = subprocess.run(["make", "-j8", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libmmpl_b200.so failed")
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the multimodal-PL hot path)")
        l = ctypes.CDLL(_LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = l
    return _lib


def last_error() -> str:
    return lib().mmpl_last_error().decode()


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"libmmpl_b200 {what} failed (code {rc}): {last_error()}")


_device_ok = False


def require_device():
    """Raise unless the current CUDA device is a B200-class part (the library refuses anything else)."""
    global _device_ok
    if not _device_ok:
        if not torch.cuda.is_available():
            raise RuntimeError("multimodal-pl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        check(lib().mmpl_check_device(), "mmpl_check_device")
        _device_ok = True


def launch_count() -> int:
    return int(lib().mmpl_launch_count())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise RuntimeError(f"unsupported activation dtype {dt} (float32 or bfloat16)")


def p(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
