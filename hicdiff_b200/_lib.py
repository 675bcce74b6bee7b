"""ctypes binding of `include/hicdiff_b200.h` (the C ABI of libhicdiff_b200.so).

The library is the ONLY compute path: if it is missing (or cannot be loaded) every entry point raises --
there is deliberately no eager/PyTorch/CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HD_ABI_VERSION = 1
HD_UNET, HD_UNET_SR3, HD_HICEDRN, HD_HICEDRN_SR3 = 0, 1, 2, 3

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libhicdiff_b200.so"
_lib = None


class hd_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("variant", C.c_int32),
        ("self_condition", C.c_int32),
        ("dim", C.c_int32),
        ("num_mults", C.c_int32),
        ("dim_mults", C.c_int32 * 8),
        ("image_size", C.c_int32),
        ("timesteps", C.c_int32),
        ("num_blocks", C.c_int32),
        ("debug_keep", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


_vp, _i32, _i64, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64

# name -> (restype, argtypes); every symbol declared in include/hicdiff_b200.h appears here.
SIGNATURES = {
    "hd_plan_create": (C.c_int, [C.POINTER(hd_config), C.POINTER(_vp)]),
    "hd_plan_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32, _vp]),
    "hd_plan_set_schedule": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "hd_plan_finalize": (C.c_int, [_vp, _vp]),
    "hd_plan_destroy": (None, [_vp]),
    "hd_eps_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "hd_ddpm_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _u64, _u64, _vp]),
    "hd_sample": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _u64, _u64, _i32, _i32, _vp]),
    "hd_ddrm_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_float, _i64, _u64, _u64, C.c_uint32, _vp]),
    "hd_ddim_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i32, _i64, _u64, _u64,
                               C.c_uint32, _vp]),
    "hd_ssim_mse_tiles": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "hd_trainer_create": (C.c_int, [C.POINTER(hd_config), _i32, C.POINTER(_vp)]),
    "hd_trainer_bind": (C.c_int, [_vp, C.c_char_p, _vp, _vp, C.POINTER(_i64), _i32]),
    "hd_trainer_finalize": (C.c_int, [_vp, _vp]),
    "hd_trainer_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "hd_trainer_set_grad_buckets": (C.c_int, [_vp, C.POINTER(C.c_char_p), _i32]),
    "hd_trainer_wait_grad_bucket": (C.c_int, [_vp, _i32, _vp]),
    "hd_trainer_num_launches": (C.c_int, [_vp]),
    "hd_trainer_profile": (C.c_int, [_vp, _i32, C.c_char_p, _i64, _vp]),
    "hd_trainer_device_bytes": (_i64, [_vp]),
    "hd_trainer_destroy": (None, [_vp]),
    "hd_adam_create": (C.c_int, [C.POINTER(_vp), C.POINTER(_i64), _i32, C.POINTER(_vp)]),
    "hd_adam_step": (C.c_int, [_vp, C.POINTER(_vp), C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _vp]),
    "hd_adam_state": (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "hd_adam_set_step": (C.c_int, [_vp, _i64]),
    "hd_adam_destroy": (None, [_vp]),
    "hd_op_conv3x3_wgrad": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "hd_op_conv_wgrad": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hd_op_groupnorm_silu_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "hd_op_channel_layernorm_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "hd_op_weight_standardize_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "hd_op_attention_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "hd_op_conv3x3_dgrad": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "hd_coo_to_dense": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "hd_remove_empty_bins": (C.c_int, [_vp, _i64, _vp, _vp, C.POINTER(_i64), _vp]),
    "hd_select_ranks": (C.c_int, [_vp, _i64, C.POINTER(_i64), _i32, C.POINTER(C.c_float), _vp]),
    "hd_normalize_contacts": (C.c_int, [_vp, _i64, C.c_float, _vp]),
    "hd_add_noise": (C.c_int, [_vp, _vp, C.c_float, _i64, _vp, _vp]),
    "hd_tile_count": (_i64, [_i64, _i32, _i32]),
    "hd_tile_extract": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp]),
    "hd_tile_scatter": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "hd_op_conv2d": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hd_op_conv_gn": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hd_op_groupnorm_silu": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "hd_op_channel_layernorm": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hd_op_linear_attention": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "hd_op_linattn_block": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, C.POINTER(C.c_float), _vp]),
    "hd_op_full_attention": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "hd_op_stem_conv": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "hd_op_philox_normal": (C.c_int, [_vp, _i64, _u64, _u64, _vp]),
    "hd_debug_read": (C.c_int, [_vp, _i32, C.c_char_p, _vp, C.POINTER(_i64), C.POINTER(_i32), _vp]),
    "hd_debug_names": (C.c_int, [_vp, _i32, C.c_char_p, _i64]),
    "hd_plan_launches_per_step": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "hd_plan_profile_step": (C.c_int, [_vp, _i32, _i32, C.c_char_p, _i64, _vp]),
    "hd_plan_device_bytes": (_i64, [_vp]),
    "hd_last_error": (C.c_char_p, []),
    "hd_abi_version": (C.c_int, []),
}


def lib_path() -> Path:
    return Path(os.environ.get("HICDIFF_B200_LIB", _LIB_PATH))


def load():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -m hicdiff_b200.build` (needs nvcc). "
            "hicdiff_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.hd_abi_version() != HD_ABI_VERSION:
        raise RuntimeError(f"libhicdiff_b200 ABI {lib.hd_abi_version()} != binding ABI {HD_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Translate a non-zero C return code into RuntimeError carrying hd_last_error()."""
    if rc != 0:
        msg = load().hd_last_error()
        raise RuntimeError(f"hicdiff_b200{': ' + what if what else ''}: {msg.decode() if msg else 'unknown error'}")


def ptr(t) -> int:
    """Device pointer of a (contiguous) torch tensor, or None."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
