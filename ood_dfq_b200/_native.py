"""ctypes binding of include/oodfq_b200.h -- the only way Python reaches the kernels.

There is no fallback: if the library is missing or a call fails, a RuntimeError is
raised.  Build it with ``python -m ood_dfq_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liboodfq_b200.so")

MODE_FAKEQUANT, MODE_QUANTIZE, MODE_DEQUANTIZE = 0, 1, 2
SYMMETRIC, PARAMS_GIVEN, RELU_FIRST, NO_ONCHIP, ONCHIP_TMA = 1, 2, 4, 8, 16
BN_RELU, BN_QUANT, BN_NHWC, BN_POOL_REGISTER = 1, 2, 4, 16
AUG_SRC_NHWC = 8
ABI_VERSION = 4

_vp, _ll, _i, _d = C.c_void_p, C.c_longlong, C.c_int, C.c_double


class WeightDesc(C.Structure):
    """oodfq_weight_desc"""
    _fields_ = [("w", _vp), ("wq", _vp), ("lo", _vp), ("hi", _vp), ("codes", _vp),
                ("rows", _ll), ("row_len", _ll), ("k", _i), ("flags", _i)]


# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "oodfq_abi_version": (_i, []),
    "oodfq_last_error": (C.c_char_p, []),
    "oodfq_launch_count": (C.c_ulonglong, []),
    "oodfq_reset_launch_count": (None, []),
    "oodfq_workspace_bytes": (C.c_size_t, []),
    "oodfq_quant_params": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "oodfq_fq_forward": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "oodfq_act_calib_forward": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "oodfq_act_calib_stats_forward": (_i, [_vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "oodfq_minmax": (_i, [_vp, _ll, _vp, _vp, _vp]),
    "oodfq_weight_fq_multi": (_i, [C.POINTER(WeightDesc), _i, _vp]),
    "oodfq_bn_stats_forward": (_i, [_vp, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "oodfq_bn_stats_finalize": (_i, [_vp, _vp, _i, _d, _vp, _vp, _vp]),
    "oodfq_bns_loss": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_i), C.POINTER(_d), _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "oodfq_bn_stats_backward": (_i, [_vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _d, _vp, _i, _vp]),
    "oodfq_bn_eval_forward": (_i, [_vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _i, _vp, _vp]),
    "oodfq_bn_pool_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _i, _vp]),
    "oodfq_bn_pool_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp]),
    "oodfq_defer_folds_begin": (_i, [_vp, C.c_size_t]),
    "oodfq_defer_folds_flush": (_i, [_vp]),
    "oodfq_defer_folds_end": (_i, [_vp]),
    "oodfq_defer_folds_pending": (_i, []),
    "oodfq_global_avgpool_forward": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp]),
    "oodfq_global_avgpool_backward": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp]),
    "oodfq_channel_energy_scratch_floats": (C.c_size_t, [_i, _i]),
    "oodfq_channel_energy_forward": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp, _vp]),
    "oodfq_channel_energy_backward": (_i, [_vp, _vp, _vp, _i, _i, _ll, _i, _vp]),
    "oodfq_res_tail_scratch_floats": (C.c_size_t, [_i, _i]),
    "oodfq_res_tail_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.c_float,
                                    _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _i, _vp]),
    "oodfq_res_tail_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.c_float,
                                     _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _vp]),
    "oodfq_act_mse_scratch_doubles": (C.c_size_t, [_i]),
    "oodfq_act_mse_search": (_i, [_vp, _ll, _vp, _i, _i, _d, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp]),
    "oodfq_bn_eval_stats_forward": (_i, [_vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _i, _vp, _vp,
                                         _vp, _vp]),
    "oodfq_bn_eval_tap_backward": (_i, [_vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _vp, _d,
                                        _vp, _vp]),
    "oodfq_fa_loss_max_layers": (_i, []),
    "oodfq_fa_loss_forward": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), _i, _i, C.c_float, _vp, _vp, _vp, _vp]),
    "oodfq_fa_loss_backward": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), _i, _i, C.c_float, _vp,
                                    C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "oodfq_s2d_stem_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "oodfq_s2d_stem_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "oodfq_crop_resize_flip": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "oodfq_crop_resize_flip_backward": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "oodfq_bn_eval_backward": (_i, [_vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.c_float, _i, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load (once) and type the library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("OODFQ_LIB")       # A/B runs of two builds of the kernels (tools/): load exactly this file
    if override:
        lib = C.CDLL(override)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.oodfq_abi_version() != ABI_VERSION:
            raise RuntimeError(f"ood_dfq_b200: ABI {lib.oodfq_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib
    if os.path.exists(LIB_PATH):
        # a library older than its sources has the wrong ABI as often as not: rebuild it when nvcc is at hand
        # (one process at a time: the ranks of a torchrun job all come through here)
        try:
            import fcntl
            from . import build as _build
            if _build.stale():
                with open(LIB_PATH + ".lock", "w") as lock:
                    fcntl.flock(lock, fcntl.LOCK_EX)
                    if _build.stale():
                        _build.build()
        except Exception:          # no nvcc / read-only tree: use what is there (the ABI version check still runs)
            pass
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"ood_dfq_b200: native library not found at {LIB_PATH}. This package has no CPU or "
            "PyTorch fallback; build the sm_100a kernels with `python -m ood_dfq_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.oodfq_abi_version() != ABI_VERSION:
        raise RuntimeError(f"ood_dfq_b200: ABI {lib.oodfq_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().oodfq_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"ood_dfq_b200.{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().oodfq_launch_count())


def reset_launch_count():
    load().oodfq_reset_launch_count()
