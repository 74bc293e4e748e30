"""ctypes binding of libccst_b200.so (the C ABI declared in include/ccst_b200.h).

There is no fallback: if the shared library is missing or a call fails, the
caller gets an exception.  The library itself refuses every compute call on a
device that is not sm_100.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libccst_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "ccst_b200.h")

OK, EINVAL, EARCH, ECUDA, ESTATE = 0, -1, -2, -3, -4
PREC_FP32, PREC_BF16, PREC_FP16, PREC_FP16X3, PREC_BF16X3 = 0, 1, 2, 3, 4
ABI_VERSION = 2
FUSE_POOL, FUSE_UPSAMPLE, FUSE_STATS, FUSE_ADAIN, FUSE_TOTENSOR, FUSE_ALL = 1, 2, 4, 8, 16, 31

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every function declared in the header
PROTOTYPES = {
    "ccst_abi_version": (_i, []),
    "ccst_last_error": (C.c_char_p, []),
    "ccst_check_device": (_i, [_i]),
    "ccst_set_device": (_i, [_i]),
    "ccst_stats_nchw_f32": (_i, [_vp, _i64, _i64, _f, _i, _vp, _vp, _vp]),
    "ccst_welford_accumulate_nchw_f32": (_i, [_vp, _i, _i, _i64, _vp, _vp, _vp]),
    "ccst_welford_finalize": (_i, [_vp, _i, _f, _vp, _vp, _vp]),
    "ccst_welford_finalize_unbiased": (_i, [_vp, _i, _f, _vp, _vp, _vp]),
    "ccst_welford_to_sums": (_i, [_vp, _i, _vp, _vp, _vp]),
    "ccst_welford_to_moments": (_i, [_vp, _i, _vp, _vp]),
    "ccst_welford_from_moments": (_i, [_vp, _i, _vp, _vp]),
    "ccst_allreduce_moments": (_i, [_vp, _vp, _i64, _vp]),
    "ccst_adain_stat_nchw_f32": (_i, [_vp, _i, _i, _i64, _vp, _vp, _i64, _f, _f, _vp, _vp]),
    "ccst_adain_feat_nchw_f32": (_i, [_vp, _vp, _i, _i, _i64, _i64, _f, _f, _vp, _vp, _vp]),
    "ccst_create": (_vp, [_i]),
    "ccst_destroy": (None, [_vp]),
    "ccst_set_encoder_weights": (_i, [_vp, _pp, _pp]),
    "ccst_set_decoder_weights": (_i, [_vp, _pp, _pp]),
    "ccst_encoder_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "ccst_decoder_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "ccst_style_transfer": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i64, _f, _vp, _i, _vp]),
    "ccst_style_transfer_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i64, _f, _vp, _i, _vp]),
    "ccst_u8_to_tensor": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "ccst_quantize_u8": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "ccst_resize_pil_scratch_bytes": (_i64, [_i, _i, _i, _i, _i, _i]),
    "ccst_resize_pil_bilinear_u8": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ccst_resize_bilinear_aa_f32": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "ccst_encoder_levels": (_i, [_vp, _vp, _i, _i, _i, _vp, _pp, _pp, _f, _i, _vp]),
    "ccst_mse_f32": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "ccst_encoder_accumulate": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "ccst_encoder_accumulate_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "ccst_saturation_snapshot": (_i, [_vp, _vp, _vp]),
    "ccst_saturation_reset": (_i, [_vp, _vp]),
    "ccst_set_fusion": (_i, [_vp, _i]),
    "ccst_feature_hw": (None, [_i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "ccst_launch_count": (_i64, []),
    "ccst_profile_enable": (_i, [_vp, _i]),
    "ccst_profile_read": (_i, [_vp, _i, C.POINTER(_f), C.POINTER(C.c_double), C.POINTER(C.c_double),
                               C.POINTER(_i)]),
    "ccst_debug_conv3x3": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp]),
}

_lib = None


class CcstError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libccst_b200 error {code}: {msg}")
        self.code = code


def header_functions(path: str = HEADER_PATH):
    """Names of all functions declared in include/ccst_b200.h."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccst_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ccst_b200.build` "
                "(ccst_b200 has no CPU / PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.ccst_abi_version() != ABI_VERSION:
            raise RuntimeError("libccst_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(code: int) -> int:
    if code < 0:
        raise CcstError(code, lib().ccst_last_error().decode(errors="replace"))
    return code


class on_device:
    """Context manager: make `device` current for torch AND for the library's own runtime."""

    def __init__(self, device):
        import torch

        self.device = device
        self._ctx = torch.cuda.device(device)

    def __enter__(self):
        self._ctx.__enter__()
        idx = self.device.index
        if idx is None:
            import torch

            idx = torch.cuda.current_device()
        check(lib().ccst_set_device(idx))
        return self

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)


def feature_hw(h: int, w: int):
    fh, fw = C.c_int(), C.c_int()
    lib().ccst_feature_hw(h, w, C.byref(fh), C.byref(fw))
    return fh.value, fw.value
