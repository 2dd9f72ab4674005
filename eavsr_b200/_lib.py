"""ctypes binding of libeavsr_b200.so (the C ABI declared in include/eavsr_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a
call fails, the error is raised -- loudly -- to the caller.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint, c_uint64, c_void_p, POINTER
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libeavsr_b200.so"
_lib = None

F32, BF16 = 0, 1
FLOW_N2HW, FLOW_NHW2 = 0, 1
PAD_ZEROS, PAD_BORDER = 0, 1
DCN_FORCE_GENERIC = 1
DCN_FORCE_V1 = 2
DCN_FORCE_WS = 4
DCN_BLEND_FP32 = 8
DCN_WS_PACKED = 16
DCN_BWD_GENERIC_DATA = 32
DCN_BWD_GENERIC_WEIGHT = 64
DCN_FORCE_WIN1 = 128
DCN_FORCE_WIN2 = 256
CONV_SUMS_PREZEROED = 1
CORR_TF32 = 2

Strides = c_int64 * 4
_P64 = POINTER(c_int64)
_PF = c_void_p  # float* passed as raw device address


class EavsrError(RuntimeError):
    pass


class FlowTerm(ctypes.Structure):
    """EavsrFlowTerm of include/eavsr_b200.h (one term of a pyramid flow)."""
    _fields_ = [("flow", c_void_p), ("scaled_out", c_void_p), ("h", c_int), ("w", c_int), ("scale", c_float)]


class ConvLayer(ctypes.Structure):
    """EavsrConvLayer of include/eavsr_b200.h (one convolution of eavsr_conv3x3_chain_forward)."""
    _fields_ = [("x", c_void_p), ("packed_weight", c_void_p), ("bias", c_void_p), ("out", c_void_p),
                ("channel_sums", c_void_p), ("res", c_void_p), ("res_sums", c_void_p), ("w1", c_void_p),
                ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p), ("y_out", c_void_p),
                ("negative_slope", c_float)]


_SIGNATURES = {
    "eavsr_version": (c_int, []),
    "eavsr_last_error": (c_char_p, []),
    "eavsr_launch_count": (c_uint64, []),
    "eavsr_flow_warp_forward": (c_int, [c_void_p, _P64, _PF, c_int, c_void_p, _P64, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p]),
    "eavsr_flow_warp_backward": (c_int, [c_void_p, _P64, c_void_p, _P64, _PF, c_int, _PF, _P64, _PF, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_void_p]),
    "eavsr_flow_warp2_forward": (c_int, [c_void_p, _P64, c_void_p, _P64, _PF, c_int, c_void_p, _P64, c_void_p, _P64] +
                                 [c_int] * 6 + [c_void_p]),
    "eavsr_flow_warp_pyramid_forward": (c_int, [c_void_p, _P64, c_void_p, _P64, POINTER(FlowTerm), c_int, c_void_p, _P64,
                                                c_void_p, _P64, _PF] + [c_int] * 6 + [c_void_p]),
    "eavsr_spynet_level_input_forward": (c_int, [_PF, _PF, _PF, _PF] + [c_int] * 5 + [c_void_p]),
    "eavsr_backwarp_forward": (c_int, [c_void_p, _P64, _PF, c_void_p, _P64, c_void_p] + [c_int] * 5 + [c_void_p]),
    "eavsr_backwarp_backward": (c_int, [c_void_p, _P64, c_void_p, _P64, _PF, _PF, _P64, _PF] + [c_int] * 5 +
                                [c_void_p]),
    "eavsr_dcn_forward_workspace": (c_size_t, [c_int] * 7),
    "eavsr_dcn_forward_uses_tensor_cores": (c_int, [_P64, _P64] + [c_int] * 12 + [c_uint]),
    "eavsr_dcn_forward": (c_int, [c_void_p, _P64, _PF, _PF, c_void_p, c_void_p, c_void_p, _P64] + [c_int] * 16 +
                          [c_void_p, c_size_t, c_uint, c_void_p]),
    "eavsr_dcn_backward": (c_int, [c_void_p, _P64, c_void_p, _P64, _PF, _PF, c_void_p, _PF, _P64, _PF, _PF, _PF,
                                   _PF] + [c_int] * 16 + [c_void_p, c_size_t, c_uint, c_void_p]),
    "eavsr_dcn_backward_workspace": (c_size_t, [c_int] * 7),
    "eavsr_dcn_affine_forward": (c_int, [c_void_p, _P64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _P64] +
                                 [c_int] * 5 + [c_void_p, c_size_t, c_uint, c_void_p]),
    "eavsr_correlation_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]),
    "eavsr_correlation_forward_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_uint,
                                             c_void_p]),
    "eavsr_correlation_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                           c_int, c_int, c_void_p]),
    "eavsr_adapt_mix_forward": (c_int, [c_void_p] * 7 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "eavsr_affine_offsets_forward": (c_int, [c_void_p, _P64, c_void_p, _P64, c_void_p, _P64, c_void_p, c_void_p,
                                             c_void_p, _PF, _PF] + [c_int] * 5 + [c_void_p]),
    "eavsr_ca_residual_forward": (c_int, [c_void_p] * 9 + [c_int] * 6 + [c_void_p]),
    "eavsr_bias_act_forward": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_longlong, c_float, c_int, c_void_p]),
    "eavsr_grouped_conv3x3_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                              c_int, c_void_p]),
    "eavsr_grouped_conv3x3_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                               c_int, c_int, c_int, c_int, c_void_p]),
    "eavsr_channel_sum_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_int, c_void_p]),
    "eavsr_channel_dot_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_int, c_void_p]),
    "eavsr_nhwc_cat_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, ctypes.c_longlong, c_int,
                                       c_void_p]),
    "eavsr_bias_act_shuffle_forward": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "eavsr_ca_scale_forward": (c_int, [c_void_p, c_void_p, _PF] + [c_void_p] * 6 + [c_int] * 6 + [c_void_p]),
    "eavsr_conv3x3_packed_weight_bytes": (c_size_t, []),
    "eavsr_conv3x3_pack_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "eavsr_conv3x3_ca_forward": (c_int, [c_void_p, c_void_p, _PF] + [c_void_p] * 8 + [_PF] + [c_int] * 3 +
                                 [c_float, c_int, c_uint, c_void_p]),
    "eavsr_conv3x3_chain_forward": (c_int, [POINTER(ConvLayer), c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "eavsr_conv3x3_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _PF] + [c_int] * 5 +
                              [c_float, c_int, c_uint, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> Path:
    return Path(os.environ.get("EAVSR_B200_LIB", _LIB_PATH))


def load():
    """Load (once) and return the ctypes handle; raises EavsrError if the .so is not built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not path.exists():
            raise EavsrError(
                f"{path} not found: build it with `python -m eavsr_b200.build` "
                "(eavsr_b200 has no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(str(path))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().eavsr_last_error().decode("utf-8", "replace")
        raise EavsrError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().eavsr_launch_count())
