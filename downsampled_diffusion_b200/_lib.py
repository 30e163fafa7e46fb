"""ctypes binding of libddb200.so (the C-ABI declared in include/ddb200.h).

There is no CPU or library fallback: if the shared object is missing this module raises at the
first call, and every kernel entry point raises ``RuntimeError`` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DD_LIB_PATH") or os.path.join(_HERE, "libddb200.so")      # DD_LIB_PATH: instrumented builds (scripts/timeline.py)

DD_F32, DD_BF16 = 0, 1
CONV_PRE_MISH, CONV_TANH, CONV_OUT_NCHW, CONV_IN_NCHW = 1, 2, 4, 8
TC_CONV3x3, TC_CONV1x1, TC_DOWN, TC_UPT = 0, 1, 2, 3
TC_W_PER_SAMPLE = 1
TC_SPLITK = 2
TC_PAIR = 4
TC_STRIDED_IN = 8

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes (restype is int unless noted); mirrors include/ddb200.h one to one
SIGNATURES = {
    "dd_q_sample": [_p, _p, _p, _p, _p, _p, _i, _i64, _p],
    "dd_posterior_step": [_p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _p, _i, _i64, _p],
    "dd_predict_x0": [_p, _p, _p, _p, _p, _i, _p, _i, _i64, _p],
    "dd_tick": [_p, _i, _p],
    "dd_mse_rowsum": [_p, _p, _p, _i, _i64, _f, _p],
    "dd_mse_rowsum_bwd": [_p, _p, _p, _p, _i, _i64, _f, _p],
    "dd_ema_update": [_p, _p, _i, _i, _f, _f, _p],
    "dd_grad_norm": [_p, _p, _i, _i, _f, _p, _p, _p],
    "dd_adam_ema_step": [_p, _p, _i, _i, _p, _f, _f, _f, _f, _f, _f, _i, _f, _f, _i, _p],
    "dd_bicubic2d": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_bicubic2d_bwd": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_gather_f32": [_p, _p, _i, _p, _p, _p, _p],
    "dd_q_sample_step": [_p, _p, _i64, _i, _p, _p, _i, _p, _i, _i64, _p],
    "dd_vlb_terms": [_p, _p, _p, _p, _i64, _i, _p, _p, _i, _i, _p, _p, _i64, _i, _i, _i64, _p],
    "dd_prior_kl": [_p, _f, _f, _p, _i, _i64, _p],
    "dd_fix_samples": [_p, _p, _i, _i, _i, _i, _p],
    "dd_nchw_to_nhwc": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_nchw_to_nhwc_pad": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_nhwc_to_nchw": [_p, _i, _p, _i, _i, _i, _i, _p],
    "dd_im2col3x3_nchw": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_time_bias": [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p],
    "dd_gn_stats": [_p, _i, _i, _i, _i, _i, _f, _p, _p],
    "dd_gn_mish": [_p, _p, _i, _i, _i, _i, _i, _p, _i, _f, _p, _p, _p, _i, _p, _i, _p, _p],
    "dd_gn_mish_sum": [_p, _i, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _i, _p, _i, _p, _p, _p],
    "dd_conv_tc_ln": [_p, _i, _p, _i, _p, _p, _p, _i, _f, _p, _i, _i, _i, _i, _p],
    "dd_layernorm_c": [_p, _p, _i, _i64, _i, _p, _p, _f, _p],
    "dd_linattn_core": [_p, _p, _i, _i, _i, _i, _i, _p, _i64, _p],
    "dd_linattn_mix": [_p, _i, _i, _i, _i, _i, _p, _i64, _p, _i, _p, _p],
    "dd_conv_direct": [_p, _p, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dd_avgpool2": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_upsample_nearest2": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_space_to_depth2": [_p, _p, _i, _i, _i, _i, _p],
    "dd_zero": [_p, _i64, _p],
    "dd_debug_set_timeline": [_p],
    "dd_debug_set_attn_timeline": [_p],
    "dd_conv_wgrad": [_p, _p, _i, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dd_colsum": [_p, _p, _i64, _i, _i, _i64, _p],
    "dd_dropout": [_p, _p, _i64, C.c_uint32, _p, _f, _p],
    "dd_gn_mish_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p],
    "dd_layernorm_c_bwd": [_p, _p, _p, _f, _i64, _i, _p, _i, _p, _p, _p],
    "dd_linattn_save": [_p, _i, _i, _i, _p, _p],
    "dd_linattn_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "dd_ew": [_i, _p, _p, _p, _i64, _f, _i, _p],
    "dd_sincos_emb": [_p, _p, _p, _i, _i, _p],
    "dd_pool2_sum": [_p, _p, _i, _i, _i, _i, _f, _p],
    "dd_unpool2": [_p, _p, _i, _i, _i, _i, _f, _p],
    "dd_conv_tc32": [_i, _p, _p, _i, _i, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "dd_conv_wgrad_tc32": [_i, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p],
    "dd_conv1x1_thin_in": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dd_conv1x1_thin_out": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dd_conv1x1_thin_wgrad": [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p],
    "dd_s2d_f32": [_p, _p, _i, _i, _i, _i, _i, _p],
    "dd_nhwc_to_chw_pad": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p],
    "dd_conv_tc_gn": [_i, _p, _i, _p, _i, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p],
    "dd_conv_tc": [_i, _p, _i, _p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i64, _p, _i, _p],
}
PLAIN = {"dd_conv_tc_splits": (C.c_int, [_i, _i, _i, _i, _i, _i]), "dd_conv_tc_gn_cluster": (C.c_int, [_i, _i, _i, _i, _i, _i]), "dd_gn_mish_sum_parts": (C.c_int, [_i, _i]), "dd_debug_max_clusters": (C.c_int, [_i]), "dd_conv_tc_gn_ws_floats": (C.c_int64, [_i, _i, _i, _i, _i, _i]), "dd_conv_tc_gn_ln_parts": (C.c_int, [_i, _i, _i, _i, _i, _i, _i]), "dd_conv_tc_tile_n": (C.c_int, [_i, _i, _i, _i, _i, _i]), "dd_linattn_ws_floats": (C.c_int64, [_i, _i, _i]), "dd_linattn_mix_ws_floats": (C.c_int64, [_i, _i, _i]), "dd_version": (C.c_int, []), "dd_device_ok": (C.c_int, []), "dd_last_error": (C.c_char_p, [])}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libddb200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C downsampled_diffusion_b200/csrc). There is no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name, (res, args) in PLAIN.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = res
        _lib = l
    return _lib


def exported_symbols():
    return list(SIGNATURES) + list(PLAIN)


class DDError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dd_last_error().decode("utf-8", "replace")
        raise DDError(f"libddb200 {what} failed (rc={rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "libddb200 takes CUDA tensors only (no CPU fallback)"
    assert t.is_contiguous(), "libddb200 takes contiguous tensors"
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return DD_F32
    if dt == torch.bfloat16:
        return DD_BF16
    raise ValueError(f"unsupported activation dtype {dt}")


# launch counter: bench.py reports how many of OUR kernels ran in the timed region
class _Counter:
    n = 0


def call(name: str, *args) -> None:
    _Counter.n += 1
    check(getattr(lib(), name)(*args), name)


def launches() -> int:
    return _Counter.n
