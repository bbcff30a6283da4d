"""ctypes binding of libhfbgpu.so (include/hfbgpu.h).  No torch types cross this boundary.

The library is the product; if it cannot be loaded this module raises -- there is no
Python or CPU fallback for the E-step.
"""
from __future__ import annotations

import ctypes as C
import os

from .flat import (hfb_acc_layout, hfb_batch, hfb_beams, hfb_compressed, hfb_model, hfb_options, hfb_stats,
                   hfb_utt_result)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhfbgpu.so")

EXPORTS = [
    "hfbgpu_abi_version", "hfbgpu_acc_layout", "hfbgpu_default_options", "hfbgpu_strerror",
    "hfbgpu_last_error", "hfbgpu_device_count", "hfbgpu_create", "hfbgpu_destroy", "hfbgpu_zero_accs",
    "hfbgpu_accumulate", "hfbgpu_accumulate_device", "hfbgpu_acc_device_ptr", "hfbgpu_acc_count",
    "hfbgpu_get_accs", "hfbgpu_set_accs", "hfbgpu_state_loglik", "hfbgpu_get_min_durs",
    "hfbgpu_get_stats", "hfbgpu_reset_stats", "hfbgpu_set_timing", "hfbgpu_set_stream", "hfbgpu_submit", "hfbgpu_wait",
    "hfbgpu_last_ticket", "hfbgpu_wait_ticket", "hfbgpu_create_multi", "hfbgpu_num_devices", "hfbgpu_reduce_accs",
    "hfbgpu_host_alloc", "hfbgpu_host_free", "hfbgpu_mstep", "hfbgpu_set_qualifiers", "hfbgpu_expand_features", "hfbgpu_accumulate_retrain",
    "hfbgpu_submit_compressed", "hfbgpu_accumulate_compressed", "hfbgpu_decompress_features",
]

_lib = None


class HfbError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        l = load()
        msg = l.hfbgpu_strerror(code).decode()
        detail = l.hfbgpu_last_error().decode()
        super().__init__("%s: %s (%d)%s" % (where, msg, code, (": " + detail) if detail else ""))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libhfbgpu.so is missing (%s): build it with `python -m htk_b200.build`; "
                           "the E-step has no CPU fallback" % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    l.hfbgpu_abi_version.restype = C.c_int
    l.hfbgpu_acc_layout.argtypes = [C.POINTER(hfb_model), C.POINTER(hfb_acc_layout)]
    l.hfbgpu_default_options.argtypes = [C.POINTER(hfb_options)]
    l.hfbgpu_default_options.restype = None
    l.hfbgpu_strerror.argtypes = [C.c_int]
    l.hfbgpu_strerror.restype = C.c_char_p
    l.hfbgpu_last_error.restype = C.c_char_p
    l.hfbgpu_device_count.restype = C.c_int
    l.hfbgpu_create.argtypes = [C.POINTER(vp), C.POINTER(hfb_model), C.POINTER(hfb_options)]
    l.hfbgpu_destroy.argtypes = [vp]
    l.hfbgpu_create_multi.argtypes = [C.POINTER(vp), C.POINTER(hfb_model), C.POINTER(hfb_options), C.POINTER(i32), i32]
    l.hfbgpu_num_devices.argtypes = [vp]
    l.hfbgpu_reduce_accs.argtypes = [vp]
    l.hfbgpu_zero_accs.argtypes = [vp]
    l.hfbgpu_accumulate.argtypes = [vp, C.POINTER(hfb_batch), C.POINTER(hfb_utt_result), C.POINTER(hfb_beams)]
    l.hfbgpu_accumulate_device.argtypes = l.hfbgpu_accumulate.argtypes
    l.hfbgpu_submit.argtypes = l.hfbgpu_accumulate.argtypes + [C.c_int]
    l.hfbgpu_wait.argtypes = [vp]
    l.hfbgpu_last_ticket.argtypes = [vp]
    l.hfbgpu_last_ticket.restype = i64
    l.hfbgpu_wait_ticket.argtypes = [vp, i64]
    l.hfbgpu_accumulate_retrain.argtypes = [vp, C.POINTER(hfb_batch), vp, C.POINTER(hfb_utt_result), C.POINTER(hfb_beams), C.c_int]
    l.hfbgpu_submit_compressed.argtypes = [vp, C.POINTER(hfb_batch), C.POINTER(hfb_compressed), C.POINTER(hfb_utt_result), C.POINTER(hfb_beams)]
    l.hfbgpu_accumulate_compressed.argtypes = l.hfbgpu_submit_compressed.argtypes
    l.hfbgpu_decompress_features.argtypes = [vp, C.POINTER(hfb_compressed), vp, i32, i32, vp]
    l.hfbgpu_acc_device_ptr.argtypes = [vp]
    l.hfbgpu_acc_device_ptr.restype = vp
    l.hfbgpu_acc_count.argtypes = [vp]
    l.hfbgpu_acc_count.restype = i64
    l.hfbgpu_get_accs.argtypes = [vp, vp]
    l.hfbgpu_set_accs.argtypes = [vp, vp]
    l.hfbgpu_state_loglik.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    l.hfbgpu_get_min_durs.argtypes = [vp, vp]
    l.hfbgpu_mstep.argtypes = [vp, vp, vp]
    l.hfbgpu_set_qualifiers.argtypes = [vp, vp]
    l.hfbgpu_expand_features.argtypes = [vp, vp, vp, i32, vp]
    l.hfbgpu_get_stats.argtypes = [vp, C.POINTER(hfb_stats)]
    l.hfbgpu_reset_stats.argtypes = [vp]
    l.hfbgpu_set_timing.argtypes = [vp, C.c_int]
    l.hfbgpu_set_stream.argtypes = [vp, vp]
    _lib = l
    return l
