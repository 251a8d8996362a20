"""ctypes binding of libtvit_b200.so (C ABI declared in include/tvit.h).

The library is built in-tree by ``neural_vit_b200.build.build()`` (nvcc, sm_100a).  There is no CPU
or PyTorch fallback: if the shared object is missing or the device is not a B200 the product path
raises immediately.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TVIT_LIB_PATH") or os.path.join(_HERE, "libtvit_b200.so")   # override: A/B builds only

F32, BF16 = 0, 1
ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1
EPI_STORE, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_GELU_BWD, EPI_ACCUM_F32, EPI_PATCH_EMBED, EPI_SOFTMAX_PROBS = range(7)

c_void_p, c_int, c_float, c_ll, c_size_t = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_longlong, ctypes.c_size_t)


class Dropout(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_ulonglong), ("site", ctypes.c_uint), ("p", ctypes.c_float)]


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("engine", c_int), ("dtype", c_int), ("trans_a", c_int), ("trans_b", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_ll), ("B", c_void_p), ("ldb", c_ll),
        ("epilogue", c_int), ("out", c_void_p), ("ldo", c_ll),
        ("bias", c_void_p), ("aux", c_void_p), ("ldaux", c_ll),
        ("resid", c_void_p), ("ldres", c_ll), ("gamma", c_void_p), ("row_scale", c_void_p),
        ("rows_per_group", c_int), ("drop", Dropout),
        ("pos_k", c_void_p), ("pos_f", c_void_p), ("pos_t", c_void_p),
        ("Kp", c_int), ("Fp", c_int), ("Tp", c_int), ("split_k", c_int), ("alpha", c_float),
        ("colsum", c_void_p),
    ]


class ShadowDesc(ctypes.Structure):
    """tvit_shadow_desc (include/tvit.h): one row of the device-resident table of tvit_shadow_t_multi."""
    _fields_ = [("w", c_void_p), ("row_scale", c_void_p), ("out_t", c_void_p), ("R", c_int), ("C", c_int),
                ("tile_begin", c_int), ("tiles_x", c_int)]


# name -> (restype, argtypes); every symbol include/tvit.h declares
SIGNATURES = {
    "tvit_last_error": (ctypes.c_char_p, []),
    "tvit_version": (c_int, []),
    "tvit_device_check": (c_int, [c_int]),
    "tvit_gemm": (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    "tvit_attn_keepbits_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tvit_attn_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              ctypes.POINTER(Dropout), c_void_p, c_void_p]),
    "tvit_attn_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "tvit_attn_bwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                              c_int, c_int, c_int, c_int, ctypes.POINTER(Dropout), c_void_p, c_void_p, c_void_p]),
    "tvit_attn_bwd_variant": (c_int, [c_int]),
    "tvit_attn_probs": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tvit_im2col": (c_int, [c_void_p, c_void_p, c_int] + [c_int] * 7 + [c_void_p]),
    "tvit_ln_fwd": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll, c_int,
                            c_float, c_void_p]),
    "tvit_ln_bwd": (c_int, [c_void_p, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(Dropout), c_void_p, c_ll,
                            c_int, c_void_p]),
    "tvit_branch_grad_prep": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, ctypes.POINTER(Dropout), c_void_p,
                                      c_int, c_void_p, c_void_p]),
    "tvit_colsum": (c_int, [c_void_p, c_int, c_ll, c_int, c_ll, c_void_p, c_void_p]),
    "tvit_cast_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "tvit_ls_finalize": (c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p]),
    "tvit_cls_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.POINTER(Dropout), c_void_p]),
    "tvit_embed_bwd_prep": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(Dropout), c_void_p, c_int,
                                    c_void_p, c_void_p, c_int, c_void_p]),
    "tvit_pos_grad_reduce": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_void_p]),
    "tvit_adamw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float,
                           c_float, c_int, c_float, c_void_p]),
    "tvit_shadow_t_multi": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "tvit_ce_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p]),
}

_lock = threading.Lock()
_lib = None
_checked_devices = set()


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().tvit_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libtvit_b200 {what} failed (code {rc}): {msg}")


def require_device(index: int) -> None:
    """Fail loudly unless CUDA device `index` is a B200-class (sm_100) GPU."""
    if index in _checked_devices:
        return
    check(load().tvit_device_check(int(index)), "device check")
    _checked_devices.add(index)
