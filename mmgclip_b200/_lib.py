"""ctypes binding of ``libmmgclip_b200.so`` (the C ABI declared in ``include/mmgclip_b200.h``).

The library is the only way this package reaches the GPU.  There is deliberately no fallback: if the shared object
is missing the import fails with a build hint, and every entry point refuses host pointers.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libmmgclip_b200.so"
# MMGCLIP_B200_LIB: load an alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("MMGCLIP_B200_LIB") or os.path.join(_HERE, LIB_NAME)

MMG_PREC_FP32 = 0
MMG_PREC_BF16 = 1
MMG_PREC_F16 = 2
MMG_STORE = 0
MMG_ACCUMULATE = 1
MMG_ATOMIC_ADD = 2

# name -> (restype, argtypes); mirrors include/mmgclip_b200.h one to one (tests/test_cabi_symbols.py checks that
# every prototype in the header is present here and exported by the shared object).
SIGNATURES = {
    "mmg_version": (c_int, []),
    "mmg_last_error_string": (c_char_p, []),
    "mmg_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mmg_kernel_launch_count": (c_longlong, []),
    "mmg_gemm": (c_int, [c_int, c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_int, c_void_p, c_longlong,
                         c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mmg_gemm_split": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_longlong, c_int, c_void_p,
                               c_longlong, c_int, c_int, c_int, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mmg_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_cast_f32_to_f16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_cast_f32_to_bf16_split": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_push_rows": (c_int, [c_void_p, c_longlong, POINTER(c_void_p), c_int, c_longlong, c_void_p]),
    "mmg_l2norm_fwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mmg_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_longlong, c_void_p]),
    "mmg_dropout_apply": (c_int, [c_void_p, c_void_p, c_float, c_longlong, c_void_p]),
    "mmg_dropout_draw_apply": (c_int, [c_void_p, c_void_p, c_float, c_longlong, c_void_p, c_void_p]),
    "mmg_relu_dropout_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_longlong, c_void_p]),
    "mmg_colsum": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mmg_add": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_gelu_fwd": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "mmg_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "mmg_infonce_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mmg_infonce_fwd": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmg_infonce_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_void_p, c_void_p]),
    "mmg_infonce_row_part": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mmg_infonce_loss_cols": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "mmg_infonce_bwd_prep": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "mmg_infonce_bwd_prep_diag": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    "mmg_infonce_bwd_diag": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mmg_infonce_bwd": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mmg_infonce_bwd_owners": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "mmg_infonce_stored_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "mmg_infonce_fwd_store": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_longlong, c_void_p]),
    "mmg_infonce_bwd_stored": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int,
                                       c_void_p, c_size_t, c_void_p]),
    "mmg_eos_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mmg_eos_pool_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mmg_adamw_step": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                               POINTER(c_longlong), c_int, c_float, c_void_p, c_float, c_float, c_float, c_float,
                               c_void_p, c_void_p]),
    "mmg_debug_fused_trace_region": (c_int, [c_int, c_int, c_int, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_int),
                                             POINTER(c_int)]),
    "mmg_tune": (c_int, [c_char_p, c_int]),
    "mmg_fused_bwd_schedule": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int,
                                       POINTER(c_int)]),
    "mmg_ce_fwd": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "mmg_ce_bwd": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                           c_longlong, c_void_p]),
    "mmg_dot_sum": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    "mmg_zeroshot_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mmg_zeroshot_score_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmg_zeroshot_score": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class MmgError(RuntimeError):
    """A C entry point returned a negative status."""


def load() -> ctypes.CDLL:
    """Load the shared object (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C mmgclip_b200/csrc`) -- mmgclip_b200 has no CPU / PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = symbol missing from the build
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().mmg_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """Translate a C status into the reference's exception conventions (ValueError for bad arguments)."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (status {rc})" if what else f"{last_error()} (status {rc})"
    if rc in (-1, -2, -3):
        raise ValueError(msg)
    raise MmgError(msg)


def device_info():
    sm, major, minor = c_int(0), c_int(0), c_int(0)
    check(load().mmg_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "mmg_device_info")
    return sm.value, major.value, minor.value
