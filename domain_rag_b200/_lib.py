"""ctypes binding of libdomainrag_b200.so (the C ABI declared in include/domainrag_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception
is raised. The CPU oracle under oracle/ is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libdomainrag_b200.so"

c_void_pp = C.POINTER(C.c_void_p)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/domainrag_b200.h declares must be listed here
# (tests/test_cabi.py checks the header against this table and against the built library).
SIGNATURES = {
    "drag_last_error": (C.c_char_p, []),
    "drag_version": (C.c_int, []),
    "drag_device_count": (C.c_int, [c_int_p]),
    "drag_device_info": (C.c_int, [C.c_int, c_int_p, c_int_p, c_int_p, C.c_char_p, C.c_int]),
    "drag_index_create": (C.c_int, [C.c_int, C.c_int, c_void_pp]),
    "drag_index_destroy": (C.c_int, [C.c_void_p]),
    "drag_index_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "drag_index_adopt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "drag_index_reset": (C.c_int, [C.c_void_p]),
    "drag_index_ntotal": (C.c_int, [C.c_void_p, c_i64_p]),
    "drag_index_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "drag_index_search_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "drag_index_last_launch": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "drag_index_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "drag_index_last_scan_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "drag_topk_merge_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "drag_gemm_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_void_p]),
    "drag_gemm_qkv_rope": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "drag_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "drag_layernorm_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "drag_timestep_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "drag_euler_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                  C.c_void_p]),
    "drag_redux_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "drag_l2_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "drag_gemm_qkv_split": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "drag_vit_patchify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "drag_vit_assemble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p]),
    "drag_topk_exchange_buffer_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, c_i64_p]),
    "drag_topk_exchange_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drag_index_search_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drag_vit_create": (C.c_int, [C.c_void_p, c_void_pp]),
    "drag_vit_destroy": (C.c_int, [C.c_void_p]),
    "drag_vit_set_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "drag_vit_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "drag_vit_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "drag_flux_create": (C.c_int, [C.c_void_p, c_void_pp]),
    "drag_flux_destroy": (C.c_int, [C.c_void_p]),
    "drag_flux_set_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "drag_flux_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p]),
    "drag_conv2d_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drag_groupnorm_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "drag_upsample2x_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "drag_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "drag_nchw_to_nhwc_pad": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_float, C.c_float, C.c_void_p]),
    "drag_nhwc_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_float, C.c_float, C.c_void_p]),
    "drag_image_postprocess_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "drag_image_preprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "drag_pack_latents": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "drag_unpack_latents": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "drag_pack_fill_inputs": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int64, C.c_void_p]),
    "drag_axpby_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "drag_prof_enable": (C.c_int, [C.c_int]),
    "drag_prof_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "drag_debug_set": (C.c_int, [C.c_int, C.c_int]),
    "drag_launch_count": (C.c_int, [C.POINTER(C.c_int64), C.c_int]),
    "drag_stem_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_float, C.c_void_p, C.c_void_p]),
    "drag_stem_stats_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_float, C.c_void_p, C.c_void_p]),
}

_lib = None


class DragError(RuntimeError):
    """A libdomainrag_b200 call returned a non-zero status."""


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -m domain_rag_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    for item in filter(None, os.environ.get("DRAG_DEBUG_SET", "").split(",")):      # A/B knobs, see drag_debug_set
        key, _, value = item.partition("=")
        check(lib.drag_debug_set(int(key), int(value)), f"DRAG_DEBUG_SET={item}")
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().drag_last_error()
        raise DragError(f"{what or 'libdomainrag_b200'} failed (status {status}): "
                        f"{msg.decode(errors='replace') if msg else ''}")


def ptr(t) -> C.c_void_p:
    """Device/host pointer of a torch tensor or numpy array as c_void_p (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def current_stream_ptr(device=None) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count(reset: bool = False) -> int:
    """Kernels launched by the library in this process since the last reset (drag_launch_count)."""
    n = C.c_int64(0)
    check(load().drag_launch_count(C.byref(n), int(reset)), "drag_launch_count")
    return int(n.value)


def device_count() -> int:
    n = C.c_int(0)
    lib = load()
    rc = lib.drag_device_count(C.byref(n))
    return n.value if rc == 0 else 0
