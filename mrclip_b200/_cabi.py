"""ctypes binding of ``libmrclip.so`` (declared in ``include/mrclip.h``).

The library is built in-tree by ``make`` / ``__graft_entry__.build()``; there is no fallback of any
kind: if it is missing, or no sm_100 device is present, the callers raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrclip.so")

DT_F32, DT_BF16, DT_F16 = 0, 1, 2


class Shape(C.Structure):
    """``mrclip_shape`` (include/mrclip.h)."""
    _fields_ = [("m_rows", C.c_int), ("n_cols", C.c_int), ("d", C.c_int), ("label_offset", C.c_int)]


# name -> (restype, argtypes); every symbol include/mrclip.h declares
_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_long, C.c_float
SIGNATURES = {
    "mrclip_version": (_I, []),
    "mrclip_last_error": (C.c_char_p, []),
    "mrclip_device_ok": (_I, []),
    "mrclip_padded_dim": (_I, [_I]),
    "mrclip_padded_cols": (_I, [_I]),
    "mrclip_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "mrclip_fwd_col_granule": (_I, [_I, _I]),
    "mrclip_pack_bf16": (_I, [_P, _I, _I, _I, _L, _P, _I, _P]),
    "mrclip_transpose_bf16": (_I, [_P, _I, _I, _L, _P, _L, _P]),
    "mrclip_clip_fwd_tiles": (_I, [_P, _P, Shape, _I, _P, _I, _I, _P, _P]),
    "mrclip_clip_fwd_reduce": (_I, [Shape, _P, _P, _P, _P, _P, _P]),
    "mrclip_lse2_merge": (_I, [_P, _P, _I, _L, _I, _P, _P]),
    "mrclip_clip_loss": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "mrclip_clip_bwd": (_I, [_P, _P, _P, _L, Shape, _I, _P, _P, _P, _F, _F, _F, _P, _P, _P, _I, _L, _P, _I, _P]),
    "mrclip_siglip_fwd": (_I, [_P, _P, Shape, _I, _P, _P, _P, _P, _P]),
    "mrclip_siglip_bwd": (_I, [_P, _P, _P, _L, Shape, _I, _P, _P, _F, _P, _P, _P, _I, _L, _P, _P, _I, _P]),
    "mrclip_gmat_bytes": (C.c_size_t, [_I, _I]),
    "mrclip_clip_gwrite": (_I, [_P, _P, Shape, _I, _P, _P, _P, _F, _F, _F, _P, _P, _P, _P, _I, _I, _P]),
    "mrclip_siglip_gwrite": (_I, [_P, _P, Shape, _I, _P, _P, _F, _P, _P, _P, _P, _P, _I, _P]),
    "mrclip_gmat_gemm": (_I, [_I, _P, Shape, _P, _I, _F, _P, _P, _P, _P, _I, _L, _P]),
    "mrclip_clip_fwd_tiles_e": (_I, [_P, _P, Shape, _I, _P, _I, _I, _P, _P, _P]),
    "mrclip_emat_check": (_I, [Shape, _P, _P, _P, _P]),
    "mrclip_emat_flag": (_P, [Shape, _P]),
    "mrclip_clip_gwrite_if": (_I, [_P, _P, Shape, _I, _P, _P, _P, _F, _F, _P, _P, _P, _P]),
    "mrclip_emat_transform": (_I, [Shape, _P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _I, _I, _I, _P]),
    "mrclip_gmat_gemm_dot": (_I, [_I, _P, Shape, _P, _I, _F, _P, _P, _P, _P, _I, _L, _P, _P, _P]),
    "mrclip_siglip_fwd_e": (_I, [_P, _P, Shape, _I, _P, _P, _P, _P, _P, _P]),
    "mrclip_siglip_e_scalars": (_I, [Shape, _P, _F, _P, _P, _P, _I, _P]),
    "mrclip_gmat_gemm_push": (_I, [_P, Shape, _P, _I, _F, _P, _P, _P, _P, _I, _I, _P]),
    "mrclip_push_copy": (_I, [_P, C.c_size_t, _P, _I, C.c_size_t, _I, _P]),
    "mrclip_sum_slots": (_I, [_P, _I, _I, _I, _P, _I, _L, _P]),
    "mrclip_fwd_row_ent_ok": (_I, [_I, _I, _I]),
    "mrclip_clip_fwd_tiles_eu": (_I, [_P, _P, Shape, _I, _P, _I, _I, _P, _P, _P]),
    "mrclip_row_ent_split": (_I, [Shape, _P, _P, _I, _I, _P, _P]),
    "mrclip_sum_slots_dot": (_I, [_P, _I, _I, _I, _P, _I, _L, _P, _L, _P, _P]),
    "mrclip_gmat_gemm_push_bf16": (_I, [_P, Shape, _P, _I, _F, _P, _P, _P, _P, _I, _I, _P]),
    "mrclip_sum_slots_bf16": (_I, [_P, _I, _I, _I, _P, _I, _L, _P, _L, _P, _P]),
    "mrclip_peer_block_bytes": (C.c_size_t, []),
    "mrclip_step_small_floats": (C.c_size_t, []),
    "mrclip_step_struct_bytes": (C.c_size_t, []),
    "mrclip_peer_struct_bytes": (C.c_size_t, []),
    "mrclip_step_uses_fwd_ds": (_I, [_P]),
    "mrclip_step_forward": (_I, [_P, _P, _I, _L, _P, _I, _L, _P, _P, _I, _I, _P, _P]),
    "mrclip_normalize_bwd": (_I, [_P, _L, _P, _I, _I, _P, _I, _L, _P]),
    "mrclip_step_backward": (_I, [_P, _P, _P, _F, _P, _I, _L, _P, _I, _L, _P, _P, _P]),
    "mrclip_class_means": (_I, [_P, _I, _P, _P, _P, _I, _P, _P]),
    "mrclip_mpos_forward": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _F, _P, _P]),
    "mrclip_mpos_backward": (_I, [_P, _I, _L, _P, _I, _L, _P, _P, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P]),
    "mrclip_rank_collect": (_I, [_P, _P, Shape, _I, _P, _P, _P, _P, _P, _P]),
    "mrclip_rank_lmax": (_I, [_P, _P, _P, _I, _P, _P]),
    "mrclip_rank_count": (_I, [_P, _P, Shape, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "mrclip_prof_enable": (_I, [_I]),
    "mrclip_prof_report": (_I, [C.c_char_p, C.c_size_t]),
    "mrclip_launch_count": (_L, []),
}

_lib = None


def load():
    """Load the shared library and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "mrclip_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MrclipError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().mrclip_last_error()
        raise MrclipError(f"libmrclip error {rc}: {msg.decode() if msg else '?'}")
