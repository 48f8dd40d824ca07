"""ctypes binding of libmoc_b200.so (the C ABI declared in include/moc_b200.h).

There is no fallback: if the library is missing or a call fails, a ``MocError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOC_B200_LIB", os.path.join(PKG, "libmoc_b200.so"))  # override: developer experiments only

OK, E_ARG, E_SHAPE, E_WORKSPACE, E_CUDA = 0, -1, -2, -3, -4
CLS_TOPK, CLS_DELTA_SOFTMAX, CLS_DELTA_DIFF, CLS_BOTTOMK, CLS_ALL = 1, 2, 4, 8, 15
HEAD_WIDE_DOMAIN = 0x100   # MOC_HEAD_WIDE_DOMAIN: force the range-free 3xTF32 gate kernel
NUM_PARAMS = 64 * 512 + 64 + 4 * 64 + 4
MAX_COLS = 64

_CLS_BITS = {"topk": CLS_TOPK, "delta_softmax": CLS_DELTA_SOFTMAX, "delta_diff": CLS_DELTA_DIFF,
             "bottomk": CLS_BOTTOMK}


class MocError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("moc_b200 error %d: %s" % (code, msg))
        self.code = code


# named key planes (include/moc_b200.h: MOC_PLANE_*)
PLANE_TOP0, PLANE_SOFTMAX0, PLANE_DIFF, PLANE_BG_SUM, PLANE_BG_MAX, PLANE_LSE = range(6)

p, i32, i64, u32, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every function include/moc_b200.h declares
SIGNATURES = {
    "moc_last_error": (C.c_char_p, []),
    "moc_version": (i32, []),
    "moc_num_key_planes": (i32, [i32]),
    "moc_key_plane": (i32, [i32, i32]),
    "moc_expand_keys": (i32, [p, i64, i32, i64, p, i64, p]),
    "moc_packed_cols": (i32, [i32, i32]),
    "moc_pack_prompts": (i32, [p, i32, p, i32, p, p]),
    "moc_collapse_prompt_bank": (i32, [p, p, i32, p, p]),
    "moc_prompt_bank_tc_bytes": (sz, [i32, i32]),
    "moc_prompt_bank_tc_flag_offset": (sz, [i32, i32]),
    "moc_prepare_prompt_bank_tc": (i32, [p, p, i32, i32, p, i32, p, sz, p]),
    "moc_score_keys_bank_tc": (i32, [p, i64, p, i32, i32, i32, i32, p, i64, p]),
    "moc_score_keys": (i32, [p, i64, p, i32, i32, i32, p, i64, p]),
    "moc_score_keys_ex": (i32, [p, i64, p, i32, i32, i32, p, i64, i32, p]),
    "moc_prompts_tc_bytes": (sz, [i32, i32]),
    "moc_prompts_tc_flag_offset": (sz, [i32, i32]),
    "moc_prepare_prompts_tc": (i32, [p, i32, i32, p, sz, p]),
    "moc_score_keys_tc": (i32, [p, i64, p, i32, i32, i32, p, i64, p]),
    "moc_select_capacity": (i64, [i64, i32, i32]),
    "moc_select_workspace_bytes": (sz, [i64, i32]),
    "moc_select_union": (i32, [p, i64, p, i32, i64, i32, i32, u32, p, p, p, p, p, p, sz, p]),
    "moc_topj_sorted": (i32, [p, i64, i64, i32, i64, i32, i32, p, i64, p, p]),
    "moc_row_keys": (i32, [p, i64, i64, i32, i32, p, i64, p]),
    "moc_take_rows": (i32, [p, i64, p, i64, i32, p, p]),
    "moc_col_prefix_mean": (i32, [p, i64, i32, i32, p, p]),
    "moc_head_forward_workspace_bytes": (sz, []),
    "moc_head_domain_flag_offset": (sz, []),
    "moc_head_forward": (i32, [p, p, i64, i32, p, p, p, i32, i64, p, p, p, p, u32, i32, p, p, p, p, p, sz, p]),
    "moc_ablation_forward": (i32, [p, i64, i32, p, p, p, i32, i64, i32, i32, p, p, p]),
    "moc_pool_topk": (i32, [p, i64, p, i32, i32, i32, i32, i32, i32, i32, i32, p, p]),
    "moc_linear_workspace_bytes": (sz, [i32, i32]),
    "moc_linear_forward": (i32, [p, i64, i64, i32, p, p, i32, i32, i32, i32, p, i64, p, sz, p]),
    "moc_adapter_scores": (i32, [p, p, f32, p, i32, i64, p, i64, p]),
    "moc_gated_attention_scores": (i32, [p, i64, i32, p, f32, i64, p, p]),
    "moc_attention_pool_workspace_bytes": (sz, [i64, i32]),
    "moc_attention_pool": (i32, [p, p, i64, i32, i64, p, p, i32, p, p, p, p, p, sz, p]),
    "moc_row_softmax": (i32, [p, i64, i32, i64, p, i64, p]),
    "moc_cross_entropy": (i32, [p, p, i32, i32, f32, p, p, p, p]),
    "moc_head_backward_workspace_bytes": (sz, [i32, i32, i32]),
    "moc_head_backward": (i32, [p, p, i64, i32, p, p, p, i32, p, p, p, p, u32, i32, p, p, p, p, sz, p]),
    "moc_gather_selected": (i32, [p, p, i64, i32, p, i64, p, p, p, p, p, p]),
    "moc_senet_forward": (i32, [p, i64, p, p, p, p, p, u32, p, sz, p]),
    "moc_senet_backward_workspace_bytes": (sz, [i64]),
    "moc_senet_backward": (i32, [p, i64, p, p, p, p, p, p, p, sz, p]),
    "moc_linear_wgrad_workspace_bytes": (sz, [i64, i32, i32]),
    "moc_linear_wgrad": (i32, [p, i64, i32, p, i64, i32, i64, p, i64, i32, p, sz, p]),
    "moc_abmil_backward_workspace_bytes": (sz, [i64, i32, i32, i32]),
    "moc_abmil_backward": (i32, [p, i64, i32, i64, p, i64, i32, p, i64, i32, p, p, p, p, p, i32, p,
                                 p, p, p, p, p, p, p, p, p, sz, p]),
    "moc_adapter_backward_rows": (i32, [p, p, f32, p, i32, p, p, i64, p, p]),
    "moc_mask_positive": (i32, [p, p, i64, p]),
    "moc_transpose": (i32, [p, i32, i32, p, p]),
    "moc_mil_fc_backward": (i32, [p, i32, p, i32, p, i32, p, p, p, p, p, p]),
    "moc_h5_open": (i32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "moc_h5_close": (None, [p]),
    "moc_h5_dataset_info": (i32, [p, C.c_char_p, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
    "moc_h5_read": (i32, [p, C.c_char_p, p, sz]),
    "moc_adam_step": (i32, [p, p, p, p, i64, i64, f32, f32, f32, f32, f32, p]),
    "moc_accumulate": (i32, [p, p, i64, p]),
    "moc_adam_prepare_dev": (i32, [p, p, f32, f32, f32, p]),
    "moc_adam_apply_dev": (i32, [p, p, p, p, i64, p, f32, f32, f32, f32, p]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MocError(E_ARG, "%s not found - build it with `python -m moc_b200.build` "
                                      "(there is no CPU fallback)" % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(code: int) -> None:
    if code != OK:
        raise MocError(code, load().moc_last_error().decode("utf-8", "replace"))


def discard_bits(names) -> int:
    """--discard_classifiers names (main_moc.py:39) -> bit mask; unknown names are ignored like the reference."""
    bits = 0
    for n in names or ():
        bits |= _CLS_BITS.get(n, 0)
    return bits


def active_bits(discard_classifiers=(), mode: str = "train") -> int:
    """Which gated planes enter the sum.  train honours all four names (main_moc.py:396-403); evaluation
    always keeps top-k and tests the never-matching "delta_bottomk" (main_moc.py:486-492)."""
    d = set(discard_classifiers or ())
    if mode == "train":
        return CLS_ALL & ~discard_bits(d)
    bits = CLS_TOPK | CLS_BOTTOMK
    if "delta_softmax" not in d:
        bits |= CLS_DELTA_SOFTMAX
    if "delta_diff" not in d:
        bits |= CLS_DELTA_DIFF
    if "delta_bottomk" in d:
        bits &= ~CLS_BOTTOMK
    return bits
