"""ctypes binding of libhct_b200.so -- the thin layer between the PyTorch host code and the C ABI.

PyTorch is used for device memory, streams and autograd bookkeeping only; every kernel on the hot
path is reached through the `extern "C"` entry points declared in include/hct_b200.h.  There is
no CPU or eager-PyTorch fallback: if the library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libhct_b200.so")
_lib: Optional[C.CDLL] = None

EPI_BF16, EPI_GELU_BF16, EPI_RES_F32, EPI_POS_F32, EPI_DGELU_BF16, EPI_F32, EPI_ATOMIC_F32, EPI_GELU_DERIV_BF16, EPI_MUL_BF16 = range(9)


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("epilogue", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("bias", C.c_void_p),
        ("res", C.c_void_p), ("ldres", C.c_int64),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("pos", C.c_void_p), ("ldpos", C.c_int64),
        ("pos_idx", C.c_void_p), ("pos_period", C.c_int32),
        ("rows_in", C.c_int32), ("rows_out", C.c_int32), ("row_off", C.c_int32),
        ("alpha", C.c_float),
        ("splits", C.c_int32),
        ("colsum", C.c_void_p),
    ]


_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGNATURES = {
    "hct_gemm_bf16": [C.POINTER(GemmDesc), _P],
    "hct_layernorm_fwd": [_P, _P, _P, _P, _I32, _P, _P, _I64, _I32, _F, _P],
    "hct_layernorm_bwd": [_P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    "hct_cast_f32_to_bf16": [_P, _P, _I64, _P],
    "hct_cast_bf16_to_f32": [_P, _P, _I64, _P],
    "hct_colsum": [_P, _I32, _I64, _P, _I64, _I32, _P],
    "hct_broadcast_rows": [_P, _P, _I32, _I32, _I64, _I32, _I32, _P],
    "hct_reduce_rows": [_P, _P, _I32, _I32, _I64, _I32, _I32, _P],
    "hct_copy_rows_f32_to_bf16": [_P, _I64, _I64, _I32, _P, _I64, _I64, _I32, _I32, _P],
    "hct_scatter_add_rows": [_P, _P, _P, _I64, _I32, _P],
    "hct_window_scale_stack": [_P, _I32, _P, _I32, _I64, _I64, _I32, C.POINTER(_F), C.POINTER(_F), _P],
    "hct_patchify": [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_mask_indices": [_P, _P, _P, _P, _I32, _I32, _I32, _P],
    "hct_gather_tokens": [_P, _P, _P, _I32, _I32, _I32, _I64, _I32, _I32, _P],
    "hct_scatter_tokens": [_P, _P, _P, _I32, _I32, _I32, _I64, _I32, _I32, _P],
    "hct_decoder_assemble": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_decoder_assemble_bwd": [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_mae_loss_fwd": [_P, _I32, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_mae_loss_bwd": [_P, _I32, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_gelu_bwd": [_P, _P, _P, _I64, _P],
    "hct_attention_fwd": [_P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_attention_bwd": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_attention_bwd_bias": [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_l2norm_fwd": [_P, _I32, _P, _P, _I64, _I32, _P],
    "hct_l2norm_bwd": [_P, _P, _P, _P, _I64, _I32, _P],
    "hct_weightnorm_fwd": [_P, _P, _P, _P, _I64, _I32, _P],
    "hct_weightnorm_bwd": [_P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    "hct_dino_loss": [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _F, _P],
    "hct_center_ema": [_P, _P, _F, _F, _I32, _P],
    "hct_ema_multi": [_P, _I32, _F, _P],
    "hct_grad_norms_multi": [_P, _I32, _P, _P],
    "hct_profile_enable": [_I32],
    "hct_gemm_set_cta_pair": [_I32],
    "hct_crop_resize_set_rows": [_I32],
    "hct_layernorm_set_bulk": [_I32],
    "hct_attention_set_tcgen05": [_I32],
    "hct_attention_set_merge_tail": [_I32],
    "hct_attention_set_dkdv32": [_I32],
    "hct_attention_set_bwd3": [_I32],
    "hct_attention_set_fwd2": [_I32],
    "hct_attention_set_tail_key": [_I32],
    "hct_attention_set_poly": [_I32, _I32],
    "hct_set_pdl": [_I32],
    "hct_attention_set_bwd3_drain": [_I32],
    "hct_attention_trace": [C.c_void_p],
    "hct_attention_trace3": [C.c_void_p],
    "hct_crop_resize_area": [_P, _I32, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_adjust_contrast": [_P, _P, _P, _I64, _I64, _P],
    "hct_flip_shift": [_P, _I32, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P],
    "hct_gaussian_smooth_axis": [_P, _P, _P, _P, _P, _I32, _I64, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_gemm_trace": [C.c_void_p],
    "hct_profile_collect": [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)],
    "hct_profile_collect_class": [_I32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)],
    "hct_adamw_multi": [_P, _I32, _P, _F, _F, _F, _F, _F, _F, _I32, _P],
    "hct_adamw_multi_dev": [_P, _I32, _P, _F, _P, _F, _F, _F, _P],
    "hct_lora_shuffle": [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P],
    "hct_colnorm_stats": [_P, _P, _I64, _I32, _F, _F, _P, _P, _P, _P, _P],
    "hct_colnorm_apply": [_P, _P, _P, _I32, _F, _P, _P, _I32, _I64, _I32, _P],
    "hct_colnorm_bwd": [_P, _I32, _P, _P, _P, _P, _P, _I64, _I32, _P],
    "hct_pool_attention_fwd": [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _F, _P],
    "hct_pool_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _F, _P],
    # fp32 mode
    "hct_split3_bf16": [_P, _I64, _I64, _I32, _I32, _P, _I64, _I32, _I32, _I32, _P],
    "hct_gelu_f32": [_P, _P, _I64, _P],
    "hct_gelu_bwd_f32": [_P, _P, _P, _I64, _P],
    "hct_copy_rows_f32": [_P, _I64, _I64, _I32, _P, _I64, _I32, _I32, _P],
    "hct_scatter_add_rows_f32": [_P, _P, _P, _I64, _I32, _P],
    "hct_attention_f32_fwd": [_P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_attention_f32_bwd": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_patchify_f32": [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_decoder_assemble_f32": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_decoder_assemble_bwd_f32": [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P],
    "hct_mae_loss_fwd_f32": [_P, _I32, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "hct_mae_loss_bwd_f32": [_P, _I32, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P],
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES) + ("hct_last_error", "hct_abi_version", "hct_launch_count")


def library_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m headct_foundation_b200.build` "
                "(there is no CPU / eager fallback for the hot path)")
        L = C.CDLL(_LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        L.hct_last_error.restype = C.c_char_p
        L.hct_last_error.argtypes = []
        L.hct_abi_version.restype = C.c_int
        L.hct_launch_count.restype = C.c_longlong
        if os.environ.get("HCT_GEMM_CTA_PAIR", "1") == "0":      # debugging aid: single-CTA tcgen05 kernel
            L.hct_gemm_set_cta_pair(0)
        if os.environ.get("HCT_ATTN_TCGEN05", "1") == "0":       # debugging aid: mma.sync attention only
            L.hct_attention_set_tcgen05(0)
        if "HCT_PDL" in os.environ:                              # A/B: programmatic dependent launch on / off
            L.hct_set_pdl(int(os.environ["HCT_PDL"] != "0"))
        if "HCT_ATTN_FWD2" in os.environ:                        # A/B: pipelined persistent forward on / off
            L.hct_attention_set_fwd2(int(os.environ["HCT_ATTN_FWD2"] != "0"))
        if "HCT_ATTN_BWD3" in os.environ:                        # A/B: pipelined persistent backward on / off
            L.hct_attention_set_bwd3(int(os.environ["HCT_ATTN_BWD3"] != "0"))
        _lib = L
    return _lib


PROF_CLASSES = {"gemm": 0, "attention_fwd": 1, "attention_bwd": 2, "layernorm_fwd": 3, "layernorm_bwd": 4, "mae_loss": 5,
                "clip_adamw": 6, "window": 7, "patchify": 8}


def profile_enable(*classes: str) -> None:
    """Bracket every launch of the named kernel classes with CUDA events (measurement aid, see hct_b200.h)."""
    mask = 0
    for c in classes:
        mask |= 1 << PROF_CLASSES[c]
    lib().hct_profile_enable(mask)


def profile_collect(cls: str):
    """(total_ms, total_work, launches) of a class since its last collect; work = flops (tensor classes) or bytes."""
    ms, wk, n = C.c_double(), C.c_double(), C.c_longlong()
    check(lib().hct_profile_collect_class(PROF_CLASSES[cls], C.byref(ms), C.byref(wk), C.byref(n)), "hct_profile_collect_class")
    return ms.value, wk.value, int(n.value)


def launch_count() -> int:
    return int(lib().hct_launch_count())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().hct_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)
