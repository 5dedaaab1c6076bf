"""ctypes binding of libamt_sm100.so (the C ABI of include/amt.h).

The library is the only compute path: if it is missing this module raises, and
every call that fails inside the library raises too -- there is no CPU or
PyTorch fallback behind these functions.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libamt_sm100.so")

AMT_ERR_ARG, AMT_ERR_DEVICE, AMT_ERR_CUDA, AMT_ERR_STATE, AMT_ERR_WORKSPACE = -1, -2, -3, -4, -5


class AmtError(RuntimeError):
    pass


class ModelConfig(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_mels", C.c_int), ("hidden", C.c_int), ("layers", C.c_int),
                ("heads", C.c_int), ("use_attention", C.c_int), ("use_onset_offset", C.c_int), ("precision", C.c_int)]


class LstmSeq(C.Structure):
    _fields_ = [("whh", C.c_void_p), ("gx", C.c_void_p), ("out_bf16", C.c_void_p), ("out_f32", C.c_void_p),
                ("H", C.c_int), ("reverse", C.c_int), ("ld_gx", C.c_int), ("ld_out", C.c_int), ("ld_out32", C.c_int)]


_SIGS = {
    "amt_version": (C.c_char_p, []),
    "amt_last_error": (C.c_char_p, []),
    "amt_device_check": (C.c_int, []),
    "amt_launch_count": (C.c_uint64, []),
    "amt_model_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "amt_pcm16_to_mono_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "amt_bce_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_model_profile_in_flight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "amt_model_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "amt_frontend_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_void_p)]),
    "amt_frontend_destroy": (C.c_int, [C.c_void_p]),
    "amt_frontend_num_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "amt_frontend_filterbank_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "amt_logmel_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    "amt_model_create": (C.c_int, [C.POINTER(ModelConfig), C.POINTER(C.c_void_p)]),
    "amt_model_destroy": (C.c_int, [C.c_void_p]),
    "amt_model_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "amt_model_finalize": (C.c_int, [C.c_void_p]),
    "amt_model_load": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "amt_model_get_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "amt_model_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "amt_model_workspace_layout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "amt_model_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_model_forward_db": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_sigmoid_threshold": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_pack_roll_u32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "amt_threshold_notes_scratch_ints": (C.c_size_t, [C.c_int, C.c_int]),
    "amt_threshold_notes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_float,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_bits_notes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_onset_notes_scratch_ints": (C.c_size_t, [C.c_int]),
    "amt_onset_notes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_f1_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p]),
    "amt_resample_poly_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "amt_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]),
    "amt_conv_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "amt_split3_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "amt_lstm_scratch_bytes": (C.c_size_t, [C.POINTER(LstmSeq), C.c_int, C.c_int]),
    "amt_lstm_recurrence": (C.c_int, [C.POINTER(LstmSeq), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (python -m music_transcription_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AmtError(f"{LIB_PATH} is missing: build it with `python -m music_transcription_b200.build` "
                           "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status == 0:
        return
    msg = lib().amt_last_error().decode("utf-8", "replace")
    if status == AMT_ERR_ARG:
        raise ValueError(msg)
    raise AmtError(f"libamt_sm100 error {status}: {msg}")


def ptr(t) -> int:
    """Raw device pointer of a tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise AmtError(f"{what} must be a CUDA tensor: the sm_100a kernels are the only compute path (got {t.device})")
