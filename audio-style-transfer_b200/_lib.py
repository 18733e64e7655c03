"""ctypes binding of ``libast_frontend.so`` (``include/ast_frontend.h``).

There is no CPU fallback: if the library is missing it is built in-tree with nvcc, and if that
fails, or no CUDA device is present when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libast_frontend.so")

AST_OK = 0
AST_ERR_INVALID_ARG = -1
AST_ERR_TOO_SHORT = -2
AST_ERR_NO_SECTIONS = -3
AST_ERR_WORKSPACE = -4
AST_ERR_CUDA = -5
AST_ERR_SHAPE = -6
LAYOUT_FLAT = 0
LAYOUT_SECTIONS = 1


class AstConfig(ctypes.Structure):
    _fields_ = [
        ("sample_rate", c_int32),
        ("n_fft", c_int32),
        ("hop", c_int32),
        ("n_bins", c_int32),
        ("window_size", c_int32),
        ("overlap_frames", c_int32),
        ("device", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/ast_frontend.h declares
SIGNATURES = {
    "ast_default_config": (c_int, [POINTER(AstConfig)]),
    "ast_plan_create": (c_int, [POINTER(AstConfig), POINTER(c_void_p)]),
    "ast_plan_destroy": (c_int, [c_void_p]),
    "ast_last_error": (c_char_p, []),
    "ast_version": (c_char_p, []),
    "ast_num_frames": (c_int32, [c_int64]),
    "ast_num_sections": (c_int32, [c_int32, c_int32, c_int32]),
    "ast_istft_length": (c_int64, [c_int32]),
    "ast_host_decimator_taps": (c_int, [POINTER(c_double), c_int32, POINTER(c_int32)]),
    "ast_host_cqt_kernel": (c_int, [POINTER(c_double)]),
    "ast_host_cqt_lengths": (c_int, [POINTER(c_double)]),
    "ast_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int64]),
    "ast_stats_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int64]),
    "ast_stft_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_int32, c_void_p]),
    "ast_cqt_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_size_t, c_void_p,
                                c_int32, c_void_p]),
    "ast_features_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                                     c_int32, c_float, c_void_p, c_size_t, c_void_p, c_int32, c_int32, c_void_p,
                                     c_void_p]),
    "ast_istft_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                  c_int64, c_void_p]),
    "ast_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "ast_concat": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "ast_overlap_windows": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "ast_sections_merge": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "ast_stats_accumulate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p,
                                     c_size_t, c_int32, c_void_p, c_void_p, c_void_p]),
    "ast_stats_accumulate_features": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                              c_void_p, c_size_t, c_int32, c_void_p, c_void_p, c_void_p]),
    "ast_stats_finalize": (c_int, [POINTER(c_double), c_double, POINTER(c_float), POINTER(c_float)]),
    "ast_resample_geometry": (c_int, [c_int32, c_int32, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "ast_resample_length": (c_int64, [c_int64, c_int32, c_int32]),
    "ast_host_resample_taps": (c_int, [c_int32, c_int32, POINTER(c_float), c_int32]),
    "ast_resampler_create": (c_int, [c_int32, c_int32, c_int32, POINTER(c_void_p)]),
    "ast_resampler_destroy": (c_int, [c_void_p]),
    "ast_load_audio_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p,
                                       c_int64, c_void_p]),
    "ast_mse_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "ast_mse_spectrogram": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p, c_void_p]),
    "ast_instrumentation_similarity_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "ast_instrumentation_similarity": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p,
                                                c_void_p]),
    "ast_synth_clips": (c_int, [c_void_p, c_int64, c_int32, c_int64, c_int64, c_int64, c_void_p]),
    "ast_profile_enable": (c_int, [c_int32]),
    "ast_profile_collect": (c_int, [c_char_p, POINTER(c_float), POINTER(c_int32), c_int32, POINTER(c_int32)]),
}

_lock = threading.Lock()
_lib = None


class AstError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[ast status {code}] {message}")
        self.code = code


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building in-tree first if needed) the CUDA library.  Never falls back to a CPU path."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"{LIB_PATH} is missing; run audio-style-transfer_b200/build.py")
            from . import build as _build

            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def profile_enable(on: bool) -> None:
    check(load().ast_profile_enable(1 if on else 0))


def profile_collect(capacity: int = 32) -> dict:
    """``{kernel name: (total_ms, launches)}`` recorded since ``profile_enable(True)``; clears the record."""
    lib = load()
    names = ctypes.create_string_buffer(32 * capacity)
    ms = (c_float * capacity)()
    cnt = (c_int32 * capacity)()
    n = c_int32(0)
    check(lib.ast_profile_collect(names, ms, cnt, capacity, ctypes.byref(n)))
    out = {}
    for i in range(n.value):
        name = names.raw[32 * i : 32 * (i + 1)].split(b"\0", 1)[0].decode()
        out[name] = (float(ms[i]), int(cnt[i]))
    return out


def check(code: int) -> None:
    """Map a status to the exception the reference would raise at the same point."""
    if code == AST_OK:
        return
    msg = load().ast_last_error().decode("utf-8", "replace")
    if code == AST_ERR_SHAPE:
        raise ValueError(msg)  # concat_stft_cqt raises ValueError, utilityFunctions.py:292-297
    if code in (AST_ERR_TOO_SHORT, AST_ERR_NO_SECTIONS):
        raise RuntimeError(msg)  # torch.stft / torch.stack raise RuntimeError in the reference
    raise AstError(code, msg)
