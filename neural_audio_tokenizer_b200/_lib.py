"""ctypes binding of the C ABI in include/nat_b200.h (the shipped form of the stub shown in INTEGRATION.md).

The library is the product: using this module without `libnat_b200.so` raises, there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAT_B200_LIB") or os.path.join(_PKG_DIR, "libnat_b200.so")   # override: A/B builds

NAT_OK = 0
LAYOUT_BCT, LAYOUT_ROWS = 0, 1
CODES_I64, CODES_I32, CODES_I16 = 0, 1, 2
RVQ_DEFAULT, RVQ_EXACT_SCAN, RVQ_SINGLE_STREAM = 0, 1, 2
STAT_FIELDS = 4
PROF_FIELDS = 8
PROF_NAMES = ("prep", "gemm", "decide_update", "full_scan", "loss", "output", "gemm_launches", "wall")

# every symbol include/nat_b200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("nat_last_error", c_char_p, []),
    ("nat_abi_version", c_int, []),
    ("nat_device_info", c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), c_char_p, c_size_t]),
    ("nat_rvq_codebooks_create", c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    ("nat_rvq_codebooks_update", c_int, [c_void_p, POINTER(c_void_p), c_void_p]),
    ("nat_rvq_codebooks_destroy", c_int, [c_void_p]),
    ("nat_rvq_codebooks_dims", c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    ("nat_rvq_workspace_bytes", c_size_t, [c_void_p, c_int64]),
    ("nat_rvq_encode_f32", c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                   c_float, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    ("nat_rvq_sample_f32", c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                   c_float, POINTER(c_float), c_void_p, ctypes.c_uint64, ctypes.c_uint64, c_void_p,
                                   c_size_t, c_int, c_void_p]),
    ("nat_rvq_encode_profile_f32", c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_void_p,
                                           c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_int, c_void_p,
                                           POINTER(c_float)]),
    ("nat_launch_count", ctypes.c_ulonglong, []),
    ("nat_rvq_decode_f32", c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    ("nat_rvq_encode_host_f32", c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    ("nat_rvq_stacks_workspace_bytes", c_size_t, [POINTER(c_void_p), c_int, c_int64]),
    ("nat_rvq_encode_stacks_f32", c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), c_int, c_int64, c_int64, c_void_p,
                                          c_int, c_void_p, c_size_t, c_int, c_void_p]),
    ("nat_rvq_stacks_fused", c_int, [POINTER(c_void_p), c_int, c_int64]),
    ("nat_rvq_encode_stacks_aligned_f32", c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), POINTER(c_int64), c_int64,
                                                  c_int64, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    ("nat_rvq_encode_stacks_profile_f32", c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), c_int, c_int64, c_int64,
                                                  c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p,
                                                  POINTER(c_float)]),
    ("nat_host_ctx_create", c_int, [POINTER(c_void_p)]),
    ("nat_host_ctx_destroy", c_int, [c_void_p]),
    ("nat_tokenize_host_f32", c_int, [c_void_p, POINTER(c_void_p), c_int, c_void_p, c_int, c_int64, c_int64, c_void_p,
                                      c_int, c_void_p]),
    ("nat_mel_power_f32", c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    ("nat_mel_filterbank_bytes", c_size_t, [c_int]),
    ("nat_mel_filterbank_prepare", c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    ("nat_mel_power_banded_f32", c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    ("nat_spectral_stats_f32", c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    ("nat_mel_num_frames", c_int64, [c_int64, c_int]),
    ("nat_spectral_num_frames", c_int64, [c_int64, c_int, c_int]),
    ("nat_debug_rvq_scores", c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    ("nat_interp_linear_f32", c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    ("nat_token_histogram", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    ("nat_token_joint_histogram", c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    ("nat_ndjson_emit_frames", c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_int, c_int,
                                       c_char_p, ctypes.c_double, POINTER(c_void_p), POINTER(c_size_t)]),
    ("nat_free_host", None, [c_void_p]),
    ("nat_debug_stack_counters", c_int, [c_void_p, c_int, c_void_p, c_int, POINTER(c_int), POINTER(c_int)]),
    ("nat_peer_create", c_int, [c_int, c_int, c_size_t, c_size_t, POINTER(c_void_p)]),
    ("nat_peer_export", c_int, [c_void_p, c_void_p]),
    ("nat_peer_connect", c_int, [c_void_p, c_void_p]),
    ("nat_peer_all_gather", c_int, [c_void_p, c_void_p, c_size_t, c_void_p, POINTER(c_void_p)]),
    ("nat_peer_buffer", c_void_p, [c_void_p, c_int]),
    ("nat_peer_disconnect", c_int, [c_void_p]),
    ("nat_peer_destroy", None, [c_void_p]),
    ("nat_peer_last_error", c_char_p, []),
]
PEER_HANDLE_BYTES = 128
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib = None


class NatError(RuntimeError):
    """A non-zero status from the native library (message from nat_last_error())."""

    def __init__(self, code: int, message: str):
        super().__init__(f"nat_b200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load libnat_b200.so and bind every declared symbol. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m neural_audio_tokenizer_b200.build` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback for the RVQ / front-end path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in _SIGNATURES:
        fn = getattr(lib, name)            # AttributeError here means the .so is older than the header
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != NAT_OK:
        raise NatError(status, (load().nat_last_error() or b"").decode("utf-8", "replace"))


def check_peer(status: int) -> None:
    """check() for the nat_peer_* functions, which keep their own message."""
    if status != NAT_OK:
        raise NatError(status, (load().nat_peer_last_error() or b"").decode("utf-8", "replace"))
