"""ctypes binding of libsba_attn.so (include/sba_attn.h).  No torch types cross this line:
only device pointers, sizes and a stream handle.  There is no CPU fallback: if the library
is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsba_attn.so")

SBA_F32, SBA_BF16 = 0, 1
SBA_MASK_REFERENCE, SBA_MASK_PER_SAMPLE = 0, 1
SBA_ALGO_AUTO, SBA_ALGO_SIMT, SBA_ALGO_MMA, SBA_ALGO_TCGEN05 = 0, 1, 2, 3
ABI_VERSION = 4
SBA_PHASE_ALL, SBA_PHASE_FIRST, SBA_PHASE_SECOND = 0, 1, 2

# every symbol include/sba_attn.h declares: (restype, argtypes)
SYMBOLS = {
    "sba_abi_version": (c_int, []),
    "sba_last_error": (c_char_p, []),
    "sba_last_launch_count": (c_int, []),
    "sba_attn_fwd": (c_int, [c_void_p] * 8 + [c_int] * 8 + [c_void_p]),
    "sba_last_algo": (c_int, []),
    "sba_attn_supported": (c_int, [c_int] * 8),
    "sba_attn_bwd_workspace_floats": (c_size_t, [c_int] * 4),
    "sba_attn_bwd": (c_int, [c_void_p] * 10 + [c_size_t] + [c_void_p] * 2 + [c_int] * 8 + [c_void_p]),
    "sba_attn_fwd_phase": (c_int, [c_void_p] * 8 + [c_int] * 8 + [c_void_p]),
    "sba_attn_bwd_phase": (c_int, [c_void_p] * 10 + [c_size_t] + [c_void_p] * 2 + [c_int] * 8 + [c_void_p]),
    "sba_attn_fwd_into": (c_int, [c_void_p] * 5 + [c_int] * 2 + [c_void_p] * 3 + [c_int] * 7 + [c_void_p]),
    "sba_attn_bwd_from": (c_int, [c_void_p] * 7 + [c_int] * 2 + [c_void_p] * 3 + [c_size_t] + [c_void_p] * 2 + [c_int] * 7
                          + [c_void_p]),
    "sba_words_sim_fwd": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_float] * 4 + [c_void_p]),
    "sba_words_sim_fwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "sba_words_sim_fwd_ws": (c_int, [c_void_p] * 6 + [c_size_t] + [c_int] * 6 + [c_float] * 4 + [c_void_p]),
    "sba_words_sim_bwd_tc_workspace_bytes": (c_size_t, [c_int] * 6),
    "sba_words_sim_bwd_tc": (c_int, [c_void_p] * 7 + [c_size_t] + [c_int] * 6 + [c_float] * 4 + [c_void_p]),
    "sba_words_sim_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "sba_words_sim_bwd": (c_int, [c_void_p] * 7 + [c_int] * 6 + [c_float] * 4 + [c_void_p]),
    "sba_func_attention": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float] + [c_void_p]),
    "sba_adain_fwd": (c_int, [c_void_p] * 3 + [c_int] * 2 + [c_void_p] + [c_int] * 4 + [c_float] + [c_void_p]),
    "sba_adain_bwd": (c_int, [c_void_p] * 4 + [c_int] * 2 + [c_void_p] + [c_int] + [c_void_p] + [c_int] * 4 + [c_void_p]),
    "sba_match_ce_fwd": (c_int, [c_void_p] * 5 + [c_int] + [c_void_p]),
    "sba_match_ce_bwd": (c_int, [c_void_p] * 6 + [c_int] + [c_void_p]),
    "sba_sent_scores_fwd": (c_int, [c_void_p] * 4 + [c_int] * 2 + [c_float] * 2 + [c_void_p]),
    "sba_sent_scores_bwd": (c_int, [c_void_p] * 7 + [c_int] * 2 + [c_float] * 2 + [c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sba_gan_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sba_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libsba_attn.so ABI {lib.sba_abi_version()} != binding {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sba_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def last_launch_count() -> int:
    return int(load().sba_last_launch_count())
