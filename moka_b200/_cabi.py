"""ctypes binding of libmoka.so (the C ABI declared in include/moka.h).

There is no fallback: if the library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64,
                    c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmoka.so')

MOKA_BF16, MOKA_F32 = 0, 1
VARIANT_AVT, VARIANT_VT = 0, 1
MOD_NONE, MOD_TEXT, MOD_VIS, MOD_AUD = 0, 1, 2, 3
MOD_MASK, KEY_VALID = 0x7, 0x8
FLAG_FORCE_SIMT, FLAG_FORCE_TC, FLAG_DELTA_OUT = 1, 2, 4
MOKA_EUNSUPPORTED = -4
META_EMPTY_QUESTION, META_OVERLAP, META_KEY_OVERFLOW, META_BAD_SEGMENT = 1, 2, 4, 8
SEG_END, SEG_TEXT, SEG_VIDEO, SEG_AUDIO, SEG_QUESTION = 0, 1, 2, 3, 4

EXPORTS = ['moka_workspace_bytes', 'moka_forward', 'moka_backward', 'moka_build_token_meta',
           'moka_build_token_meta_from_ids', 'moka_build_token_meta_from_segments',
           'moka_dropout_keep_mask', 'moka_effective_dropout_p', 'moka_adamw_step', 'moka_launch_count',
           'moka_uses_tensor_cores', 'moka_supports_delta', 'moka_strerror', 'moka_version', 'moka_rope', 'moka_debug_read',
           'moka_debug_trap_buffer', 'moka_debug_mark']


class MokaDesc(Structure):
    _fields_ = [('B', c_int32), ('S', c_int32), ('d_in', c_int32), ('d_out', c_int32),
                ('r', c_int32), ('M', c_int32), ('dtype', c_int32), ('Lk', c_int32),
                ('pre', c_float), ('post', c_float * 3), ('attn_w', c_float),
                ('dropout_p', c_float), ('flags', c_int32)]


class MokaError(RuntimeError):
    pass


_lib = None


def load():
    """Load libmoka.so, declaring every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MokaError(f'{LIB_PATH} not found: run `python -m moka_b200.build` (or __graft_entry__.build()); '
                        'there is no CPU or PyTorch fallback for MokaLinear')
    lib = ctypes.CDLL(LIB_PATH)
    D = POINTER(MokaDesc)
    lib.moka_workspace_bytes.restype = c_size_t
    lib.moka_workspace_bytes.argtypes = [D]
    lib.moka_forward.restype = c_int32
    lib.moka_forward.argtypes = [D, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64,
                                 c_void_p, c_size_t, c_void_p]
    lib.moka_backward.restype = c_int32
    lib.moka_backward.argtypes = [D, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_void_p,
                                  c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]
    lib.moka_build_token_meta.restype = c_int32
    lib.moka_build_token_meta.argtypes = [c_int32, c_int32, c_int32, c_int32, POINTER(c_void_p), c_int32,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.moka_build_token_meta_from_ids.restype = c_int32
    lib.moka_build_token_meta_from_ids.argtypes = [c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.moka_build_token_meta_from_segments.restype = c_int32
    lib.moka_build_token_meta_from_segments.argtypes = [c_int32, c_int32, c_int32, c_int32, c_void_p,
                                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.moka_dropout_keep_mask.restype = c_int32
    lib.moka_dropout_keep_mask.argtypes = [D, c_uint64, c_uint64, c_void_p, c_void_p]
    lib.moka_effective_dropout_p.restype = c_float
    lib.moka_effective_dropout_p.argtypes = [c_float]
    lib.moka_adamw_step.restype = c_int32
    lib.moka_adamw_step.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                    c_float, c_float, c_float, c_float, c_float, c_int32, c_float, c_void_p]
    lib.moka_rope.restype = c_int32
    lib.moka_rope.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p]
    lib.moka_debug_trap_buffer.restype = c_int32
    lib.moka_debug_trap_buffer.argtypes = [c_void_p]
    lib.moka_debug_mark.restype = c_int32
    lib.moka_debug_mark.argtypes = [c_int32, c_void_p]
    lib.moka_debug_read.restype = c_int32
    lib.moka_debug_read.argtypes = [c_void_p, c_int32]
    lib.moka_launch_count.restype = c_int64
    lib.moka_launch_count.argtypes = []
    lib.moka_uses_tensor_cores.restype = c_int32
    lib.moka_uses_tensor_cores.argtypes = [D]
    lib.moka_supports_delta.restype = c_int32
    lib.moka_supports_delta.argtypes = [D, c_int32]
    lib.moka_strerror.restype = c_char_p
    lib.moka_strerror.argtypes = [c_int32]
    lib.moka_version.restype = c_int32
    lib.moka_version.argtypes = []
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().moka_strerror(code).decode()
        raise MokaError(f'{what} failed: {msg} (code {code})')


def launch_count() -> int:
    return int(load().moka_launch_count())


def ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
