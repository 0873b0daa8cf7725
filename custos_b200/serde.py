"""Host side of the buffer (de)serialisation: the codec of `cbm_buffer_serialize` / `cbm_buffer_deserialize`
without a device (`cb_serde_encode` / `cb_serde_decode`).

A `CUDAPtr<T>` serialises as the sequence of its elements (src/devices/cuda/cuda_ptr.rs:122-157); the two
encodings are what serde_json and bincode 1.x make of such a sequence.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .expr import NP_DTYPE, dtype_code

JSON, BINCODE = N.SER_JSON, N.SER_BINCODE


def encode(values, dtype, fmt: int = JSON) -> bytes:
    dt = dtype_code(dtype)
    arr = np.ascontiguousarray(values, NP_DTYPE[dt])
    need = C.c_size_t()
    ptr = arr.ctypes.data_as(C.c_void_p)
    N.call("cb_serde_encode", dt, fmt, ptr, arr.size, None, 0, C.byref(need))
    out = C.create_string_buffer(max(need.value, 1))
    N.call("cb_serde_encode", dt, fmt, ptr, arr.size, out, need.value, C.byref(need))
    return out.raw[:need.value]


def decode(data: bytes, dtype, fmt: int = JSON) -> np.ndarray:
    dt = dtype_code(dtype)
    n = C.c_size_t()
    N.call("cb_serde_decode", dt, fmt, C.c_char_p(data), len(data), None, 0, C.byref(n))
    out = np.empty(n.value, NP_DTYPE[dt])
    N.call("cb_serde_decode", dt, fmt, C.c_char_p(data), len(data), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n))
    return out
