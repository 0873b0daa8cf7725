"""Thin object wrapper over the device-level C ABI (cb_* in include/custos_b200.h).

This is what a Rust `CUDA<Mods>` device would call through FFI; tests and bench.py use it
to drive the kernels directly.  All arithmetic happens in libcustos_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import _native as N
from .expr import Chain, NP_DTYPE, dtype_code


class Expr:
    """A compiled (chain of) expression(s): one fused sm_100a kernel."""

    def __init__(self, handle, chain: Chain, kind: int):
        self.handle, self.chain, self.kind, self.dtype = handle, chain, kind, (chain.dtype if chain is not None else None)


class Event:
    def __init__(self, dev: "RawDevice"):
        self.dev = dev
        self.h = C.c_void_p()
        N.call("cb_event_create", dev.h, C.byref(self.h))

    def record(self):
        N.call("cb_event_record", self.dev.h, self.h)
        return self

    def sync(self):
        N.call("cb_event_sync", self.h)

    def elapsed_ms(self, end: "Event") -> float:
        ms = C.c_float()
        N.call("cb_event_elapsed_ms", self.h, end.h, C.byref(ms))
        return ms.value

    def __del__(self):
        try:
            if self.h:
                N.load().cb_event_destroy(self.h)
        except Exception:
            pass


class RawDevice:
    """`CUDA::new(idx)` at the FFI level (reference: src/devices/cuda/cuda.rs:53-67)."""

    def __init__(self, ordinal: int = -1):
        self.h = C.c_void_p()
        N.call("cb_device_create", ordinal, C.byref(self.h))
        self._owned = True

    @classmethod
    def from_handle(cls, handle) -> "RawDevice":
        self = cls.__new__(cls)
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self._owned = False
        return self

    def close(self):
        if self.h and self._owned:
            N.call("cb_device_destroy", self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------ properties
    @property
    def sm_count(self) -> int:
        v = C.c_int32()
        N.call("cb_device_sm_count", self.h, C.byref(v))
        return v.value

    @property
    def ordinal(self) -> int:
        v = C.c_int32()
        N.call("cb_device_ordinal", self.h, C.byref(v))
        return v.value

    @property
    def stream(self) -> int:
        v = C.c_void_p()
        N.call("cb_device_stream", self.h, C.byref(v))
        return v.value or 0

    @property
    def launches(self) -> int:
        v = C.c_uint64()
        N.call("cb_launch_count", self.h, C.byref(v))
        return v.value

    def sync(self):
        N.call("cb_sync", self.h)

    def event(self) -> Event:
        return Event(self)

    # ------------------------------------------------------------ memory
    def alloc(self, nbytes: int, zero: bool = True) -> int:
        p = C.c_uint64()
        N.call("cb_alloc", self.h, nbytes, 1 if zero else 0, C.byref(p))
        return p.value

    def free(self, dptr: int):
        N.call("cb_free", self.h, dptr)

    def mem_info(self):
        r, u = C.c_size_t(), C.c_size_t()
        N.call("cb_mem_info", self.h, C.byref(r), C.byref(u))
        return r.value, u.value

    def cache_retrieve(self, cursor: int, nbytes: int):
        p, hit = C.c_uint64(), C.c_int32()
        N.call("cb_cache_retrieve", self.h, cursor, nbytes, C.byref(p), C.byref(hit))
        return p.value, bool(hit.value)

    def h2d(self, dptr: int, arr: np.ndarray, offset_bytes: int = 0):
        arr = np.ascontiguousarray(arr)
        N.call("cb_h2d", self.h, dptr + offset_bytes, arr.ctypes.data_as(C.c_void_p), arr.nbytes)

    def d2h(self, dptr: int, n: int, dtype, offset_bytes: int = 0) -> np.ndarray:
        out = np.empty(n, dtype=NP_DTYPE[dtype_code(dtype)])
        N.call("cb_d2h", self.h, out.ctypes.data_as(C.c_void_p), dptr + offset_bytes, out.nbytes)
        return out

    def upload(self, arr: np.ndarray) -> int:
        """alloc_from_slice (src/devices/cuda/cuda.rs:124-137)."""
        arr = np.ascontiguousarray(arr)
        p = self.alloc(arr.nbytes, zero=False)
        self.h2d(p, arr)
        return p

    def host_alloc(self, nbytes: int, write_combined: bool = False) -> int:
        p = C.c_void_p()
        if write_combined:
            N.call("cb_host_alloc_ex", nbytes, 1, C.byref(p))
        else:
            N.call("cb_host_alloc", nbytes, C.byref(p))
        return p.value

    def host_free(self, p: int):
        N.call("cb_host_free", C.c_void_p(p))

    def h2d_async(self, dptr: int, host_ptr: int, nbytes: int):
        N.call("cb_h2d_async", self.h, dptr, C.c_void_p(host_ptr), nbytes)

    def d2h_async(self, host_ptr: int, dptr: int, nbytes: int):
        N.call("cb_d2h_async", self.h, C.c_void_p(host_ptr), dptr, nbytes)

    # ------------------------------------------------------------ plain kernels
    def clear(self, dtype, dptr: int, n: int):
        N.call("cb_clear", self.h, dtype_code(dtype), dptr, n)

    def fill(self, dtype, dptr: int, n: int, value):
        dt = dtype_code(dtype)
        isf = dt in (N.F32, N.F64, N.F16, N.BF16)
        N.call("cb_fill", self.h, dt, dptr, n, float(value) if isf else 0.0, 0 if isf else int(value))

    def copy(self, dtype, dst: int, dst_off: int, src: int, src_off: int, n: int):
        N.call("cb_copy", self.h, dtype_code(dtype), dst, dst_off, src, src_off, n)

    def binary(self, dtype, op: int, lhs: int, rhs: int, out: int, n: int):
        N.call("cb_binary", self.h, dtype_code(dtype), op, lhs, rhs, out, n)

    # ------------------------------------------------------------ expression kernels
    def compile(self, fs: Sequence[Callable] | Callable, dtype, kind: int = N.KERNEL_APPLY) -> Expr:
        if callable(fs) or not isinstance(fs, (list, tuple)):
            fs = [fs]
        chain = Chain(fs, dtype, 2 if kind == N.KERNEL_BINARY else 1)
        h = C.c_void_p()
        N.call("cb_expr_compile", self.h, chain.dtype, kind, chain.progs, chain.n_nodes, chain.n_progs, C.byref(h))
        return Expr(h, chain, kind)

    def set_lut(self, e: Expr, enabled: bool):
        """f16 / bf16 chains: allow / forbid the table-lookup kernel for this expression (cb_expr_set_lookup)."""
        N.call("cb_expr_set_lookup", e.handle, 1 if enabled else 0)

    def has_lut(self, e: Expr) -> bool:
        v = C.c_int32()
        N.call("cb_expr_has_lookup", e.handle, C.byref(v))
        return bool(v.value)

    def apply(self, e: Expr, src: int, dst: int, n: int):
        N.call("cb_apply", self.h, e.handle, src, dst, n)

    def unary_grad(self, e: Expr, lhs: int, lhs_grad: int, out_grad: int, n: int):
        N.call("cb_unary_grad", self.h, e.handle, lhs, lhs_grad, out_grad, n)

    def unary_grad_ex(self, e: Expr, lhs: int, lhs_grad: int, out_grad: int, n: int, flags: int = 0):
        """add_unary_grad or the backward of a whole fused chain (an Expr of kind KERNEL_CHAIN_GRAD compiled from
        the K forward closures followed by their K grad closures); flags: N.GRAD_SEED_ONES."""
        N.call("cb_unary_grad_ex", self.h, e.handle, lhs, lhs_grad, out_grad, n, flags)

    def apply_host(self, e: Expr, host_in: int, host_out: int, n: int):
        """out_host[i] = f(in_host[i]) with H2D / kernel / D2H overlapped chunk by chunk (raw host addresses)."""
        N.call("cb_apply_host", self.h, e.handle, C.c_void_p(host_in), C.c_void_p(host_out), n)

    def apply2(self, e: Expr, lhs: int, rhs: int, out: int, n: int):
        N.call("cb_apply2", self.h, e.handle, lhs, rhs, out, n)

    # ------------------------------------------------------------ reductions
    @staticmethod
    def acc_dtype(dtype):
        dt = dtype_code(dtype)
        if dt == N.U64:
            return np.uint64  # u64 sums wrap as unsigned and the mean divides unsigned
        return np.float32 if dt in (N.F32, N.F16, N.BF16) else (np.float64 if dt == N.F64 else np.int64)

    def sum(self, dtype, dptr: int, n: int):
        out = np.zeros(1, dtype=self.acc_dtype(dtype))
        N.call("cb_sum_host", self.h, dtype_code(dtype), dptr, n, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def mean(self, dtype, dptr: int, n: int):
        out = np.zeros(1, dtype=self.acc_dtype(dtype))
        N.call("cb_mean_host", self.h, dtype_code(dtype), dptr, n, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def sum_into(self, dtype, dptr: int, n: int, out_dptr: int):
        N.call("cb_sum", self.h, dtype_code(dtype), dptr, n, out_dptr)

    # ------------------------------------------------------------ CUDA graphs
    def graph_begin(self):
        N.call("cb_graph_begin", self.h)

    def graph_end(self):
        g = C.c_void_p()
        N.call("cb_graph_end", self.h, C.byref(g))
        return g

    def graph_launch(self, g):
        N.call("cb_graph_launch", self.h, g)

    def graph_destroy(self, g):
        N.call("cb_graph_destroy", g)

    def graph_kernel_nodes(self, g) -> int:
        n = C.c_size_t()
        N.call("cb_graph_node_count", g, C.byref(n))
        return n.value


def sum_plan(dtype, n: int):
    b, t, v, t2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    ch = C.c_size_t()
    N.call("cb_sum_plan", dtype_code(dtype), n, C.byref(b), C.byref(ch), C.byref(t), C.byref(v), C.byref(t2))
    return dict(blocks=b.value, chunk=ch.value, threads=t.value, vec=v.value, threads2=t2.value)


def shard_range(n: int, elem_bytes: int, n_ranks: int, rank: int):
    b, e = C.c_size_t(), C.c_size_t()
    N.call("cb_shard_range", n, elem_bytes, n_ranks, rank, C.byref(b), C.byref(e))
    return b.value, e.value


class Comm:
    """One NCCL communicator per process/GPU; only reductions communicate."""

    def __init__(self, dev: RawDevice, n_ranks: int, rank: int, unique_id: bytes):
        self.dev = dev
        self.h = C.c_void_p()
        N.call("cb_comm_create", dev.h, n_ranks, rank, unique_id, C.byref(self.h))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(N.COMM_ID_BYTES)
        N.call("cb_comm_unique_id", buf)
        return buf.raw

    @property
    def uses_peer_memory(self) -> bool:
        v = C.c_int32()
        N.call("cb_comm_uses_peer_memory", self.h, C.byref(v))
        return bool(v.value)

    def sum_into(self, dtype, dptr: int, n_local: int, out_dptr: int):
        N.call("cb_comm_sum", self.h, dtype_code(dtype), dptr, n_local, out_dptr)

    def mean_into(self, dtype, dptr: int, n_local: int, n_global: int, out_dptr: int):
        N.call("cb_comm_mean", self.h, dtype_code(dtype), dptr, n_local, n_global, out_dptr)

    def check(self):
        """Synchronises and raises if a peer missed an exchange (the device scalar is then not the global sum)."""
        N.call("cb_comm_check", self.h)

    def sum(self, dtype, dptr: int, n_local: int):
        out = np.zeros(1, dtype=RawDevice.acc_dtype(dtype))
        N.call("cb_comm_sum_host", self.h, dtype_code(dtype), dptr, n_local, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def mean(self, dtype, dptr: int, n_local: int, n_global: int):
        out = np.zeros(1, dtype=RawDevice.acc_dtype(dtype))
        N.call("cb_comm_mean_host", self.h, dtype_code(dtype), dptr, n_local, n_global, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def close(self):
        if self.h:
            N.call("cb_comm_destroy", self.h)
            self.h = None
