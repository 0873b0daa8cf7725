"""Python mirror of the reference's operator interface over the module-layer C ABI (cbm_*).

    dev = CUDA("Graph", "Lazy", "Base")            # CUDA::<Graph<Lazy<Base>>>::new(0)
    buf = dev.buffer([1., 2., 3., 4., 5.])
    out = dev.apply_fn(buf, lambda x: x.sin())     # ApplyFunction::apply_fn
    dev.optimize_mem_graph(); dev.unary_fusing(); dev.run()
    out.replace().read()

Names, argument meaning and error behaviour follow the reference (src/unary.rs,
src/op_traits.rs, src/features.rs, src/buffer.rs) so that the parity tests read like the
reference's own tests.  All work happens in libcustos_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Iterable, Sequence

import numpy as np

from . import _native as N
from .expr import NP_DTYPE, dtype_code, flatten
from .raw import RawDevice

MODULE_BITS = {"Base": N.CBM_BASE, "Cached": N.CBM_CACHED, "Lazy": N.CBM_LAZY, "Graph": N.CBM_GRAPH,
               "Autograd": N.CBM_AUTOGRAD}


class Buffer:
    """`Buffer<'a, T, CUDA<Mods>>` (src/buffer.rs:43-49): a handle, not the data."""

    def __init__(self, device: "CUDA", handle: int, dtype: int):
        self.device, self.handle, self.dtype = device, handle, dtype

    def __len__(self) -> int:
        n = C.c_size_t()
        N.call("cbm_buffer_len", self.device.h, self.handle, C.byref(n))
        return n.value

    len = __len__

    def id(self) -> int:
        v = C.c_uint64()
        N.call("cbm_buffer_id", self.device.h, self.handle, C.byref(v))
        return v.value

    def ptr(self) -> int:
        """Device address after `replace()`; 0 while a lazy buffer has no storage."""
        v = C.c_uint64()
        N.call("cbm_buffer_ptr", self.device.h, self.handle, C.byref(v))
        return v.value

    def shard(self):
        """(begin, end, global_len): the slice of the global buffer this rank holds."""
        b, e, g = C.c_size_t(), C.c_size_t(), C.c_size_t()
        N.call("cbm_buffer_shard", self.device.h, self.handle, C.byref(b), C.byref(e), C.byref(g))
        return b.value, e.value, g.value

    def replace(self) -> "Buffer":
        """Buffer::replace (src/modules/lazy.rs:430-455): reads always resolve by id."""
        return self

    def read(self) -> np.ndarray:
        if self.dtype is None:
            raise TypeError("an untyped buffer has no element type: use read_typed(dtype) or to_typed(dtype)")
        out = np.empty(len(self), NP_DTYPE[self.dtype])
        N.call("cbm_buffer_read", self.device.h, self.handle, out.ctypes.data_as(C.c_void_p), out.size)
        return out

    read_to_vec = read

    def write(self, data) -> None:
        arr = np.ascontiguousarray(data, NP_DTYPE[self.dtype])
        N.call("cbm_buffer_write", self.device.h, self.handle, arr.ctypes.data_as(C.c_void_p), arr.size)

    def clear(self) -> None:
        N.call("cbm_clear", self.device.h, self.handle)

    def require_grad(self) -> "Buffer":
        N.call("cbm_buffer_require_grad", self.device.h, self.handle)
        return self

    def requires_grad(self) -> bool:
        v = C.c_int32()
        N.call("cbm_buffer_requires_grad", self.device.h, self.handle, C.byref(v))
        return bool(v.value)

    def checkpoint(self) -> "Buffer":
        N.call("cbm_buffer_checkpoint", self.device.h, self.handle)
        return self

    def grad(self) -> "Buffer":
        g = C.c_uint64()
        N.call("cbm_grad", self.device.h, self.handle, C.byref(g))
        return Buffer(self.device, g.value, self.dtype)

    def backward(self) -> None:
        N.call("cbm_backward", self.device.h, self.handle)

    def backward_with(self, seed) -> None:
        arr = np.ascontiguousarray(seed, NP_DTYPE[self.dtype])
        N.call("cbm_backward_with", self.device.h, self.handle, arr.ctypes.data_as(C.c_void_p), arr.size)

    def empty_like(self) -> "Buffer":
        return self.device.new_buffer(self.dtype, len(self))

    def clone(self) -> "Buffer":
        out = C.c_uint64()
        N.call("cbm_clone_buf", self.device.h, self.handle, C.byref(out))
        return Buffer(self.device, out.value, self.dtype)

    def drop(self) -> None:
        """End of the Rust scope of the buffer."""
        N.call("cbm_buffer_drop", self.device.h, self.handle)

    # ------------------------------------------------------------ untyped views (src/devices/untyped/mod.rs:17-84)
    def storage_dtype(self) -> int:
        """The run-time tag of the storage (`UntypedData` / `CudaStorage` variant)."""
        v = C.c_int32()
        N.call("cbm_buffer_dtype", self.device.h, self.handle, C.byref(v))
        return v.value

    def to_untyped(self) -> "Buffer":
        """Buffer::to_untyped / as_untyped: the same buffer without static type information."""
        return Buffer(self.device, self.handle, None)

    as_untyped = to_untyped

    def to_typed(self, dtype) -> "Buffer | None":
        """Buffer::to_typed / as_typed: `None` when the storage holds another type."""
        dt = dtype_code(dtype)
        rc = N.load().cbm_buffer_matches_type(self.device.h, self.handle, dt)
        if rc == N.CB_ERR_TYPE_MISMATCH:
            return None
        N.check(rc)
        return Buffer(self.device, self.handle, dt)

    as_typed = to_typed

    def read_typed(self, dtype) -> "np.ndarray | None":
        """Buffer::read_typed::<OT>()"""
        typed = self.to_typed(dtype)
        return None if typed is None else typed.read()

    # ------------------------------------------------------------ serde (src/devices/cuda/cuda_ptr.rs:122-157)
    def serialize(self, fmt: int = N.SER_JSON) -> bytes:
        """The buffer as a serialised sequence of its elements: serde_json text or bincode bytes."""
        need = C.c_size_t()
        N.call("cbm_buffer_serialize", self.device.h, self.handle, fmt, None, 0, C.byref(need))
        out = C.create_string_buffer(max(need.value, 1))
        N.call("cbm_buffer_serialize", self.device.h, self.handle, fmt, out, need.value, C.byref(need))
        return out.raw[:need.value]

    def to_tokens(self) -> list:
        """The serde data-model view `serde_test` asserts on: Seq { len }, one token per element, SeqEnd."""
        dt = self.storage_dtype()
        name = {N.F32: "F32", N.F64: "F64", N.I8: "I8", N.I16: "I16", N.I32: "I32", N.I64: "I64", N.U8: "U8",
                N.U16: "U16", N.U32: "U32", N.U64: "U64", N.BOOL: "Bool"}[dt]
        vals = Buffer(self.device, self.handle, dt).read().tolist()
        return [("Seq", len(vals))] + [(name, v) for v in vals] + [("SeqEnd",)]


class CUDA:
    """`CUDA<Mods>`; the module stack is given outermost first, e.g. CUDA("Graph", "Lazy", "Base")."""

    def __init__(self, *modules: str, ordinal: int = 0, dtype=np.float32):
        bits = 0
        for m in modules:
            bits |= MODULE_BITS[m]
        self.modules = modules
        self.h = C.c_void_p()
        N.call("cbm_device_create", ordinal, bits, dtype_code(dtype), C.byref(self.h))
        raw = C.c_void_p()
        N.call("cbm_device_raw", self.h, C.byref(raw))
        self.raw = RawDevice.from_handle(raw)

    def close(self):
        if getattr(self, "comm", None) is not None:
            N.call("cbm_device_set_comm", self.h, None)
            self.comm.close()
            self.comm = None
        if self.h:
            N.call("cbm_device_destroy", self.h)
            self.h = None

    def __enter__(self):
        return self

    # ------------------------------------------------------------ sharded device (one process per GPU)
    def shard(self, n_ranks: int, rank: int, unique_id: bytes) -> "CUDA":
        """Makes this device one rank of a sharded device: `buffer_sharded` / `new_buffer_sharded` hand out the
        rank's contiguous slice, every operator works on the slice unchanged, `sum` / `mean` return the global value
        (one scalar per rank exchanged over NVLink, folded in rank order — identical bits on every rank)."""
        from .raw import Comm
        self.comm = Comm(self.raw, n_ranks, rank, unique_id)
        N.call("cbm_device_set_comm", self.h, self.comm.h)
        return self

    def buffer_sharded(self, global_data, dtype=None) -> Buffer:
        """`global_data` is the whole array; only this rank's slice is uploaded."""
        if dtype is None:
            dtype = global_data.dtype if isinstance(global_data, np.ndarray) else np.float32
        dt = dtype_code(dtype)
        arr = np.ascontiguousarray(global_data, NP_DTYPE[dt])
        out = C.c_uint64()
        N.call("cbm_buffer_from_host_sharded", self.h, dt, arr.ctypes.data_as(C.c_void_p), arr.size, C.byref(out))
        return Buffer(self, out.value, dt)

    def new_buffer_sharded(self, dtype, global_length: int) -> Buffer:
        dt = dtype_code(dtype)
        out = C.c_uint64()
        N.call("cbm_buffer_new_sharded", self.h, dt, global_length, C.byref(out))
        return Buffer(self, out.value, dt)

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------ buffers
    def buffer(self, data, dtype=None) -> Buffer:
        """device.buffer([..]) / Buffer::from((&device, [..]))"""
        if dtype is None:
            if isinstance(data, np.ndarray):
                dtype = data.dtype
            else:  # Python literals: f32 for floats (the modules' default T), i32 for integers (Rust's default)
                dtype = np.float32 if np.asarray(data).dtype.kind == "f" else np.int32
        dt = dtype_code(dtype)
        arr = np.ascontiguousarray(data, NP_DTYPE[dt])
        out = C.c_uint64()
        N.call("cbm_buffer_from_host", self.h, dt, arr.ctypes.data_as(C.c_void_p), arr.size, C.byref(out))
        return Buffer(self, out.value, dt)

    def new_buffer(self, dtype, length: int) -> Buffer:
        """Buffer::<T, _>::new(&device, len): zero initialised."""
        dt = dtype_code(dtype)
        out = C.c_uint64()
        N.call("cbm_buffer_new", self.h, dt, length, C.byref(out))
        return Buffer(self, out.value, dt)

    def retrieve(self, length: int, parents: Sequence[Buffer] = (), dtype=np.float32) -> Buffer:
        """Retriever::retrieve (src/devices.rs:172-186)"""
        dt = dtype_code(dtype)
        arr = (C.c_uint64 * max(len(parents), 1))(*[p.handle for p in parents])
        out = C.c_uint64()
        N.call("cbm_retrieve", self.h, dt, length, arr, len(parents), C.byref(out))
        return Buffer(self, out.value, dt)

    # ------------------------------------------------------------ operator traits
    def apply_fn(self, buf: Buffer, f: Callable) -> Buffer:
        """ApplyFunction::apply_fn (src/unary.rs:7-28)"""
        nodes, n = flatten(f, buf.dtype)
        out = C.c_uint64()
        N.call("cbm_apply_fn", self.h, buf.handle, nodes, n, C.byref(out))
        return Buffer(self, out.value, buf.dtype)

    def add_unary_grad(self, lhs: Buffer, lhs_grad: Buffer, out: Buffer, lhs_grad_fn: Callable) -> None:
        """UnaryGrad::add_unary_grad (src/unary.rs:31-58): lhs_grad += out * lhs_grad_fn(lhs)"""
        nodes, n = flatten(lhs_grad_fn, lhs.dtype)
        N.call("cbm_add_unary_grad", self.h, lhs.handle, lhs_grad.handle, out.handle, nodes, n)

    def unary_ew(self, buf: Buffer, forward_fn: Callable, grad_fn: Callable) -> Buffer:
        """UnaryElementWiseMayGrad::unary_ew (src/unary.rs:62-132)"""
        fwd, nf = flatten(forward_fn, buf.dtype)
        grd, ng = flatten(grad_fn, buf.dtype)
        out = C.c_uint64()
        N.call("cbm_unary_ew", self.h, buf.handle, fwd, nf, grd, ng, C.byref(out))
        return Buffer(self, out.value, buf.dtype)

    def _binary(self, op: int, lhs: Buffer, rhs: Buffer) -> Buffer:
        out = C.c_uint64()
        N.call("cbm_binary", self.h, op, lhs.handle, rhs.handle, C.byref(out))
        return Buffer(self, out.value, lhs.dtype)

    def add(self, lhs: Buffer, rhs: Buffer) -> Buffer:
        """AddEw::add (src/lib.rs:293-301)"""
        return self._binary(N.BIN_ADD, lhs, rhs)

    def mul(self, lhs: Buffer, rhs: Buffer) -> Buffer:
        """MulBuf::mul (README.md:96-122)"""
        return self._binary(N.BIN_MUL, lhs, rhs)

    def sub(self, lhs: Buffer, rhs: Buffer) -> Buffer:
        return self._binary(N.BIN_SUB, lhs, rhs)

    def div(self, lhs: Buffer, rhs: Buffer) -> Buffer:
        return self._binary(N.BIN_DIV, lhs, rhs)

    def binary_into(self, op: int, lhs: Buffer, rhs: Buffer, out: Buffer) -> None:
        """`launch_kernel1d(len, src, "add", &[&lhs, &rhs, &mut out, &len])` of the reference's tests
        (src/devices/cuda/lazy.rs:96-141): the caller owns `out`; recorded under Lazy."""
        N.call("cbm_binary_into", self.h, op, lhs.handle, rhs.handle, out.handle)

    def add_into(self, lhs: Buffer, rhs: Buffer, out: Buffer) -> None:
        self.binary_into(N.BIN_ADD, lhs, rhs, out)

    def mul_into(self, lhs: Buffer, rhs: Buffer, out: Buffer) -> None:
        self.binary_into(N.BIN_MUL, lhs, rhs, out)

    def clear(self, buf: Buffer) -> None:
        buf.clear()

    def clear_op(self, buf: Buffer) -> None:
        """`add_op(&mut out, |out, _| out.clear())` (src/modules/lazy.rs:733-738): recorded under Lazy."""
        N.call("cbm_clear_op", self.h, buf.handle)

    def copy_slice_to(self, source: Buffer, source_range: range, dest: Buffer, dest_range: range) -> None:
        """CopySlice::copy_slice_to (src/op_traits.rs:34-60)"""
        assert len(source_range) == len(dest_range)
        N.call("cbm_copy_slice", self.h, source.handle, source_range.start, dest.handle, dest_range.start,
               len(source_range))

    def copy_slice_all(self, source: Buffer, dest: Buffer, ranges: Iterable) -> None:
        for sr, dr in ranges:
            self.copy_slice_to(source, sr, dest, dr)

    def write_buf(self, dst: Buffer, src: Buffer) -> None:
        """WriteBuf::write_buf"""
        N.call("cbm_copy_slice", self.h, src.handle, 0, dst.handle, 0, len(src))

    def sum(self, buf: Buffer):
        out = np.zeros(1, RawDevice.acc_dtype(buf.dtype))
        N.call("cbm_sum", self.h, buf.handle, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def mean(self, buf: Buffer):
        out = np.zeros(1, RawDevice.acc_dtype(buf.dtype))
        N.call("cbm_mean", self.h, buf.handle, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def deserialize(self, data: bytes, dtype, fmt: int = N.SER_JSON) -> Buffer:
        """Deserialize for CUDAPtr<T>: collect the sequence, allocate, write (cuda_ptr.rs:141-156)."""
        dt = dtype_code(dtype)
        out = C.c_uint64()
        N.call("cbm_buffer_deserialize", self.h, dt, fmt, C.c_char_p(data), len(data), C.byref(out))
        return Buffer(self, out.value, dt)

    # ------------------------------------------------------------ Lazy
    def run(self) -> None:
        N.call("cbm_run", self.h)

    def exec_now(self, begin: int = 0, end: int | None = None) -> None:
        N.call("cbm_exec_now", self.h, begin, (1 << 64) - 1 if end is None else end)

    def exec_last_n(self, n: int) -> None:
        N.call("cbm_exec_last_n", self.h, n)

    def ops_count(self) -> int:
        v = C.c_size_t()
        N.call("cbm_ops_count", self.h, C.byref(v))
        return v.value

    def alloc_later(self) -> None:
        N.call("cbm_alloc_later", self.h)

    def op_hint_src(self, i: int) -> str:
        buf = C.create_string_buffer(1 << 14)
        N.call("cbm_op_hint_src", self.h, i, buf, len(buf))
        return buf.value.decode()

    def op_expr(self, i: int):
        """The compiled kernel behind recorded op i (after the fusing passes), usable with RawDevice.apply / apply_host;
        None for no-ops and the ahead-of-time kernels."""
        from .raw import Expr
        h = C.c_void_p()
        N.call("cbm_op_expr", self.h, i, C.byref(h))
        return Expr(h, None, N.KERNEL_APPLY) if h.value else None

    def set_graph_replay(self, enabled: bool) -> None:
        N.call("cbm_set_graph_replay", self.h, 1 if enabled else 0)

    def replay_kernel_nodes(self) -> int:
        v = C.c_size_t()
        N.call("cbm_replay_kernel_nodes", self.h, C.byref(v))
        return v.value

    # ------------------------------------------------------------ Graph
    def optimize_mem_graph(self) -> None:
        N.call("cbm_optimize_mem_graph", self.h)

    def unary_fusing(self) -> None:
        N.call("cbm_unary_fusing", self.h)

    def elementwise_fusing(self) -> None:
        """Beyond the reference: binary ops fuse with neighbouring unary chains (at most two inputs per kernel)."""
        N.call("cbm_elementwise_fusing", self.h)

    def cache_traces(self):
        cap = 1 << 16
        buf = (C.c_int64 * cap)()
        w = C.c_size_t()
        N.call("cbm_cache_traces", self.h, buf, cap, C.byref(w))
        flat, out, i = [int(buf[k]) for k in range(w.value)], [], 0
        while i < len(flat):
            out.append((flat[i], flat[i + 2:i + 2 + flat[i + 1]]))
            i += 2 + flat[i + 1]
        return out

    # ------------------------------------------------------------ Cached
    def cursor(self) -> int:
        v = C.c_uint64()
        N.call("cbm_cursor", self.h, C.byref(v))
        return v.value

    def set_cursor(self, cursor: int) -> None:
        N.call("cbm_set_cursor", self.h, cursor)

    def bump_cursor(self) -> None:
        """Cursor::bump_cursor (src/features.rs:68-111): what every cached retrieve does."""
        self.set_cursor(self.cursor() + 1)

    def range(self, *args):
        """device.range(..) (src/range.rs:9-60): the cursor at loop entry is restored at the start of every
        iteration; after the loop it stays where the last iteration left it.  `range(None)` / `range(a, None)`
        are the unbounded `..` / `a..` forms."""
        start = self.cursor()
        if len(args) == 1:
            lo, hi = 0, args[0]
        else:
            lo, hi = args[0], args[1]
        i = lo
        while hi is None or i < hi:
            self.set_cursor(start)
            yield i
            i += 1

    def span(self, storage: dict) -> None:
        """`span!(device, span_storage)` (src/range/span.rs:8-30): the first visit of a call site records the
        cursor, every later visit of the same site restores it.  The site is the caller's (file, line)."""
        import sys
        f = sys._getframe(1)
        site = (f.f_code.co_filename, f.f_lineno)
        if site in storage:
            self.set_cursor(storage[site])
        else:
            storage[site] = self.cursor()

    # ------------------------------------------------------------ Autograd
    def zero_grad(self) -> None:
        N.call("cbm_zero_grad", self.h)

    def disable_grad(self) -> None:
        N.call("cbm_set_grad_enabled", self.h, 0)

    def enable_grad(self) -> None:
        N.call("cbm_set_grad_enabled", self.h, 1)

    def sync(self) -> None:
        self.raw.sync()


class Untyped(CUDA):
    """`Untyped` (src/devices/untyped/untyped_device.rs:16-72): a `CUDA<Base>` whose buffers are told apart by
    a run-time storage tag.  Only the `AsType` types are accepted (matches_type.rs:28-71); ops dispatch on
    the tag like `untyped_binary_op!` (ops.rs:80-159) and fail where the reference hits `unimplemented!()`."""

    def __init__(self, ordinal: int = 0):
        super().__init__("Base", ordinal=ordinal)

    def buffer(self, data, dtype=None) -> Buffer:
        if dtype is None:
            dtype = data.dtype if isinstance(data, np.ndarray) else (
                np.float32 if np.asarray(data).dtype.kind == "f" else np.int64)
        dt = dtype_code(dtype)
        if not N.load().cbm_untyped_supports(dt):
            raise TypeError(f"dtype {dt} has no AsType impl: an Untyped device cannot hold it")
        return super().buffer(data, dtype=dt)

    @staticmethod
    def _tag(buf: Buffer) -> int:
        return buf.storage_dtype() if buf.dtype is None else buf.dtype

    def apply_fn(self, buf: Buffer, f) -> Buffer:
        typed = Buffer(self, buf.handle, self._tag(buf))
        out = super().apply_fn(typed, f)
        return out.to_untyped() if buf.dtype is None else out

    def _binary(self, op: int, lhs: Buffer, rhs: Buffer) -> Buffer:
        tl, tr = self._tag(lhs), self._tag(rhs)
        if tl != tr:
            raise NotImplementedError(f"untyped_binary_op: storages of type {tl} and {tr}")  # `_ => unimplemented!()`
        out = super()._binary(op, Buffer(self, lhs.handle, tl), Buffer(self, rhs.handle, tr))
        return out.to_untyped() if lhs.dtype is None else out
