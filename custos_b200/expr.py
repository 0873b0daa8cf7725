"""Host-side mirror of custos' `two_way_ops` expression DSL.

`Resolve` + the `Combiner` methods (reference: src/two_way_ops/resolve.rs:25-30,
src/two_way_ops/combiner.rs:8-116) build a small expression tree; `flatten()` turns it
into the `cb_node` array of the C ABI.  A Python closure `lambda x: x.mul(2.).add(1.).sin()`
plays the role of the Rust closure `|x| x.mul(2.).add(1.).sin()`.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import _native as N

OPS = ["x", "y", "const", "add", "mul", "sub", "div", "pow", "min", "max", "sin", "cos", "tan", "tanh",
       "exp", "ln", "abs", "neg", "identity", "geq", "leq", "eq"]
OP = {name: i for i, name in enumerate(OPS)}
BINARY = {"add", "mul", "sub", "div", "pow", "min", "max", "geq", "leq", "eq"}
UNARY = {"sin", "cos", "tan", "tanh", "exp", "ln", "abs", "neg", "identity"}

# numpy has no bfloat16: a bf16 buffer travels as its uint16 bit patterns (bf16_from_f32 / bf16_to_f32 below)
NP_DTYPE = {N.F32: np.float32, N.F64: np.float64, N.F16: np.float16, N.I32: np.int32, N.I64: np.int64,
            N.U32: np.uint32, N.U8: np.uint8, N.BF16: np.uint16, N.I8: np.int8, N.I16: np.int16,
            N.U16: np.uint16, N.U64: np.uint64, N.BOOL: np.bool_}
DTYPE_OF_NP = {np.dtype(v): k for k, v in NP_DTYPE.items() if k != N.BF16}
FLOAT_DTYPES = (N.F32, N.F64, N.F16, N.BF16)


def bf16_from_f32(a) -> np.ndarray:
    """f32 -> bf16 bit patterns, round to nearest even; NaN keeps its top payload bits plus the quiet bit
    (what `half::bf16::from_f32` does)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7fffffff) > 0x7f800000
    lower, q = u & 0xffff, u >> 16
    q = q + ((lower > 0x8000) | ((lower == 0x8000) & ((q & 1) == 1)))
    return np.where(nan, (u >> 16) | 0x40, q).astype(np.uint16)


def bf16_to_f32(bits) -> np.ndarray:
    return (np.ascontiguousarray(bits, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def dtype_code(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    return DTYPE_OF_NP[np.dtype(dtype)]


class Combiner:
    """A node of the expression tree.  Method names and operand order follow Combiner."""

    __slots__ = ("op", "lhs", "rhs", "value")

    def __init__(self, op: str, lhs=None, rhs=None, value=None):
        self.op, self.lhs, self.rhs, self.value = op, lhs, rhs, value

    @staticmethod
    def _wrap(v) -> "Combiner":
        return v if isinstance(v, Combiner) else Combiner("const", value=v)

    def _bin(self, op, rhs):
        return Combiner(op, self, Combiner._wrap(rhs))

    def add(self, rhs): return self._bin("add", rhs)
    def mul(self, rhs): return self._bin("mul", rhs)
    def sub(self, rhs): return self._bin("sub", rhs)
    def div(self, rhs): return self._bin("div", rhs)
    def pow(self, rhs): return self._bin("pow", rhs)
    def min(self, rhs): return self._bin("min", rhs)
    def max(self, rhs): return self._bin("max", rhs)
    def geq(self, rhs): return self._bin("geq", rhs)
    def leq(self, rhs): return self._bin("leq", rhs)
    def eq(self, rhs): return self._bin("eq", rhs)
    def sin(self): return Combiner("sin", self)
    def cos(self): return Combiner("cos", self)
    def tan(self): return Combiner("tan", self)
    def tanh(self): return Combiner("tanh", self)
    def exp(self): return Combiner("exp", self)
    def ln(self): return Combiner("ln", self)
    def abs(self): return Combiner("abs", self)
    def neg(self): return Combiner("neg", self)
    def identity(self): return Combiner("identity", self)


class Resolve(Combiner):
    """The closure argument: the seed value on the CPU, the marker in generated source."""

    def __init__(self, which: str = "x"):
        super().__init__(which)


def trace(f: Callable | Combiner, n_args: int = 1) -> Combiner:
    """Calls the closure with marker arguments and returns the resulting tree."""
    if isinstance(f, Combiner):
        return f
    out = f(Resolve("x")) if n_args == 1 else f(Resolve("x"), Resolve("y"))
    return Combiner._wrap(out)


def flatten(f: Callable | Combiner, dtype, n_args: int = 1):
    """-> (ctypes array of cb_node, n).  Literals are rounded to `dtype` first (a Rust
    closure would hold a value of type T); shared sub-trees become shared nodes."""
    dt = dtype_code(dtype)
    root = trace(f, n_args)
    order: list[Combiner] = []
    index: dict[int, int] = {}
    marker_index: dict[str, int] = {}

    def visit(node: Combiner) -> int:
        if id(node) in index:
            return index[id(node)]
        if node.op in ("x", "y") and node.op in marker_index:
            return marker_index[node.op]
        a = visit(node.lhs) if node.lhs is not None else -1
        b = visit(node.rhs) if node.rhs is not None else -1
        order.append((node, a, b))
        index[id(node)] = len(order) - 1
        if node.op in ("x", "y"):
            marker_index[node.op] = len(order) - 1
        return len(order) - 1

    visit(root)
    arr = (N.cb_node * len(order))()
    for i, (node, a, b) in enumerate(order):
        arr[i].op, arr[i].a, arr[i].b = OP[node.op], a, b
        if node.op == "const":
            if dt == N.BF16:
                arr[i].fimm = float(bf16_to_f32(bf16_from_f32([node.value]))[0])
            elif dt in FLOAT_DTYPES:
                arr[i].fimm = float(NP_DTYPE[dt](node.value))
            elif dt == N.U64:
                arr[i].iimm = int(np.uint64(node.value).astype(np.int64))
            else:
                arr[i].iimm = int(NP_DTYPE[dt](node.value))
    return arr, len(order)


class Chain:
    """A list of flattened programs in the `cb_node**` layout the ABI takes."""

    def __init__(self, fs: Sequence[Callable | Combiner], dtype, n_args: int = 1):
        self.dtype = dtype_code(dtype)
        self._keep = [flatten(f, self.dtype, n_args) for f in fs]
        self.n_progs = len(self._keep)
        self.progs = (C.POINTER(N.cb_node) * self.n_progs)(*[C.cast(a, C.POINTER(N.cb_node)) for a, _ in self._keep])
        self.n_nodes = (C.c_int32 * self.n_progs)(*[n for _, n in self._keep])


def to_cl_source(f, dtype=N.F32, marker_x: str = "x", marker_y: str = "y", n_args: int = 1) -> str:
    """`to_cl_source()` of the closure — the reference's C source rendering."""
    arr, n = flatten(f, dtype, n_args)
    buf = C.create_string_buffer(1 << 16)
    N.call("cb_expr_to_cl_source", dtype_code(dtype), arr, n, marker_x.encode(), marker_y.encode(), buf, len(buf))
    return buf.value.decode()


def ops_to_fused_src(fs, dtype=N.F32) -> str:
    """`operations_to_fused_src` (src/devices/fusing.rs:4-19)."""
    ch = Chain(fs, dtype)
    buf = C.create_string_buffer(1 << 16)
    N.call("cb_ops_to_fused_src", ch.dtype, ch.progs, ch.n_nodes, ch.n_progs, buf, len(buf))
    return buf.value.decode()


def cuda_source(fs, dtype=N.F32, kind=N.KERNEL_APPLY, n_args: int = 1) -> str:
    ch = Chain(fs, dtype, n_args)
    buf = C.create_string_buffer(1 << 18)
    N.call("cb_expr_cuda_source", ch.dtype, kind, ch.progs, ch.n_nodes, ch.n_progs, buf, len(buf))
    return buf.value.decode()


def compile_check(fs, dtype=N.F32, kind=N.KERNEL_APPLY, n_args: int = 1) -> int:
    """NVRTC-compiles the kernel for sm_100a without a device; returns the cubin size."""
    ch = Chain(fs, dtype, n_args)
    sz = C.c_size_t(0)
    N.call("cb_expr_compile_check", ch.dtype, kind, ch.progs, ch.n_nodes, ch.n_progs, C.byref(sz))
    return sz.value
