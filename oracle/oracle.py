"""ctypes wrapper of oracle/liboracle.so — the CPU restatement of the reference's CPU device.

TEST INFRASTRUCTURE ONLY (see oracle.h): imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by custos_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"

F32, F64, F16, I32, I64, U32, U8, BF16, I8, I16, U16, U64, BOOL = range(13)
# bf16 buffers are uint16 bit patterns (numpy has no bfloat16)
NP_DTYPE = {F32: np.float32, F64: np.float64, F16: np.float16, I32: np.int32, I64: np.int64, U32: np.uint32, U8: np.uint8,
            BF16: np.uint16, I8: np.int8, I16: np.int16, U16: np.uint16, U64: np.uint64, BOOL: np.bool_}
ACC_DTYPE = {F32: np.float32, F16: np.float32, BF16: np.float32, F64: np.float64, I32: np.int64, I64: np.int64, U32: np.int64,
             U8: np.int64, I8: np.int64, I16: np.int64, U16: np.int64, U64: np.int64}


class orc_node(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("_pad", C.c_int32),
                ("fimm", C.c_double), ("iimm", C.c_int64)]


_lib = None


def build(force: bool = False) -> Path:
    src = [HERE / "oracle.c", HERE / "oracle.h"]
    if force or not LIB_PATH.exists() or any(p.stat().st_mtime > LIB_PATH.stat().st_mtime for p in src):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        nodes = C.POINTER(orc_node)
        progs = C.POINTER(C.POINTER(orc_node))
        L.orc_f32_to_f16.argtypes, L.orc_f32_to_f16.restype = [C.c_float], C.c_uint16
        L.orc_f16_to_f32.argtypes, L.orc_f16_to_f32.restype = [C.c_uint16], C.c_float
        L.orc_f32_to_bf16.argtypes, L.orc_f32_to_bf16.restype = [C.c_float], C.c_uint16
        L.orc_bf16_to_f32.argtypes, L.orc_bf16_to_f32.restype = [C.c_uint16], C.c_float
        L.orc_eval.argtypes = [i32, nodes, i32, vp, vp, vp]
        L.orc_apply_fn.argtypes = [i32, nodes, i32, vp, vp, sz]
        L.orc_apply_chain.argtypes = [i32, progs, C.POINTER(C.c_int), i32, vp, vp, sz]
        L.orc_apply_chain_mt.argtypes = [i32, progs, C.POINTER(C.c_int), i32, vp, vp, sz, i32]
        L.orc_apply_chain_boxed_f32.argtypes = [progs, C.POINTER(C.c_int), i32, vp, vp, sz]
        L.orc_apply_chain_unfused_f32.argtypes = [progs, C.POINTER(C.c_int), i32, vp, vp, sz]
        L.orc_add_unary_grad.argtypes = [i32, nodes, i32, vp, vp, vp, sz]
        L.orc_apply2.argtypes = [i32, nodes, i32, vp, vp, vp, sz]
        L.orc_binary.argtypes = [i32, i32, vp, vp, vp, sz]
        L.orc_clear.argtypes = [i32, vp, sz]
        L.orc_sum_seq.argtypes = [i32, vp, sz, vp]
        L.orc_sum_f64.argtypes, L.orc_sum_f64.restype = [i32, vp, sz], C.c_double
        L.orc_sum_two_pass.argtypes = [i32, vp, sz, i32, sz, i32, i32, i32, vp]
        L.orc_ulp_stats_f32.argtypes = [vp, vp, sz, C.POINTER(C.c_uint64), C.POINTER(sz)]
        L.orc_ulp_stats_f32.restype = C.c_uint32
        L.orc_fmaf_array.argtypes, L.orc_fmaf_array.restype = [vp, C.c_float, C.c_float, vp, sz], None
        L.orc_check_scale_add.argtypes = [C.c_uint64, C.c_uint64, C.c_float, C.c_float, i32, i32, C.POINTER(C.c_uint32)]
        L.orc_check_scale_add.restype = C.c_uint64
        L.orc_graph_new.restype = vp
        L.orc_graph_free.argtypes = [vp]
        L.orc_graph_add_leaf.argtypes, L.orc_graph_add_leaf.restype = [vp, sz], C.c_int64
        L.orc_graph_add_node.argtypes, L.orc_graph_add_node.restype = [vp, sz, C.POINTER(C.c_int64), i32], C.c_int64
        L.orc_graph_set_skip.argtypes = [vp, C.c_int64, i32]
        L.orc_graph_is_leaf.argtypes = [vp, C.c_int64]
        L.orc_graph_is_path_optimizable.argtypes = [vp, C.c_int64]
        L.orc_graph_trace_cache_path_raw.argtypes, L.orc_graph_trace_cache_path_raw.restype = [vp, C.c_int64, C.POINTER(C.c_int64), sz], sz
        L.orc_graph_cache_traces.argtypes, L.orc_graph_cache_traces.restype = [vp, C.POINTER(C.c_int64), sz], sz
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(f"oracle status {rc}")


def _nodes(f, dtype, n_args=1):
    """Uses the package's IR builder for the closure -> node array step (no arithmetic)."""
    from custos_b200.expr import flatten
    arr, n = flatten(f, dtype, n_args)
    return C.cast(arr, C.POINTER(orc_node)), n, arr


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as(dtype, a):
    return np.ascontiguousarray(a, dtype=NP_DTYPE[dtype])


def eval_scalar(f, dtype, x, y=None):
    p, n, keep = _nodes(f, dtype, 1 if y is None else 2)
    xa, out = _as(dtype, [x]), np.zeros(1, NP_DTYPE[dtype])
    ya = _as(dtype, [y]) if y is not None else None
    _check(lib().orc_eval(dtype, p, n, _ptr(xa), _ptr(ya) if ya is not None else None, _ptr(out)))
    return out[0]


def apply_fn(f, dtype, x):
    p, n, keep = _nodes(f, dtype)
    x = _as(dtype, x)
    out = np.empty_like(x)
    _check(lib().orc_apply_fn(dtype, p, n, _ptr(x), _ptr(out), x.size))
    return out


def _chain(fs, dtype):
    from custos_b200.expr import Chain
    ch = Chain(fs, dtype)
    progs = C.cast(ch.progs, C.POINTER(C.POINTER(orc_node)))
    n_nodes = C.cast(ch.n_nodes, C.POINTER(C.c_int))
    return ch, progs, n_nodes


def apply_chain(fs, dtype, x, threads: int = 1):
    ch, progs, n_nodes = _chain(fs, dtype)
    x = _as(dtype, x)
    out = np.empty_like(x)
    if threads <= 1:
        _check(lib().orc_apply_chain(dtype, progs, n_nodes, ch.n_progs, _ptr(x), _ptr(out), x.size))
    else:
        _check(lib().orc_apply_chain_mt(dtype, progs, n_nodes, ch.n_progs, _ptr(x), _ptr(out), x.size, threads))
    return out


def apply_chain_boxed(fs, x):
    """f32 only: the reference's fused CPU path as it really runs — per element and per op a heap-boxed dyn op is
    built, evaluated through a function pointer and dropped (cpu_device.rs:217-229, op_hint.rs:30-33)."""
    ch, progs, n_nodes = _chain(fs, F32)
    x = _as(F32, x)
    out = np.empty_like(x)
    _check(lib().orc_apply_chain_boxed_f32(progs, n_nodes, ch.n_progs, _ptr(x), _ptr(out), x.size))
    return out


def apply_chain_unfused(fs, x):
    """f32 only: one monomorphised apply_fn_slice loop per recorded op (cpu_stack_ops.rs:7-15), n_progs passes."""
    ch, progs, n_nodes = _chain(fs, F32)
    x = _as(F32, x)
    out = np.empty_like(x)
    _check(lib().orc_apply_chain_unfused_f32(progs, n_nodes, ch.n_progs, _ptr(x), _ptr(out), x.size))
    return out


def add_unary_grad(f, dtype, lhs, lhs_grad, out_grad):
    """Returns the updated lhs_grad (the input array is not modified)."""
    p, n, keep = _nodes(f, dtype)
    lhs, out_grad = _as(dtype, lhs), _as(dtype, out_grad)
    g = _as(dtype, lhs_grad).copy()
    _check(lib().orc_add_unary_grad(dtype, p, n, _ptr(lhs), _ptr(out_grad), _ptr(g), lhs.size))
    return g


def apply2(f, dtype, lhs, rhs):
    p, n, keep = _nodes(f, dtype, 2)
    lhs, rhs = _as(dtype, lhs), _as(dtype, rhs)
    out = np.empty_like(lhs)
    _check(lib().orc_apply2(dtype, p, n, _ptr(lhs), _ptr(rhs), _ptr(out), lhs.size))
    return out


def binary(op: int, dtype, lhs, rhs):
    lhs, rhs = _as(dtype, lhs), _as(dtype, rhs)
    out = np.empty_like(lhs)
    _check(lib().orc_binary(dtype, op, _ptr(lhs), _ptr(rhs), _ptr(out), lhs.size))
    return out


def clear(dtype, buf: np.ndarray):
    _check(lib().orc_clear(dtype, _ptr(buf), buf.size))
    return buf


def sum_seq(dtype, x):
    x = _as(dtype, x)
    out = np.zeros(1, ACC_DTYPE[dtype])
    _check(lib().orc_sum_seq(dtype, _ptr(x), x.size, _ptr(out)))
    return out[0]


def sum_f64(dtype, x) -> float:
    x = _as(dtype, x)
    return float(lib().orc_sum_f64(dtype, _ptr(x), x.size))


def sum_two_pass(dtype, x, blocks, chunk, threads, vec, threads2):
    x = _as(dtype, x)
    out = np.zeros(1, ACC_DTYPE[dtype])
    _check(lib().orc_sum_two_pass(dtype, _ptr(x), x.size, blocks, chunk, threads, vec, threads2, _ptr(out)))
    return out[0]


def ulp_stats_f32(got: np.ndarray, want: np.ndarray):
    """-> (max ulp, index of its first occurrence, histogram of distances 0, 1, 2, 3, 4, > 4)"""
    got, want = np.ascontiguousarray(got, np.float32), np.ascontiguousarray(want, np.float32)
    hist = (C.c_uint64 * 6)()
    where = C.c_size_t()
    worst = lib().orc_ulp_stats_f32(_ptr(got), _ptr(want), got.size, hist, C.byref(where))
    return int(worst), int(where.value), [int(h) for h in hist]


def fmaf_array(a: np.ndarray, b: float, c: float) -> np.ndarray:
    """fmaf(a[i], b, c) with glibc's correctly rounded fmaf"""
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty_like(a)
    lib().orc_fmaf_array(_ptr(a), b, c, _ptr(out), a.size)
    return out


def check_scale_add(P: float, Cc: float, mode: int, first: int = 0, count: int = 1 << 32, threads: int = 0):
    """Mismatches between two rounded steps and one fmaf over f32 bit patterns [first, first + count); mode 0:
    (u * P) + C, mode 1: (u + C) * P.  -> (number of mismatches, first offending bit pattern)"""
    import os
    first_bad = C.c_uint32(0)
    bad = lib().orc_check_scale_add(first, count, P, Cc, mode, threads or (os.cpu_count() or 1), C.byref(first_bad))
    return int(bad), int(first_bad.value)


def f32_to_f16_bits(v: float) -> int:
    return int(lib().orc_f32_to_f16(C.c_float(v)))


def f16_bits_to_f32(h: int) -> float:
    return float(lib().orc_f16_to_f32(C.c_uint16(h)))


def f32_to_bf16_bits(v: float) -> int:
    return int(lib().orc_f32_to_bf16(C.c_float(v)))


def bf16_bits_to_f32(h: int) -> float:
    return float(lib().orc_bf16_to_f32(C.c_uint16(h)))


class Graph:
    """OptGraph restated (src/modules/graph/opt_graph.rs, opt_graph/optimize.rs)."""

    def __init__(self):
        self.g = lib().orc_graph_new()
        self.n = 0

    def __del__(self):
        try:
            lib().orc_graph_free(self.g)
        except Exception:
            pass

    def add_leaf(self, length: int) -> int:
        self.n += 1
        return int(lib().orc_graph_add_leaf(self.g, length))

    def add_node(self, length: int, deps) -> int:
        arr = (C.c_int64 * len(deps))(*deps)
        self.n += 1
        return int(lib().orc_graph_add_node(self.g, length, arr, len(deps)))

    def set_skip(self, idx: int, skip: bool = True):
        lib().orc_graph_set_skip(self.g, idx, 1 if skip else 0)

    def is_path_optimizable(self, idx: int) -> bool:
        return bool(lib().orc_graph_is_path_optimizable(self.g, idx))

    def trace_cache_path_raw(self, idx: int):
        buf = (C.c_int64 * (self.n + 1))()
        k = lib().orc_graph_trace_cache_path_raw(self.g, idx, buf, self.n + 1)
        return [int(buf[i]) for i in range(k)]

    def cache_traces(self):
        cap = 3 * self.n + 4
        buf = (C.c_int64 * cap)()
        k = lib().orc_graph_cache_traces(self.g, buf, cap)
        return unflatten_traces([int(buf[i]) for i in range(k)])


def unflatten_traces(flat):
    out, i = [], 0
    while i < len(flat):
        idx, k = flat[i], flat[i + 1]
        out.append((idx, list(flat[i + 2:i + 2 + k])))
        i += 2 + k
    return out
