"""The C ABI takes untrusted input at the boundary: malformed expression IR, null pointers, undersized buffers, bad enum
values.  Every such call must come back with an error code — never crash, never write out of bounds.  Runs on the CPU
(no device needed for these entry points) and, through scripts/host_sanitize.sh, under ASan + UBSan."""
import ctypes as C
import random

import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import expr as E


def random_nodes(rng: random.Random, n: int, wild: bool):
    arr = (N.cb_node * max(n, 1))()
    for i in range(n):
        if wild:
            arr[i].op = rng.randint(-5, 30)
            arr[i].a = rng.randint(-3, n + 3)
            arr[i].b = rng.randint(-3, n + 3)
        else:  # mostly well formed: operands before users, unused slots -1; now and then one field is off
            op = rng.randint(0, 21) if i else rng.choice([0, 1, 2])
            name = E.OPS[op]
            arr[i].op = op
            arr[i].a = rng.randint(0, i - 1) if (name in E.BINARY or name in E.UNARY) and i else -1
            arr[i].b = rng.randint(0, i - 1) if name in E.BINARY and i else -1
            if rng.random() < 0.08:
                setattr(arr[i], rng.choice(["a", "b"]), rng.randint(-2, n + 1))
        arr[i].fimm = rng.choice([0.0, 1.5, -2.0, 0.25, 1e300, float("nan"), float("inf"), 3.1])
        arr[i].iimm = rng.choice([0, 1, -1, 2 ** 63 - 1, -2 ** 63, 255, 70000])
    return arr


def is_valid(arr, n, dtype):
    """Independent restatement of what a well-formed program is (what expr.cpp must accept, and nothing else)."""
    is_float = dtype in E.FLOAT_DTYPES
    signed = dtype in (N.I8, N.I16, N.I32, N.I64)
    if n <= 0 or n > 64 or dtype == N.BOOL:
        return False
    for i in range(n):
        op, a, b = arr[i].op, arr[i].a, arr[i].b
        if not 0 <= op <= 21:
            return False
        name = E.OPS[op]
        if not is_float and name not in ("x", "y", "const", "add", "mul", "sub", "div", "geq", "leq", "eq") and not (name == "neg" and signed):
            return False
        if name in E.BINARY and not (0 <= a < i and 0 <= b < i):
            return False
        if name in E.UNARY and not (0 <= a < i and b < 0):
            return False
        if name in ("x", "y", "const") and not (a < 0 and b < 0):
            return False
        if name == "const" and is_float and np.isfinite(arr[i].fimm):
            f = arr[i].fimm
            if abs(f) > 3e38 and dtype != N.F64:
                return False  # not representable in any of the narrower float types
            if dtype == N.F32 and float(np.float32(f)) != f:
                return False
            if dtype == N.F16 and (abs(f) > 65504 or float(np.float16(f)) != f):
                return False
            if dtype == N.BF16 and float(E.bf16_to_f32(E.bf16_from_f32([f]))[0]) != f:
                return False
    return True


@pytest.mark.parametrize("seed", range(6))
def test_malformed_ir_is_rejected_not_crashed_on(seed):
    lib = N.load()
    rng = random.Random(seed)
    buf = C.create_string_buffer(1 << 16)
    accepted = rejected = 0
    for _ in range(1500):
        n = rng.randint(0, 12)
        dtype = rng.choice(list(range(13)) + [-1, 13, 99])
        arr = random_nodes(rng, n, wild=rng.random() < 0.5)
        rc = lib.cb_expr_to_cl_source(dtype, arr, n, b"x", b"y", buf, len(buf))
        progs = (C.POINTER(N.cb_node) * 1)(C.cast(arr, C.POINTER(N.cb_node)))
        counts = (C.c_int32 * 1)(n)
        rc2 = lib.cb_expr_cuda_source(dtype, rng.randint(-1, 3), progs, counts, 1, buf, len(buf))
        rc3 = lib.cb_ops_to_fused_src(dtype, progs, counts, 1, buf, len(buf))
        ok = 0 <= dtype < 13 and is_valid(arr, n, dtype)
        if ok:
            accepted += 1
            has_y = any(arr[i].op == 1 for i in range(n))  # a unary chain cannot hold the second marker
            assert rc == N.CB_OK and (rc3 == N.CB_OK) == (not has_y), (dtype, n, [(arr[i].op, arr[i].a, arr[i].b) for i in range(n)])
        else:
            rejected += 1
            assert rc != N.CB_OK and rc3 != N.CB_OK, (dtype, n, [(arr[i].op, arr[i].a, arr[i].b) for i in range(n)])
            assert rc2 != N.CB_OK
    assert accepted > 20 and rejected > 500


def test_small_output_buffers_and_nulls():
    lib = N.load()
    arr, n = E.flatten(lambda x: x.mul(2.0).add(1.0).sin(), N.F32)
    for cap in (0, 1, 5, 20):
        buf = C.create_string_buffer(max(cap, 1) + 8)
        buf.raw = b"\xAA" * len(buf)
        assert lib.cb_expr_to_cl_source(N.F32, arr, n, b"x", b"y", buf, cap) != N.CB_OK
        assert buf.raw[max(cap, 1):] == b"\xAA" * 8 or cap == 0  # nothing written past `cap`
    assert lib.cb_expr_to_cl_source(N.F32, None, 3, b"x", b"y", C.create_string_buffer(64), 64) != N.CB_OK
    assert lib.cb_expr_to_cl_source(N.F32, arr, n, b"x", b"y", None, 64) != N.CB_OK
    assert lib.cb_expr_compile_check(N.F32, N.KERNEL_APPLY, None, None, 1, None) != N.CB_OK
    progs = (C.POINTER(N.cb_node) * 1)(C.cast(arr, C.POINTER(N.cb_node)))
    assert lib.cb_expr_compile_check(N.F32, N.KERNEL_APPLY, progs, (C.c_int32 * 1)(n), 0, None) != N.CB_OK
    assert lib.cb_expr_compile_check(N.F32, N.KERNEL_APPLY, progs, (C.c_int32 * 1)(n), 65, None) != N.CB_OK
    assert lib.cb_expr_compile_check(N.F32, 7, progs, (C.c_int32 * 1)(n), 1, None) != N.CB_OK
    # handles: every device-level call with a null device is an argument error, not a segfault
    for name, args in (("cb_sync", [None]), ("cb_free", [None, 0]), ("cb_cache_clear", [None]),
                       ("cb_clear", [None, N.F32, 0, 4]), ("cb_apply", [None, None, 0, 0, 4]),
                       ("cb_binary", [None, N.F32, 0, 0, 0, 0, 4]), ("cb_sum", [None, N.F32, 0, 4, 0]),
                       ("cb_graph_begin", [None]), ("cbm_run", [None]), ("cbm_buffer_drop", [None, 1]),
                       ("cbm_unary_fusing", [None]), ("cbm_backward", [None, 1])):
        assert getattr(lib, name)(*args) != N.CB_OK, name
    sz = C.c_size_t()
    assert lib.cb_serde_encode(N.F32, 5, None, 0, None, 0, C.byref(sz)) != N.CB_OK          # unknown format
    assert lib.cb_serde_decode(N.I32, N.SER_BINCODE, b"\x01", 1, None, 0, C.byref(sz)) == N.CB_ERR_PARSE
    assert lib.cb_shard_range(10, 4, 0, 0, C.byref(sz), C.byref(sz)) != N.CB_OK
    assert lib.cb_shard_range(10, 4, 2, 5, C.byref(sz), C.byref(sz)) != N.CB_OK


def test_optgraph_api_with_garbage_arguments():
    """Out-of-range node indices, negative counts, null / undersized output arrays: error codes, no memory errors."""
    lib = N.load()
    rng = random.Random(1)
    for _ in range(200):
        g = C.c_void_p()
        assert lib.cb_optgraph_create(C.byref(g)) == N.CB_OK
        n = 0
        for _ in range(rng.randint(0, 12)):
            idx = C.c_int64()
            if rng.random() < 0.3:
                rc = lib.cb_optgraph_add_leaf(g, rng.choice([0, 1, 10, 2 ** 40]), C.byref(idx))
            else:
                k = rng.randint(0, 4)
                deps = (C.c_int64 * max(k, 1))(*[rng.randint(-3, n + 3) for _ in range(k)])
                rc = lib.cb_optgraph_add_node(g, rng.choice([0, 1, 10]), deps, rng.choice([k, k, -1, 0]), C.byref(idx))
            if rc == N.CB_OK:
                assert idx.value == n
                n += 1
        out, w, flag = (C.c_int64 * 256)(), C.c_size_t(), C.c_int32()
        for q in range(-2, n + 3):
            rc = lib.cb_optgraph_is_path_optimizable(g, q, C.byref(flag))
            assert (rc == N.CB_OK) == (0 <= q < n)
            rc = lib.cb_optgraph_trace_cache_path_raw(g, q, out, rng.choice([0, 1, 256]), C.byref(w))
            assert rc == N.CB_OK or not 0 <= q < n or w.value > 0
            assert (lib.cb_optgraph_set_skip(g, q, 1) == N.CB_OK) == (0 <= q < n)
        lib.cb_optgraph_cache_traces(g, out, rng.choice([0, 3, 256]), C.byref(w))
        lib.cb_optgraph_cache_traces(g, None, 0, C.byref(w))
        assert lib.cb_optgraph_destroy(g) == N.CB_OK
    assert lib.cb_optgraph_destroy(None) in (N.CB_OK, N.CB_ERR_INVALID_ARG)
