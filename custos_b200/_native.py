"""ctypes binding of libcustos_b200.so — the C ABI declared in include/custos_b200.h.

Only plumbing lives here: every compute call goes straight to the CUDA library.  There is
no Python/NumPy/torch fallback: if the library is missing, or no GPU is usable, the calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libcustos_b200.so"


class CustosError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[custos_b200 status {code}] {msg}")
        self.code = code
        self.msg = msg


class cb_node(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("_pad", C.c_int32),
                ("fimm", C.c_double), ("iimm", C.c_int64)]


# status codes / enums (include/custos_b200.h)
CB_OK = 0
CB_ERR_INVALID_ARG, CB_ERR_ZERO_LENGTH, CB_ERR_NO_DEVICE, CB_ERR_UNSUPPORTED, CB_ERR_EXPR = 1, 2, 3, 4, 5
CB_ERR_INVALID_LAZY_BUF, CB_ERR_MISSING_CACHE_TRACES, CB_ERR_GRAPH_OPTIMIZATION = 6, 7, 8
CB_ERR_SHAPE, CB_ERR_STATE, CB_ERR_TYPE_MISMATCH, CB_ERR_PARSE = 9, 10, 11, 12
SER_JSON, SER_BINCODE = 0, 1
F32, F64, F16, I32, I64, U32, U8, BF16, I8, I16, U16, U64, BOOL = range(13)
KERNEL_APPLY, KERNEL_UNARY_GRAD, KERNEL_BINARY, KERNEL_CHAIN_GRAD = 0, 1, 2, 3
GRAD_SEED_ONES = 1
BIN_ADD, BIN_MUL, BIN_SUB, BIN_DIV = 0, 1, 2, 3
CBM_BASE, CBM_CACHED, CBM_LAZY, CBM_GRAPH, CBM_AUTOGRAD = 0, 1, 2, 4, 8
COMM_ID_BYTES = 128

_vp, _i32, _u32, _i64, _u64, _sz, _dbl = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64, C.c_size_t, C.c_double
_P = C.POINTER
_nodes, _cstr = _P(cb_node), C.c_char_p
_progs = _P(_P(cb_node))

# name -> argtypes; every function returns int32 status unless listed in _OTHER_RESTYPE
SIGNATURES = {
    "cb_expr_to_cl_source": [_i32, _nodes, _i32, _cstr, _cstr, C.c_char_p, _sz],
    "cb_ops_to_fused_src": [_i32, _progs, _P(_i32), _i32, C.c_char_p, _sz],
    "cb_expr_cuda_source": [_i32, _i32, _progs, _P(_i32), _i32, C.c_char_p, _sz],
    "cb_expr_compile_check": [_i32, _i32, _progs, _P(_i32), _i32, _P(_sz)],
    "cb_device_create": [_i32, _P(_vp)],
    "cb_device_destroy": [_vp],
    "cb_device_ordinal": [_vp, _P(_i32)],
    "cb_device_sm_count": [_vp, _P(_i32)],
    "cb_device_stream": [_vp, _P(_vp)],
    "cb_sync": [_vp],
    "cb_alloc": [_vp, _sz, _i32, _P(_u64)],
    "cb_free": [_vp, _u64],
    "cb_mem_info": [_vp, _P(_sz), _P(_sz)],
    "cb_cache_retrieve": [_vp, _u64, _sz, _P(_u64), _P(_i32)],
    "cb_cache_clear": [_vp],
    "cb_h2d": [_vp, _u64, _vp, _sz],
    "cb_d2h": [_vp, _vp, _u64, _sz],
    "cb_host_alloc": [_sz, _P(_vp)],
    "cb_host_alloc_ex": [_sz, _u32, _P(_vp)],
    "cb_host_free": [_vp],
    "cb_h2d_async": [_vp, _u64, _vp, _sz],
    "cb_d2h_async": [_vp, _vp, _u64, _sz],
    "cb_copy": [_vp, _i32, _u64, _sz, _u64, _sz, _sz],
    "cb_clear": [_vp, _i32, _u64, _sz],
    "cb_fill": [_vp, _i32, _u64, _sz, _dbl, _i64],
    "cb_expr_compile": [_vp, _i32, _i32, _progs, _P(_i32), _i32, _P(_vp)],
    "cb_expr_release": [_vp],
    "cb_expr_set_lookup": [_vp, _i32],
    "cb_expr_has_lookup": [_vp, _P(_i32)],
    "cb_apply": [_vp, _vp, _u64, _u64, _sz],
    "cb_unary_grad": [_vp, _vp, _u64, _u64, _u64, _sz],
    "cb_unary_grad_ex": [_vp, _vp, _u64, _u64, _u64, _sz, _u32],
    "cb_apply2": [_vp, _vp, _u64, _u64, _u64, _sz],
    "cb_apply_host": [_vp, _vp, _vp, _vp, _sz],
    "cb_binary": [_vp, _i32, _i32, _u64, _u64, _u64, _sz],
    "cb_sum": [_vp, _i32, _u64, _sz, _u64],
    "cb_mean": [_vp, _i32, _u64, _sz, _u64],
    "cb_sum_host": [_vp, _i32, _u64, _sz, _vp],
    "cb_mean_host": [_vp, _i32, _u64, _sz, _vp],
    "cb_sum_plan": [_i32, _sz, _P(_i32), _P(_sz), _P(_i32), _P(_i32), _P(_i32)],
    "cb_graph_begin": [_vp],
    "cb_graph_end": [_vp, _P(_vp)],
    "cb_graph_launch": [_vp, _vp],
    "cb_graph_destroy": [_vp],
    "cb_graph_node_count": [_vp, _P(_sz)],
    "cb_launch_count": [_vp, _P(_u64)],
    "cb_event_create": [_vp, _P(_vp)],
    "cb_event_record": [_vp, _vp],
    "cb_event_sync": [_vp],
    "cb_event_elapsed_ms": [_vp, _vp, _P(C.c_float)],
    "cb_event_destroy": [_vp],
    "cb_comm_unique_id": [C.c_char_p],
    "cb_comm_create": [_vp, _i32, _i32, C.c_char_p, _P(_vp)],
    "cb_comm_destroy": [_vp],
    "cb_comm_uses_peer_memory": [_vp, _P(_i32)],
    "cb_comm_sum": [_vp, _i32, _u64, _sz, _u64],
    "cb_comm_mean": [_vp, _i32, _u64, _sz, _sz, _u64],
    "cb_comm_sum_host": [_vp, _i32, _u64, _sz, _vp],
    "cb_comm_mean_host": [_vp, _i32, _u64, _sz, _sz, _vp],
    "cb_comm_check": [_vp],
    "cb_comm_rank": [_vp, _P(_i32), _P(_i32)],
    "cb_comm_device": [_vp, _P(_vp)],
    "cb_shard_range": [_sz, _i32, _i32, _i32, _P(_sz), _P(_sz)],
    # module layer
    "cbm_device_create": [_i32, _u32, _i32, _P(_vp)],
    "cbm_device_destroy": [_vp],
    "cbm_device_raw": [_vp, _P(_vp)],
    "cbm_device_set_comm": [_vp, _vp],
    "cbm_buffer_new_sharded": [_vp, _i32, _sz, _P(_u64)],
    "cbm_buffer_from_host_sharded": [_vp, _i32, _vp, _sz, _P(_u64)],
    "cbm_buffer_shard": [_vp, _u64, _P(_sz), _P(_sz), _P(_sz)],
    "cbm_buffer_new": [_vp, _i32, _sz, _P(_u64)],
    "cbm_buffer_from_host": [_vp, _i32, _vp, _sz, _P(_u64)],
    "cbm_buffer_drop": [_vp, _u64],
    "cbm_buffer_len": [_vp, _u64, _P(_sz)],
    "cbm_buffer_read": [_vp, _u64, _vp, _sz],
    "cbm_buffer_write": [_vp, _u64, _vp, _sz],
    "cbm_buffer_ptr": [_vp, _u64, _P(_u64)],
    "cbm_buffer_id": [_vp, _u64, _P(_u64)],
    "cbm_buffer_require_grad": [_vp, _u64],
    "cbm_buffer_requires_grad": [_vp, _u64, _P(_i32)],
    "cbm_buffer_checkpoint": [_vp, _u64],
    "cbm_retrieve": [_vp, _i32, _sz, _P(_u64), _i32, _P(_u64)],
    "cbm_apply_fn": [_vp, _u64, _nodes, _i32, _P(_u64)],
    "cbm_add_unary_grad": [_vp, _u64, _u64, _u64, _nodes, _i32],
    "cbm_unary_ew": [_vp, _u64, _nodes, _i32, _nodes, _i32, _P(_u64)],
    "cbm_binary": [_vp, _i32, _u64, _u64, _P(_u64)],
    "cbm_binary_into": [_vp, _i32, _u64, _u64, _u64],
    "cbm_clear": [_vp, _u64],
    "cbm_clear_op": [_vp, _u64],
    "cbm_copy_slice": [_vp, _u64, _sz, _u64, _sz, _sz],
    "cbm_clone_buf": [_vp, _u64, _P(_u64)],
    "cbm_sum": [_vp, _u64, _vp],
    "cbm_mean": [_vp, _u64, _vp],
    "cbm_run": [_vp],
    "cbm_exec_now": [_vp, _sz, _sz],
    "cbm_exec_last_n": [_vp, _sz],
    "cbm_ops_count": [_vp, _P(_sz)],
    "cbm_alloc_later": [_vp],
    "cbm_set_lazy_enabled": [_vp, _i32],
    "cbm_op_hint_src": [_vp, _sz, C.c_char_p, _sz],
    "cbm_op_expr": [_vp, _sz, _P(_vp)],
    "cbm_set_graph_replay": [_vp, _i32],
    "cbm_replay_kernel_nodes": [_vp, _P(_sz)],
    "cbm_optimize_mem_graph": [_vp],
    "cbm_unary_fusing": [_vp],
    "cbm_elementwise_fusing": [_vp],
    "cbm_cache_traces": [_vp, _P(_i64), _sz, _P(_sz)],
    "cbm_cursor": [_vp, _P(_u64)],
    "cbm_set_cursor": [_vp, _u64],
    "cbm_backward": [_vp, _u64],
    "cbm_backward_with": [_vp, _u64, _vp, _sz],
    "cbm_grad": [_vp, _u64, _P(_u64)],
    "cbm_zero_grad": [_vp],
    "cbm_set_grad_enabled": [_vp, _i32],
    # untyped views and serde
    "cbm_buffer_dtype": [_vp, _u64, _P(_i32)],
    "cbm_buffer_matches_type": [_vp, _u64, _i32],
    "cbm_buffer_read_typed": [_vp, _u64, _i32, _vp, _sz],
    "cb_serde_encode": [_i32, _i32, _vp, _sz, _vp, _sz, _P(_sz)],
    "cb_serde_decode": [_i32, _i32, _vp, _sz, _vp, _sz, _P(_sz)],
    "cbm_buffer_serialize": [_vp, _u64, _i32, _vp, _sz, _P(_sz)],
    "cbm_buffer_deserialize": [_vp, _i32, _i32, _vp, _sz, _P(_u64)],
    # device-free graph analysis
    "cb_optgraph_create": [_P(_vp)],
    "cb_optgraph_destroy": [_vp],
    "cb_optgraph_add_leaf": [_vp, _sz, _P(_i64)],
    "cb_optgraph_add_node": [_vp, _sz, _P(_i64), _i32, _P(_i64)],
    "cb_optgraph_set_skip": [_vp, _i64, _i32],
    "cb_optgraph_is_path_optimizable": [_vp, _i64, _P(_i32)],
    "cb_optgraph_trace_cache_path_raw": [_vp, _i64, _P(_i64), _sz, _P(_sz)],
    "cb_optgraph_cache_traces": [_vp, _P(_i64), _sz, _P(_sz)],
}
_OTHER_RESTYPE = {"cb_last_error": ([], C.c_char_p), "cb_abi_version": ([], _i32), "cb_dtype_size": ([_i32], _sz),
                   "cbm_untyped_supports": ([_i32], _i32)}

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (building it is the job of `custos_b200.build`).  Raises if
    it is missing — the product path never degrades to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("CUSTOS_B200_LIB", LIB_PATH))
    if not path.exists():
        raise CustosError(-1, f"{path} is missing: run `python -m custos_b200.build` (needs nvcc); "
                              "there is no fallback implementation")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export the declared ABI
        fn.argtypes = argtypes
        fn.restype = _i32
    for name, (argtypes, restype) in _OTHER_RESTYPE.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != CB_OK:
        raise CustosError(rc, load().cb_last_error().decode("utf-8", "replace"))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
