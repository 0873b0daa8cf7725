"""Cache-trace analysis (Graph module) without a device: every pure-graph unit test of the
reference (src/modules/graph/opt_graph/optimize.rs:137-395,692-805), run against BOTH the
product's OptGraph (C ABI) and the oracle's restatement."""
import ctypes as C

import pytest

from custos_b200 import _native as N
from oracle import oracle as orc


class ProductGraph:
    def __init__(self):
        self.g = C.c_void_p()
        N.call("cb_optgraph_create", C.byref(self.g))
        self.n = 0

    def add_leaf(self, length):
        i = C.c_int64()
        N.call("cb_optgraph_add_leaf", self.g, length, C.byref(i))
        self.n += 1
        return i.value

    def add_node(self, length, deps):
        i = C.c_int64()
        arr = (C.c_int64 * max(len(deps), 1))(*deps)
        N.call("cb_optgraph_add_node", self.g, length, arr, len(deps), C.byref(i))
        self.n += 1
        return i.value

    def set_skip(self, idx, skip=True):
        N.call("cb_optgraph_set_skip", self.g, idx, 1 if skip else 0)

    def is_path_optimizable(self, idx):
        v = C.c_int32()
        N.call("cb_optgraph_is_path_optimizable", self.g, idx, C.byref(v))
        return bool(v.value)

    def trace_cache_path_raw(self, idx):
        buf, w = (C.c_int64 * (self.n + 1))(), C.c_size_t()
        N.call("cb_optgraph_trace_cache_path_raw", self.g, idx, buf, self.n + 1, C.byref(w))
        return [int(buf[i]) for i in range(w.value)]

    def cache_traces(self):
        cap = 3 * self.n + 4
        buf, w = (C.c_int64 * cap)(), C.c_size_t()
        N.call("cb_optgraph_cache_traces", self.g, buf, cap, C.byref(w))
        return orc.unflatten_traces([int(buf[i]) for i in range(w.value)])


@pytest.fixture(params=["product", "oracle"])
def graph(request):
    return ProductGraph() if request.param == "product" else orc.Graph()


def test_cache_trace(graph):  # optimize.rs:137-161
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, b])
    assert graph.trace_cache_path_raw(c) == [3, 4]


def test_no_cache_trace(graph):  # optimize.rs:163-185
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, b])
    graph.add_node(10, [c, b])
    assert graph.trace_cache_path_raw(c) == []


def test_cache_trace_2(graph):  # optimize.rs:187-210
    a, b, u = graph.add_leaf(10), graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    graph.add_node(10, [a, u])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, b])
    assert graph.trace_cache_path_raw(c) == [5, 6]


def test_cache_trace_break_not_anymore(graph):  # optimize.rs:212-248
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, a])
    graph.add_node(10, [d, b])
    assert graph.is_path_optimizable(c) and not graph.is_path_optimizable(d)
    assert graph.trace_cache_path_raw(c) == [3]  # the first unoptimizable node still joins the trace


def test_trace_all(graph):  # optimize.rs:250-277
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, b])
    assert graph.cache_traces() == [(2, [3, 4])]


def test_leafed_diff_len_trace(graph):  # optimize.rs:279-307
    a = graph.add_leaf(10)
    graph.add_node(10, [a, a])
    graph.add_leaf(10)
    graph.add_leaf(10)
    c = graph.add_node(10, [a, a])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, a])
    assert graph.cache_traces()[0] == (4, [5, 6])


def test_cache_trace_neural_net(graph):  # optimize.rs:309-358
    inputs, targets = graph.add_leaf(1000), graph.add_leaf(100)
    w1, b1, w2, b2 = graph.add_leaf(640), graph.add_leaf(64), graph.add_leaf(4096), graph.add_leaf(64)
    w3, b3, w4, b4 = graph.add_leaf(4096), graph.add_leaf(64), graph.add_leaf(64), graph.add_leaf(1)
    a1 = graph.add_node(6400, [inputs, w1])
    a2 = graph.add_node(6400, [a1, b1])
    a2 = graph.add_node(6400, [a2, a2])
    a3 = graph.add_node(6400, [a2, w2])
    a4 = graph.add_node(6400, [a3, b2])
    a4 = graph.add_node(6400, [a4, a4])
    a5 = graph.add_node(6400, [a4, w3])
    a6 = graph.add_node(6400, [a5, b3])
    a6 = graph.add_node(6400, [a6, a6])
    a7 = graph.add_node(100, [a6, w4])
    a8 = graph.add_node(100, [a7, b4])
    graph.add_node(100, [a8, targets])
    assert graph.cache_traces() == [(10, [11, 12, 13, 14, 15, 16, 17, 18]), (19, [20, 21])]


def test_cache_trace_d(graph):  # optimize.rs:360-393
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [a, d])
    assert graph.trace_cache_path_raw(c) == [3, 4]
    assert graph.is_path_optimizable(c) and graph.is_path_optimizable(d)
    assert graph.cache_traces() == [(2, [3, 4])]


def test_sliced_chained_perf_example(graph):  # optimize.rs:449-495 (node structure of the retrieves)
    x, b = graph.add_leaf(1000), graph.add_leaf(1000)
    squared = graph.add_node(1000, [x, x])
    add = graph.add_node(1000, [b, x])
    mul_b = graph.add_node(1000, [add, b])
    mul = graph.add_node(1000, [squared, x])
    graph.add_node(1000, [mul, mul_b])
    assert graph.cache_traces() == [(2, [5, 6]), (3, [4])]


def test_no_cache_trace_in_graph(graph):  # optimize.rs:692-704
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    assert graph.trace_cache_path_raw(c) == []
    assert graph.cache_traces() == []


def test_multiple_traces_with_skips(graph):  # optimize.rs:706-757
    a = graph.add_leaf(10)
    _b = graph.add_node(10, [a, a])
    graph.add_leaf(10)
    _z = graph.add_leaf(10)
    c = graph.add_node(10, [a, a])
    graph.set_skip(c)
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, a])
    f = graph.add_node(10, [_b, _z])
    graph.set_skip(f)
    graph.add_node(10, [f, _z])
    assert graph.cache_traces() == [(5, [6])]


def test_multiple_traces(graph):  # optimize.rs:759-805
    a = graph.add_leaf(10)
    _b = graph.add_node(10, [a, a])
    graph.add_leaf(10)
    _z = graph.add_leaf(10)
    c = graph.add_node(10, [a, a])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, a])
    f = graph.add_node(10, [_b, _z])
    graph.add_node(10, [f, _z])
    assert graph.cache_traces() == [(1, [7, 8]), (4, [5, 6])]


def test_is_path_optimizable_doc_example(graph):  # optimize.rs:92-109
    a, b = graph.add_leaf(10), graph.add_leaf(10)
    c = graph.add_node(10, [a, b])
    d = graph.add_node(10, [c, c])
    graph.add_node(10, [d, a])
    graph.add_node(10, [d, b])
    assert graph.is_path_optimizable(c) and not graph.is_path_optimizable(d)


def test_self_dependent_node_is_a_leaf(graph):  # node.rs:14-35
    n0 = graph.add_node(10, [0, 0])
    assert not graph.is_path_optimizable(n0)
    n1 = graph.add_node(10, [0, 0])
    assert graph.is_path_optimizable(n1)
