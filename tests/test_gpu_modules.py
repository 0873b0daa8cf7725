"""GPU tests of the module layer (Base, Cached, Lazy, Graph, Autograd) through the C ABI.

Each test restates a test of the reference (cited) on the CUDA device and, where numbers are
involved, checks them against the CPU oracle with the tight bars of BASELINE.json rather than
the reference's loose `roughly_eq_slices` (abs 0.1).
"""
import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import CustosError
from custos_b200.device import CUDA
from oracle import oracle as orc
from tests.helpers import CHAIN8, CHAIN8_GRADS, CHEAP8, assert_bit_exact, assert_ulp, random_inputs

pytestmark = pytest.mark.gpu

SIN = [0.8414709848078965, 0.9092974268256817, 0.1411200080598672, -0.7568024953079282]
COS = [0.5403023058681398, -0.4161468365471424, -0.9899924966004454, -0.6536436208636119]


# ------------------------------------------------------------------ Base
def test_base_apply_fn_is_eager():
    # src/unary.rs:12-18 (doc test) on CUDA<Base>
    with CUDA("Base") as dev:
        a = dev.buffer([1., 2., 3., 3., 2., 1.])
        out = dev.apply_fn(a, lambda x: x.mul(2.))
        assert out.read().tolist() == [2., 4., 6., 6., 4., 2.]
        assert dev.ops_count() == 0


def test_base_add_unary_grad():
    # src/unary.rs:36-47 (doc test)
    with CUDA("Base") as dev:
        a = dev.buffer([1., 2., 3., 3., 2., 1.])
        out = dev.apply_fn(a, lambda x: x.mul(2.))
        a_grad = a.empty_like()
        out_grad = dev.buffer([1.] * 6)
        dev.add_unary_grad(a, a_grad, out_grad, lambda x: 2.0)
        assert a_grad.read().tolist() == [2.] * 6
        assert len(out) == 6


def test_clear_write_clone_copy_slice():
    # src/devices/cuda/ops.rs:244-249, tests/clear.rs:31-43, tests/write.rs:51-87, tests/clone_buf.rs:38-49
    with CUDA("Base") as dev:
        x = dev.buffer(np.array([1, 2, 3, 4, 5, 6], np.uint32))
        x.clear()
        assert x.read().tolist() == [0] * 6
        b = dev.new_buffer(np.float32, 5)
        b.write([1., 2., 3., 4., 5.])
        assert b.read().tolist() == [1., 2., 3., 4., 5.]
        c = b.clone()
        assert c.read().tolist() == [1., 2., 3., 4., 5.] and c.ptr() != b.ptr()
        dst = dev.new_buffer(np.float32, 5)
        dev.write_buf(dst, b)
        assert dst.read().tolist() == [1., 2., 3., 4., 5.]
        # src/op_traits.rs:34-60 doc example
        dest = dev.new_buffer(np.float32, 6)
        dev.copy_slice_to(b, range(1, 3), dest, range(3, 5))
        assert dest.read().tolist() == [0., 0., 0., 2., 3., 0.]
        # src/op_traits.rs:62-93 doc example (copy_slice_all)
        dest2 = dev.new_buffer(np.float32, 6)
        dev.copy_slice_all(b, dest2, [(range(2, 4), range(4, 6)), (range(0, 2), range(0, 2))])
        assert dest2.read().tolist() == [1., 2., 0., 0., 3., 4.]


def test_zero_length_buffer_is_an_error():
    with CUDA("Base") as dev:
        with pytest.raises(CustosError) as ei:
            dev.new_buffer(np.float32, 0)
        assert ei.value.code == N.CB_ERR_ZERO_LENGTH


def test_binary_ops_on_base():
    # tests/demo_impl/cuda/mod.rs:40-66
    with CUDA("Base") as dev:
        n = 655_360
        lhs, rhs = dev.buffer(np.full(n, 1, np.float32)), dev.buffer(np.full(n, 4, np.float32))
        for _ in range(3):
            out = dev.add(lhs, rhs)
            assert np.all(out.read() == 5.0)
            out.drop()
        # README.md:96-122
        assert dev.mul(dev.buffer([1., 2., 3.]), dev.buffer([4., 5., 6.])).read().tolist() == [4., 10., 18.]


# ------------------------------------------------------------------ Lazy
def test_lazy_retrieve_ids():
    # src/modules/lazy.rs:553-566
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        buf = dev.new_buffer(np.int32, 10)
        assert buf.ptr() != 0
        x = dev.retrieve(10, (), np.int32)
        assert x.id() == 0 and x.ptr() == 0
        y = dev.retrieve(10, (), np.int32)
        assert y.id() == 1


def test_lazy_apply_fn_nothing_before_run():
    # src/modules/lazy.rs:642-655, :657-672 (alloc_later shows zeros, run shows results)
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        buf = dev.new_buffer(np.int32, 10)
        out = dev.apply_fn(buf, lambda x: x.add(3))
        with pytest.raises(CustosError):
            out.read()  # no storage yet
        dev.alloc_later()
        assert out.replace().read().tolist() == [0] * 10
        dev.run()
        assert out.replace().read().tolist() == [3] * 10


def test_lazy_dropped_buffer_gives_invalid_lazy_buf():
    # src/modules/lazy.rs:624-640
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        buf = dev.new_buffer(np.int32, 10)
        out = dev.apply_fn(buf, lambda x: x.add(3))
        out.drop()
        buf.drop()
        with pytest.raises(CustosError) as ei:
            dev.run()
        assert ei.value.code == N.CB_ERR_INVALID_LAZY_BUF


def test_lazy_add_apply_fn_with_run():
    # src/modules/lazy.rs:687-709
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        buf = dev.new_buffer(np.int32, 10)
        lhs = dev.apply_fn(buf, lambda x: x.add(3))
        rhs = dev.buffer(np.arange(1, 11, dtype=np.int32))
        assert rhs.read().tolist() == list(range(1, 11))
        out = dev.add(lhs, rhs)
        dev.run()
        assert lhs.replace().read().tolist() == [3] * 10
        assert out.replace().read().tolist() == [4, 5, 6, 7, 8, 9, 10, 11, 12, 13]


def test_lazy_unary_grad_kat():
    # src/devices/cuda/ops.rs:279-294
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        lhs = dev.buffer(np.array([1, 2, 3, 4, 5, 6], np.int32))
        lhs_grad = dev.buffer(np.array([1, 2, 3, 4, 5, 6], np.int32))
        out = dev.buffer(np.array([1, 1, 1, 1, 1, 1], np.int32))
        dev.add_unary_grad(lhs, lhs_grad, out, lambda x: x.add(2))
        assert lhs_grad.read().tolist() == [1, 2, 3, 4, 5, 6]
        dev.run()
        assert lhs_grad.read().tolist() == [4, 6, 8, 10, 12, 14]


def test_lazy_exec_now_and_exec_last_n():
    # src/modules/lazy.rs:732-789: executed ops are drained, run() replays the rest
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        a = dev.buffer(np.array([1, 2, 3, 4], np.int32))
        b = dev.buffer(np.array([1, 2, 3, 4], np.int32))
        zeros = dev.apply_fn(a, lambda x: x.mul(0))      # op 0
        out = dev.add(a, b)                              # op 1
        assert dev.ops_count() == 2
        dev.exec_last_n(1)
        assert dev.ops_count() == 1
        assert out.replace().read().tolist() == [2, 4, 6, 8]
        assert zeros.replace().read().tolist() == [0, 0, 0, 0]  # allocated (zeroed) but op 0 not run... and its result is 0 anyway
        dev.run()
        assert dev.ops_count() == 1
        dev.exec_now(0, None)
        assert dev.ops_count() == 0


def test_op_hints_are_recorded():
    # src/op_hint.rs:44-83
    with CUDA("Lazy", "Base") as dev:
        buf = dev.buffer([1., 2., 3., 4., 5.])
        out = dev.apply_fn(buf, lambda x: x.sin())
        out = dev.apply_fn(out, lambda x: x.cos())
        dev.apply_fn(out, lambda x: x.ln())
        assert [dev.op_hint_src(i) for i in range(3)] == ["sin(x)", "cos(x)", "log(x)"]


# ------------------------------------------------------------------ Graph + Lazy: fusing and aliasing
def chain_device(*extra):
    return CUDA("Graph", "Lazy", *extra, "Base")


def test_fused_sin_cos_ln():
    # src/op_hint.rs:121-140 / :199-219 (CUDA variant): reference tolerance 1e-3, ours <= 4 ulp per op
    with chain_device() as dev:
        x = np.array([1., 2., 3., 4., 5.], np.float32)
        buf = dev.buffer(x)
        out = dev.apply_fn(buf, lambda x: x.sin())
        out = dev.apply_fn(out, lambda x: x.cos())
        final = dev.apply_fn(out, lambda x: x.ln())
        dev.optimize_mem_graph()
        launches = dev.raw.launches
        dev.unary_fusing()
        dev.run()
        assert dev.raw.launches - launches == 1, "three recorded ops must run as ONE kernel"
        got = final.replace().read()
        want = orc.apply_chain([lambda x: x.sin(), lambda x: x.cos(), lambda x: x.ln()], orc.F32, x)
        assert_ulp(got, want, 16, "fused sin->cos->ln")
        assert dev.op_hint_src(0) == "UnaryFused"


def test_fusing_without_mem_optimisation():
    # src/op_hint.rs:142-170 (OpenCL variant: alloc_later instead of optimize_mem_graph)
    with chain_device() as dev:
        x = np.array([1., 2., 3., 4., 5.], np.float32)
        buf = dev.buffer(x)
        o1 = dev.apply_fn(buf, lambda x: x.sin())
        o2 = dev.apply_fn(o1, lambda x: x.cos())
        o3 = dev.apply_fn(o2, lambda x: x.ln())
        dev.alloc_later()
        dev.unary_fusing()
        dev.run()
        want = orc.apply_chain([lambda x: x.sin(), lambda x: x.cos(), lambda x: x.ln()], orc.F32, x)
        assert_ulp(o3.replace().read(), want, 16)
        # the intermediates were never written: they still hold the zeros of their allocation
        assert o1.replace().read().tolist() == [0.] * 5 and o2.replace().read().tolist() == [0.] * 5


def test_checkpoint_splits_the_chain():
    # src/op_hint.rs:172-196
    with chain_device() as dev:
        x = np.array([1., 2., 3., 4., 5.], np.float32)
        buf = dev.buffer(x)
        out1 = dev.apply_fn(buf, lambda x: x.sin()).checkpoint()
        out = dev.apply_fn(out1, lambda x: x.cos())
        final = dev.apply_fn(out, lambda x: x.ln())
        dev.alloc_later()
        dev.unary_fusing()
        dev.run()
        assert out.replace().read().tolist() == [0.] * 5
        assert_ulp(out1.replace().read(), orc.apply_fn(lambda x: x.sin(), orc.F32, x), 4)
        want = orc.apply_chain([lambda x: x.sin(), lambda x: x.cos(), lambda x: x.ln()], orc.F32, x)
        assert_ulp(final.replace().read(), want, 16)


def test_fusing_two_independent_chains():
    # src/op_hint.rs:225-251 (exact on the CPU; here each chain must equal the oracle within the ulp bar
    # and equal the unfused device result bit for bit)
    with chain_device() as dev:
        b, r = np.array([1., 2., 3., 4., 5.], np.float32), np.array([8., 2., 3., 4., 5.], np.float32)
        buf, rhs = dev.buffer(b), dev.buffer(r)
        out1 = dev.apply_fn(buf, lambda x: x.sin())
        out = dev.apply_fn(rhs, lambda x: x.sin())
        out2 = dev.apply_fn(out, lambda x: x.cos())
        out1 = dev.apply_fn(out1, lambda x: x.abs())
        final = dev.apply_fn(out1, lambda x: x.ln())
        assert dev.cache_traces() == [(2, [5, 6]), (3, [4])]
        dev.optimize_mem_graph()
        dev.unary_fusing()
        dev.run()
        assert_ulp(final.replace().read(), orc.apply_chain([lambda x: x.sin(), lambda x: x.abs(), lambda x: x.ln()], orc.F32, b), 16)
        assert_ulp(out2.replace().read(), orc.apply_chain([lambda x: x.sin(), lambda x: x.cos()], orc.F32, r), 8)
    with CUDA("Base") as eager:
        e = eager.apply_fn(eager.apply_fn(eager.buffer(r), lambda x: x.sin()), lambda x: x.cos())
        want_unfused = e.read()
    with chain_device() as dev:
        rhs = dev.buffer(r)
        o = dev.apply_fn(dev.apply_fn(rhs, lambda x: x.sin()), lambda x: x.cos())
        dev.optimize_mem_graph()
        dev.unary_fusing()
        dev.run()
        assert_bit_exact(o.replace().read(), want_unfused, "fused == unfused on the device")


def test_binary_op_on_a_trace_is_not_lost():
    # deliberate difference from lazy/optimization.rs:77-91: a non-unary op on a trace survives fusing
    with chain_device() as dev:
        a, b = dev.buffer([1., 2., 3., 4.]), dev.buffer([10., 20., 30., 40.])
        s = dev.apply_fn(a, lambda x: x.mul(2.))
        t = dev.apply_fn(s, lambda x: x.add(1.))
        u = dev.add(t, b)
        v = dev.apply_fn(u, lambda x: x.neg())
        dev.optimize_mem_graph()
        dev.unary_fusing()
        dev.run()
        assert v.replace().read().tolist() == [-13., -25., -37., -49.]


def test_optimize_mem_graph_aliases_traces():
    # src/modules/graph/opt_graph/optimize.rs:497-565 (test_lazy_from_retrieve)
    with chain_device() as dev:
        x = dev.buffer(np.full(1000, 1.0, np.float32))
        b = dev.buffer(np.full(1000, 1.1, np.float32))
        squared = dev.retrieve(1000, (x, x))
        add = dev.retrieve(1000, (b, x))
        mul_b = dev.retrieve(1000, (add, b))
        mul = dev.retrieve(1000, (squared, x))
        out = dev.retrieve(1000, (mul, mul_b))
        assert dev.cache_traces() == [(2, [5, 6]), (3, [4])]
        dev.optimize_mem_graph()
        dev.run()
        assert squared.ptr() == mul.ptr() == out.ptr() != 0
        assert add.ptr() == mul_b.ptr() != 0
        assert add.ptr() != squared.ptr()


def test_buffer_off_every_trace_still_gets_storage():
    # deliberate difference from lazy/optimization.rs:11 (drain): a lone retrieved buffer survives optimize
    with chain_device() as dev:
        x = dev.buffer([1., 2., 3.])
        y = dev.apply_fn(x, lambda v: v.add(1.))
        dev.optimize_mem_graph()
        dev.run()
        assert y.replace().read().tolist() == [2., 3., 4.]


def test_graph_cached_aliasing_in_a_loop():
    # src/modules/graph/opt_graph/optimize.rs:594-637
    with CUDA("Graph", "Cached", "Base") as dev:
        x = dev.buffer(np.full(1000, 1.0, np.float32))
        b = dev.buffer(np.full(1000, 1.1, np.float32))
        for i in dev.range(0, 2):
            squared = dev.retrieve(1000, (x, x))
            add = dev.retrieve(1000, (b, x))
            mul_b = dev.retrieve(1000, (add, b))
            mul = dev.retrieve(1000, (squared, x))
            out = dev.retrieve(1000, (mul, mul_b))
            if i == 0:
                assert squared.id() != mul.id()
            if i == 1:
                assert squared.id() == mul.id() == out.id()
                assert add.id() == mul_b.id()
                break
            dev.optimize_mem_graph()


def test_neural_net_traces_from_retrieve():
    # src/modules/graph/opt_graph/optimize.rs:395-447
    with CUDA("Graph", "Cached", "Base", dtype=np.int32) as dev:
        mk = lambda n, v=1: dev.buffer(np.full(n, v, np.int32))
        w1, b1, w2, b2, w3, b3, w4, b4 = mk(640), mk(64), mk(4096), mk(64), mk(4096), mk(64), mk(64), mk(1)
        inputs, targets = mk(1000), mk(100, 2)
        r = lambda n, *p: dev.retrieve(n, p, np.int32)
        a1 = r(6400, inputs, w1); a2 = r(6400, a1, b1); a2 = r(6400, a2, a2)
        a3 = r(6400, a2, w2); a4 = r(6400, a3, b2); a4 = r(6400, a4, a4)
        a5 = r(6400, a4, w3); a6 = r(6400, a5, b3); a6 = r(6400, a6, a6)
        a7 = r(100, a6, w4); a8 = r(100, a7, b4); r(100, a8, targets)
        assert dev.cache_traces() == [(10, [11, 12, 13, 14, 15, 16, 17, 18]), (19, [20, 21])]


# ------------------------------------------------------------------ Cached
def test_cached_reuses_allocations_in_a_loop():
    # src/modules/cached.rs:184-196, src/modules/autograd.rs:388-400
    with CUDA("Cached", "Base") as dev:
        buf = dev.buffer([1., 2., 3., 4.])
        seen = []
        for _ in dev.range(10):
            a = dev.apply_fn(buf, lambda x: x.add(1.))
            b = dev.apply_fn(a, lambda x: x.mul(2.))
            seen.append((a.ptr(), b.ptr()))
            assert b.read().tolist() == [4., 6., 8., 10.]
        assert len(set(seen)) == 1 and seen[0][0] != seen[0][1]
        assert dev.cursor() == 2


# ------------------------------------------------------------------ Autograd
@pytest.mark.parametrize("mods", [("Autograd", "Base"), ("Autograd", "Cached", "Base")])
def test_unary_ew_and_backward(mods):
    # src/unary.rs:160-242 (test_unary_autograd on CUDA)
    with CUDA(*mods) as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float64)).require_grad()
        out = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
        assert_ulp(out.read(), np.array(SIN), 4)
        out.backward()
        assert_ulp(buf.grad().read(), np.array(COS), 4)


def test_backward_multiple_times_cached():
    # src/unary.rs:244-277
    with CUDA("Autograd", "Cached", "Base") as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float64)).require_grad()
        for _ in range(10):
            out = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
            assert_ulp(out.read(), np.array(SIN), 4)
            out.backward()
            assert_ulp(buf.grad().read(), np.array(COS), 4)
            buf.grad().clear()


def test_backward_accumulates_under_lazy():
    # src/unary.rs:279-326 (run_several_times!): the tape is kept, grads add up i * cos(x)
    with CUDA("Autograd", "Lazy", "Base", dtype=np.float64) as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float64)).require_grad()
        out = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
        g = np.zeros(4)
        for i in range(1, 10):
            dev.run()
            assert_ulp(out.replace().read(), np.array(SIN), 4)
            out.backward()
            g = orc.add_unary_grad(lambda x: x.cos(), orc.F64, [1., 2., 3., 4.], g, np.ones(4))
            got = buf.grad().read()
            np.testing.assert_allclose(got, np.array(COS) * i, rtol=1e-14)


def test_backward_with_lazy_input():
    # src/unary.rs:328-346
    with CUDA("Autograd", "Lazy", "Base", dtype=np.float64) as dev:
        buf = dev.buffer(np.array([0., 1., 2., 3.], np.float64)).require_grad()
        buf1 = dev.apply_fn(buf, lambda x: x.add(1.))
        out = dev.unary_ew(buf1, lambda x: x.sin(), lambda x: x.cos())
        for i in range(1, 5):
            dev.run()
            out.backward()
            np.testing.assert_allclose(buf1.grad().read(), np.array(COS) * i, rtol=1e-14)


def test_backwards_at_end_of_cached_loop():
    # src/unary.rs:372-398: ten recorded ops, one backward -> 10 * cos(x)
    with CUDA("Autograd", "Cached", "Base") as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float64)).require_grad()
        for i in dev.range(0, 9):
            o = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
            if i == 0:
                o.grad().write([1., 1., 1., 1.])
        out = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
        out.backward()
        np.testing.assert_allclose(buf.grad().read(), np.array(COS) * 10, rtol=1e-14)


def test_requires_grad_chaining():
    # src/modules/autograd.rs:544-590
    with CUDA("Autograd", "Base", dtype=np.int32) as dev:
        lhs = dev.buffer(np.array([1, 2, 3, 4], np.int32)).require_grad()
        no_grad = dev.buffer(np.array([1, 2, 3, 4], np.int32))
        rhs = dev.buffer(np.array([1, 2, 3, 4], np.int32))
        assert lhs.requires_grad() and not rhs.requires_grad() and not no_grad.requires_grad()
        assert dev.retrieve(4, (lhs, rhs), np.int32).requires_grad()
        assert dev.retrieve(4, (lhs,), np.int32).requires_grad()
        assert not dev.retrieve(4, (rhs,), np.int32).requires_grad()
        assert not dev.retrieve(4, (no_grad, rhs), np.int32).requires_grad()


def test_grads_from_x_plus_3_and_disabling():
    # src/modules/autograd.rs:496-542
    with CUDA("Autograd", "Base", dtype=np.int32) as dev:
        lhs = dev.buffer(np.array([1, 2, 3, 4], np.int32)).require_grad()
        dev.disable_grad()
        out = dev.unary_ew(lhs, lambda x: x.mul(1), lambda x: x.add(3))
        out.backward()
        assert lhs.grad().read().tolist() == [0, 0, 0, 0]  # nothing was recorded
        dev.enable_grad()
        out = dev.unary_ew(lhs, lambda x: x.mul(1), lambda x: x.add(3))
        out.backward()
        assert lhs.grad().read().tolist() == [4, 5, 6, 7]


def test_no_grad_for_buffers_that_do_not_require_it():
    with CUDA("Autograd", "Base") as dev:
        buf = dev.buffer([1., 2., 3., 4.])
        out = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
        out.backward()
        assert buf.grad().read().tolist() == [0.] * 4


def test_grad_without_autograd_is_an_error():
    # src/modules/autograd.rs:424-430 (should_panic)
    with CUDA("Base") as dev:
        with pytest.raises(CustosError):
            dev.new_buffer(np.float32, 10).grad()


# ------------------------------------------------------------------ the north-star stack
def test_full_stack_chain8_forward_backward():
    # CUDA<Lazy<Graph<Autograd<Base>>>>: record 8 unary_ew ops, fuse, run, backward — SURVEY §8(d) item 3
    n = 100_003
    x = random_inputs(N.F32, n, 4, -4, 4)
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(x).require_grad()
        acts = [buf]
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            acts.append(dev.unary_ew(acts[-1], f, g))
        out = acts[-1]
        assert dev.ops_count() == 8 and dev.cache_traces() == [(1, [2, 3, 4, 5, 6, 7, 8])]
        dev.run()  # unfused: every intermediate is materialised, which backward needs
        unfused = out.replace().read()
        want = orc.apply_chain(CHAIN8, orc.F32, x)
        assert float(np.max(np.abs(unfused.astype(np.float64) - want.astype(np.float64)))) < 1e-5, "chain8 forward"
        # every recorded op, checked on the device's own input to it: <= 4 ulp / bit exact
        limits = [0, 0, 4, 4, 0, 0, 4, 0]
        for k in range(8):
            a_in, a_out = acts[k].replace().read(), acts[k + 1].replace().read()
            ref = orc.apply_fn(CHAIN8[k], orc.F32, a_in)
            if limits[k]:
                assert_ulp(a_out, ref, limits[k], f"op {k}")
            else:
                assert_bit_exact(a_out, ref, f"op {k}")
        out.backward()
        got = buf.grad().read()
        # oracle backward over the oracle's own activations
        a = [x]
        for f in CHAIN8:
            a.append(orc.apply_fn(f, orc.F32, a[-1]))
        g = np.ones(n, np.float32)
        for k in reversed(range(8)):
            g = orc.add_unary_grad(CHAIN8_GRADS[k], orc.F32, a[k], np.zeros(n, np.float32), g)
        err = np.abs(got.astype(np.float64) - g.astype(np.float64))
        assert np.all(err <= 2e-5 + 1e-4 * np.abs(g.astype(np.float64))), float(err.max())
    # forward-only, fused: one kernel, same bits as the unfused device result
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        cur = dev.buffer(x)
        for f in CHAIN8:
            cur = dev.apply_fn(cur, f)
        dev.optimize_mem_graph()
        dev.unary_fusing()
        before = dev.raw.launches
        dev.run()
        assert dev.raw.launches - before == 1
        assert_bit_exact(cur.replace().read(), unfused, "fused chain8 == unfused chain8")


def test_lazy_graph_replay_of_20_ops():
    # BASELINE configs[4]: Cached+Lazy CUDA-graph replay of a 20-op sequence on 4K-element buffers
    n = 4096
    x, y = random_inputs(N.F32, n, 70), random_inputs(N.F32, n, 71)

    def record(dev):
        a, b = dev.buffer(x), dev.buffer(y)
        cur = a
        for k in range(10):  # alternate unary pieces and binary adds so fusing cannot collapse it
            cur = dev.apply_fn(cur, CHEAP8[k % 8])
            cur = dev.add(cur, b)
        return a, cur

    with CUDA("Lazy", "Cached", "Base") as dev:
        a, out = record(dev)
        assert dev.ops_count() == 20
        dev.run()
        eager = out.replace().read()
    with CUDA("Lazy", "Cached", "Base") as dev:
        dev.set_graph_replay(True)
        a, out = record(dev)
        dev.run()
        assert dev.replay_kernel_nodes() == 20
        assert_bit_exact(out.replace().read(), eager, "graph replay == eager launches")
        launches = dev.raw.launches
        for _ in range(5):
            dev.run()
        assert dev.raw.launches - launches == 100
        assert_bit_exact(out.replace().read(), eager, "replay is idempotent on fresh inputs")
        a.write(y)  # new input data, same graph
        dev.run()
        got = out.replace().read()
    cur = y
    for k in range(10):
        cur = orc.apply_fn(CHEAP8[k % 8], orc.F32, cur)
        cur = orc.binary(0, orc.F32, cur, y)
    assert_bit_exact(got, cur, "replayed 20-op sequence vs oracle")


def test_sum_and_mean_through_the_module_layer():
    with CUDA("Lazy", "Base") as dev:
        x = np.random.default_rng(5).random(100_000, dtype=np.float32)
        buf = dev.buffer(x)
        doubled = dev.apply_fn(buf, lambda v: v.mul(2.0))
        dev.run()
        s = dev.sum(doubled)
        assert abs(float(s) - 2.0 * orc.sum_f64(orc.F32, x)) <= 1e-6 * 2.0 * orc.sum_f64(orc.F32, x)
        assert abs(float(dev.mean(buf)) - float(np.mean(x.astype(np.float64)))) < 1e-6


# ------------------------------------------------------------------ element-wise fusing
def test_elementwise_fusing_binary_with_unary_neighbours():
    a_np, b_np = random_inputs(N.F32, 50_003, 80), random_inputs(N.F32, 50_003, 81)

    def record(dev):
        a, b = dev.buffer(a_np), dev.buffer(b_np)
        s = dev.apply_fn(a, lambda x: x.mul(2.0).add(1.0))
        c = dev.apply_fn(b, lambda x: x.abs())
        t = dev.add(s, c)                      # (2a + 1) + |b|
        u = dev.apply_fn(t, lambda x: x.neg().mul(0.5))
        return a, b, s, c, t, u

    with CUDA("Graph", "Lazy", "Base") as dev:
        a, b, s, c, t, u = record(dev)
        dev.run()
        unfused = u.replace().read()
        assert dev.ops_count() == 4
    with CUDA("Graph", "Lazy", "Base") as dev:
        a, b, s, c, t, u = record(dev)
        dev.elementwise_fusing()
        assert dev.op_hint_src(3) == "Fused: (-((((x * 2.0) + 1.0) + abs(y))) * 0.5)"
        dev.alloc_later()
        before = dev.raw.launches
        dev.run()
        assert dev.raw.launches - before == 1, "four recorded ops, two inputs: ONE kernel"
        got = u.replace().read()
        assert_bit_exact(got, unfused, "fused == unfused on the device")
        want = orc.apply_fn(lambda x: x.neg().mul(0.5), orc.F32,
                            orc.binary(0, orc.F32, orc.apply_fn(lambda x: x.mul(2.0).add(1.0), orc.F32, a_np),
                                       orc.apply_fn(lambda x: x.abs(), orc.F32, b_np)))
        assert_bit_exact(got, want, "fused expression vs oracle (exact ops only)")
        assert t.replace().read().tolist()[:4] == [0.0] * 4  # fused-away intermediates are never written


def test_elementwise_fusing_respects_readers_checkpoints_and_three_inputs():
    x = random_inputs(N.F32, 10_001, 82)
    with CUDA("Graph", "Lazy", "Base") as dev:
        a, b, c = dev.buffer(x), dev.buffer(x * 2), dev.buffer(x * 3)
        s = dev.apply_fn(a, lambda v: v.add(1.0))
        p = dev.mul(s, b)            # s has two readers (p and q): must stay materialised
        q = dev.add(s, c)
        r = dev.add(p, q)            # p and q fold into r only while <= 2 inputs remain: they read s,b and s,c -> 3
        k = dev.apply_fn(r, lambda v: v.mul(0.5)).checkpoint()
        out = dev.apply_fn(k, lambda v: v.sub(1.0))   # k is checkpointed: not fused into out
        dev.elementwise_fusing()
        dev.run()
        s_np = x + np.float32(1.0)
        want_r = (s_np * (x * 2)) + (s_np + x * 3)
        assert_bit_exact(s.replace().read(), s_np, "s is still computed")
        assert_bit_exact(k.replace().read(), want_r * np.float32(0.5), "checkpointed buffer holds its value")
        assert_bit_exact(out.replace().read(), want_r * np.float32(0.5) - np.float32(1.0), "result")


def test_elementwise_fusing_collapses_the_20_op_sequence():
    # the replay workload (10 unary pieces alternating with 10 binary adds against the same rhs): every
    # intermediate has one reader and the expression only ever reads (a, b) -> a single kernel
    n = 4096
    x, y = random_inputs(N.F32, n, 70), random_inputs(N.F32, n, 71)
    with CUDA("Graph", "Lazy", "Base") as dev:
        a, b = dev.buffer(x), dev.buffer(y)
        cur = a
        for k in range(10):
            cur = dev.apply_fn(cur, CHEAP8[k % 8])
            cur = dev.add(cur, b)
        dev.elementwise_fusing()
        dev.alloc_later()
        before = dev.raw.launches
        dev.run()
        assert dev.raw.launches - before == 1
        want = x
        for k in range(10):
            want = orc.binary(0, orc.F32, orc.apply_fn(CHEAP8[k % 8], orc.F32, want), y)
        assert_bit_exact(cur.replace().read(), want, "20 fused ops vs oracle")


def test_elementwise_fusing_keeps_buffers_a_grad_function_needs():
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float32)).require_grad()
        h = dev.unary_ew(buf, lambda x: x.sin(), lambda x: x.cos())
        out = dev.unary_ew(h, lambda x: x.mul(2.0), lambda x: 2.0)
        dev.elementwise_fusing()   # h is on the tape (backward reads it): nothing may be fused away
        dev.run()
        assert_ulp(h.replace().read(), np.sin(np.array([1., 2., 3., 4.], np.float32)), 2)
        out.backward()
        assert_ulp(buf.grad().read(), (2.0 * np.cos(np.array([1., 2., 3., 4.]))).astype(np.float32), 4)


# ------------------------------------------------------------------ fused forward + fused backward (north-star stack)
def _unfused_chain8(x, seed=None):
    """The reference behaviour on the device: 8 kernels forward, 8 add_unary_grad kernels backward."""
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(x).require_grad()
        cur = buf
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = dev.unary_ew(cur, f, g)
        dev.run()
        if seed is None:
            cur.backward()
        else:
            cur.backward_with(seed)
        return cur.replace().read(), buf.grad().read()


@pytest.mark.parametrize("order", ["fuse", "alias_then_fuse", "fuse_then_alias"])
def test_fused_chain8_forward_and_backward_are_two_kernels(order):
    # BASELINE configs[2] on CUDA<Lazy<Graph<Autograd<Base>>>>: unary_fusing on a stack whose tape reads the
    # intermediates (src/unary.rs:118-128 vs src/modules/lazy/optimization.rs:77-91).  One forward kernel, one
    # recomputing backward kernel, gradients bit-identical to the 8 + 8 kernel path.
    n = 100_003
    x = random_inputs(N.F32, n, 4, -4, 4)
    x[:12] = [0.0, -0.0, np.inf, -np.inf, np.nan, 3e5, -2.5e6, 1e-40, -1e-40, 88.0, -104.0, 1e30]  # redo path, edges
    want_out, want_grad = _unfused_chain8(x)
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(x).require_grad()
        cur = buf
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = dev.unary_ew(cur, f, g)
        if order == "alias_then_fuse":
            dev.optimize_mem_graph()
        dev.unary_fusing()
        if order == "fuse_then_alias":
            dev.optimize_mem_graph()
        dev.run()  # (the first run also zero-fills the deferred allocations)
        assert_bit_exact(cur.replace().read(), want_out, "fused forward == unfused forward")
        cur.backward()
        got = buf.grad().read()
        assert_bit_exact(got, want_grad, "one chain-grad kernel == eight add_unary_grad kernels")
        assert np.all(cur.grad().read() == 1.0)  # the seed is written by the grad kernel itself
        # steady state: forward = 1 launch, backward = 1 launch (seed folded in, gradient buffers exist)
        before = dev.raw.launches
        dev.run()
        assert dev.raw.launches - before == 1
        before = dev.raw.launches
        cur.backward()
        assert dev.raw.launches - before == 1
        twice = buf.grad().read()
        with np.errstate(all="ignore"):
            assert_bit_exact(twice, (got + got).astype(np.float32), "second backward accumulates into x.grad")


def test_fused_backward_with_an_explicit_seed_and_zero_grad():
    n = 4099
    x = random_inputs(N.F32, n, 14, -4, 4)
    seed = random_inputs(N.F32, n, 15, -2, 2)
    seed[:4] = [0.0, -0.0, 1.0, -1.0]
    _, want_grad = _unfused_chain8(x, seed)
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(x).require_grad()
        cur = buf
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = dev.unary_ew(cur, f, g)
        dev.unary_fusing()
        dev.run()
        cur.backward_with(seed)
        assert_bit_exact(buf.grad().read(), want_grad, "seeded chain-grad")
        dev.zero_grad()
        assert np.all(buf.grad().read() == 0)
        cur.backward_with(seed)
        assert_bit_exact(buf.grad().read(), want_grad, "after zero_grad")


def test_unary_fusing_leaves_a_chain_alone_when_only_part_of_it_is_on_the_tape():
    # apply_fn records no grad function: the tape reads x1 (input of the unary_ew op) — fusing it away would feed
    # zeros to add_unary_grad.  The run stays unfused and the gradient stays right.
    n = 1000
    x = random_inputs(N.F32, n, 16, -2, 2)
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(x).require_grad()
        x1 = dev.apply_fn(buf, lambda v: v.mul(2.0))
        x2 = dev.unary_ew(x1, lambda v: v.sin(), lambda v: v.cos())
        x3 = dev.apply_fn(x2, lambda v: v.add(1.0))
        dev.unary_fusing()
        dev.optimize_mem_graph()
        dev.run()
        a1 = orc.apply_fn(lambda v: v.mul(2.0), orc.F32, x)
        assert_bit_exact(x1.replace().read(), a1, "x1 is still materialised")
        x2.backward()
        g = x1.grad().read()
        dev_cos = orc.apply_fn(lambda v: v.cos(), orc.F32, a1)
        assert_ulp(g, dev_cos, 4, "d sin(x1) / d x1")
        assert float(np.max(np.abs(x3.replace().read() - (np.sin(a1.astype(np.float64)) + 1)))) < 1e-5


@pytest.mark.parametrize("dtype", ["f64", "f16", "bf16", "i32"])
def test_fused_backward_other_dtypes_match_the_unfused_device_result(dtype):
    n = 10_007
    if dtype == "i32":
        code, fwd, grads = N.I32, [lambda v: v.add(3), lambda v: v.mul(5), lambda v: v.sub(7)], \
            [lambda v: v.add(1), lambda v: v.mul(2), lambda v: 3]
        x = np.random.default_rng(3).integers(-1000, 1000, n).astype(np.int32)
    else:
        code = {"f64": N.F64, "f16": N.F16, "bf16": N.BF16}[dtype]
        fwd, grads = CHAIN8, CHAIN8_GRADS
        x = random_inputs(code, n, 17, -4, 4)
    results = []
    for fuse in (False, True):
        with CUDA("Lazy", "Graph", "Autograd", "Base", dtype=code) as dev:
            buf = dev.buffer(x, dtype=code).require_grad()
            cur = buf
            for f, g in zip(fwd, grads):
                cur = dev.unary_ew(cur, f, g)
            if fuse:
                dev.unary_fusing()
            dev.run()  # (the first run also zero-fills the deferred allocations)
            before = dev.raw.launches
            dev.run()
            assert dev.raw.launches - before == (1 if fuse else len(fwd))
            cur.backward()
            results.append((cur.replace().read(), buf.grad().read()))
    assert results[0][0].tobytes() == results[1][0].tobytes(), "forward"
    assert results[0][1].tobytes() == results[1][1].tobytes(), "backward"


def test_cached_graph_runs_in_place_after_optimize_mem_graph():
    # Graph<Cached<Base>>: ids are device addresses; after optimize_mem_graph the buffers of a trace share one
    # address, so the next iteration's apply_fn(a) -> b has out == in.  Base::add_op just runs it (base.rs:53-62).
    x = random_inputs(N.F32, 5000, 18, -1, 1)
    with CUDA("Graph", "Cached", "Base") as dev:
        buf = dev.buffer(x)
        want = orc.apply_chain([lambda v: v.mul(2.0), lambda v: v.add(1.0), lambda v: v.neg()], orc.F32, x)
        for it in dev.range(4):
            a = dev.apply_fn(buf, lambda v: v.mul(2.0))
            b = dev.apply_fn(a, lambda v: v.add(1.0))
            c = dev.apply_fn(b, lambda v: v.neg())
            assert_bit_exact(c.read(), want, f"iteration {it}")
            if it == 0:
                dev.optimize_mem_graph()
            else:
                assert a.ptr() == b.ptr() == c.ptr()
    with CUDA("Base") as dev:  # eager x op x is legal too
        v = dev.buffer(x)
        assert_bit_exact(dev.add(v, v).read(), orc.binary(0, orc.F32, x, x), "x + x")


def test_autograd_state_does_not_survive_the_buffer():
    # ids of eager buffers are device addresses and the pool recycles them: a new buffer at the address of a dropped
    # one must not inherit requires_grad or the accumulated gradient
    with CUDA("Autograd", "Base") as dev:
        seen = False
        for _ in range(8):
            a = dev.buffer([1., 2., 3., 4.]).require_grad()
            out = dev.unary_ew(a, lambda v: v.mul(2.0), lambda v: 2.0)
            out.backward()
            assert a.grad().read().tolist() == [2.] * 4
            addr = a.ptr()
            out.drop()
            a.drop()
            b = dev.buffer([5., 6., 7., 8.])
            seen = seen or b.ptr() == addr
            assert not b.requires_grad()
            assert b.grad().read().tolist() == [0.] * 4
            b.drop()
        assert seen, "the pool never recycled an address: the test did not exercise the case"


def test_fusing_and_aliasing_respect_readers_the_graph_does_not_know():
    # cbm_binary_into records an op with a caller-owned output: no retrieve(), hence no graph edge (the reference's
    # tests do this with their own kernels, src/devices/cuda/lazy.rs:96-141).  x1 is read by it AND by the next unary
    # op: fusing x1 away, or letting x2 overwrite it, would hand the binary op the wrong operand.
    n = 3000
    x = random_inputs(N.F32, n, 21, -2, 2)
    for prepare in (lambda d: d.unary_fusing(), lambda d: d.optimize_mem_graph(), lambda d: (d.optimize_mem_graph(), d.unary_fusing())):
        with CUDA("Lazy", "Graph", "Base") as dev:
            buf = dev.buffer(x)
            side = dev.new_buffer(np.float32, n)
            x1 = dev.apply_fn(buf, lambda v: v.mul(2.0))
            x2 = dev.apply_fn(x1, lambda v: v.add(1.0))
            x3 = dev.apply_fn(x2, lambda v: v.neg())
            dev.add_into(x1, buf, side)  # side = x1 + x, recorded AFTER the ops that could clobber x1
            prepare(dev)
            dev.run()
            a1 = orc.apply_fn(lambda v: v.mul(2.0), orc.F32, x)
            assert_bit_exact(side.read(), orc.binary(0, orc.F32, a1, x), "the caller-owned binary op saw the real x1")
            assert_bit_exact(x3.replace().read(), orc.apply_chain([lambda v: v.mul(2.0), lambda v: v.add(1.0), lambda v: v.neg()], orc.F32, x), "chain")


def test_a_mixed_chain_is_split_where_the_tape_starts_and_both_parts_fuse():
    # apply_fn, apply_fn | unary_ew x 3 | apply_fn, apply_fn: three runs -> three forward kernels, and the three grad
    # functions of the middle part become one chain-grad kernel; x2 (the input of the recorded part) stays materialised
    n = 20_011
    x = random_inputs(N.F32, n, 23, -2, 2)
    fns = [lambda v: v.mul(0.5), lambda v: v.add(0.25)]
    ews = [(lambda v: v.sin(), lambda v: v.cos()), (lambda v: v.mul(v), lambda v: v.mul(2.0)), (lambda v: v.tanh(), lambda v: v.tanh().mul(v.tanh()).neg().add(1.0))]
    tail = [lambda v: v.neg(), lambda v: v.add(1.0)]

    def program(dev):
        buf = dev.buffer(x)
        cur = buf
        for f in fns:
            cur = dev.apply_fn(cur, f)
        mid_in = cur.require_grad()
        for f, g in ews:
            cur = dev.unary_ew(cur, f, g)
        mid_out = cur
        for f in tail:
            cur = dev.apply_fn(cur, f)
        return mid_in, mid_out, cur
    with CUDA("Autograd", "Base") as dev:  # eager reference
        mid_in, mid_out, out = program(dev)
        mid_out.backward()
        want = (out.read(), mid_in.grad().read())
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        mid_in, mid_out, out = program(dev)
        dev.unary_fusing()
        dev.run()
        before = dev.raw.launches
        dev.run()
        assert dev.raw.launches - before == 3, "three homogeneous runs, three kernels"
        mid_out.backward()  # (the first backward also creates and zero-fills the gradient buffers)
        assert_bit_exact(out.replace().read(), want[0], "forward")
        assert_bit_exact(mid_in.grad().read(), want[1], "gradient of the recorded part")
        before = dev.raw.launches
        mid_out.backward()
        assert dev.raw.launches - before == 1, "the three grad functions of the recorded part are one chain-grad kernel"
