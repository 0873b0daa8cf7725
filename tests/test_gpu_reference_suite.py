"""More of the reference's own tests for the path, restated one to one on the CUDA device through the C ABI
(the ones tests/test_gpu_modules.py does not already hold).  Each cites the reference test it follows.

The reference's CUDA tests launch their own `add` / `mul` kernels into caller-owned buffers
(`launch_kernel1d(.., &[&lhs, &rhs, &mut out, &len])`); `binary_into` is that operation here.
"""
import numpy as np
import pytest

from custos_b200 import CustosError
from custos_b200 import _native as N
from custos_b200.device import CUDA

pytestmark = pytest.mark.gpu

ONE_TO_SIX = np.array([1, 2, 3, 4, 5, 6], np.int32)


def record_three_kernels(dev, lhs, rhs, out):
    dev.add_into(lhs, rhs, out)   # out = lhs + rhs
    dev.add_into(out, lhs, rhs)   # rhs = out + lhs
    dev.mul_into(out, rhs, lhs)   # lhs = out * rhs


# ------------------------------------------------------------------ src/devices/cuda/lazy.rs
def test_lazy_cuda_run():
    # :95-141: nothing is visible before run(); afterwards the three kernels ran in order
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        lhs, rhs = dev.buffer(ONE_TO_SIX), dev.buffer(ONE_TO_SIX)
        out = lhs.empty_like()
        record_three_kernels(dev, lhs, rhs, out)
        assert out.read().tolist() == [0] * 6
        assert lhs.read().tolist() == [1, 2, 3, 4, 5, 6] and rhs.read().tolist() == [1, 2, 3, 4, 5, 6]
        dev.run()
        assert out.read().tolist() == [2, 4, 6, 8, 10, 12]
        assert rhs.read().tolist() == [3, 6, 9, 12, 15, 18]
        assert lhs.read().tolist() == [6, 24, 54, 96, 150, 216]


@pytest.mark.parametrize("graph_replay", [False, True])
def test_lazy_cuda_run_multiple_times(graph_replay):
    # :143-196: the recorded ops persist across run(); inputs rewritten between runs (CUDA-graph replay here too)
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        dev.set_graph_replay(graph_replay)
        lhs, rhs = dev.buffer(ONE_TO_SIX), dev.buffer(ONE_TO_SIX)
        out = lhs.empty_like()
        out.clear()
        record_three_kernels(dev, lhs, rhs, out)
        assert out.read().tolist() == [0] * 6
        for _ in range(10):
            lhs.write(ONE_TO_SIX)
            rhs.write(ONE_TO_SIX)
            dev.run()
        assert out.read().tolist() == [2, 4, 6, 8, 10, 12]
        assert rhs.read().tolist() == [3, 6, 9, 12, 15, 18]
        assert lhs.read().tolist() == [6, 24, 54, 96, 150, 216]
        if graph_replay:
            assert dev.replay_kernel_nodes() == 3


def test_cuda_eager_without_lazy():
    # :198-244: on CUDA<Base> the same launches run immediately
    with CUDA("Base", dtype=np.int32) as dev:
        lhs, rhs = dev.buffer(ONE_TO_SIX), dev.buffer(ONE_TO_SIX)
        out = lhs.empty_like()
        assert out.read().tolist() == [0] * 6
        record_three_kernels(dev, lhs, rhs, out)
        assert out.read().tolist() == [2, 4, 6, 8, 10, 12]
        assert rhs.read().tolist() == [3, 6, 9, 12, 15, 18]
        assert lhs.read().tolist() == [6, 24, 54, 96, 150, 216]


def test_cuda_add_ew_op_and_lazy_retrieving_exec_op():
    # :275-284 (retrieve + add_op on Base) and :286-300 (the same twice under Lazy, read through replace())
    with CUDA("Base", dtype=np.int32) as dev:
        lhs, rhs = dev.buffer(ONE_TO_SIX), dev.buffer(ONE_TO_SIX)
        assert dev.add(lhs, rhs).read().tolist() == [2, 4, 6, 8, 10, 12]
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        lhs, rhs = dev.buffer(ONE_TO_SIX), dev.buffer(ONE_TO_SIX)
        out = dev.add(lhs, rhs)
        out2 = dev.add(out, rhs)
        dev.run()
        assert out.replace().read().tolist() == [2, 4, 6, 8, 10, 12]
        assert out2.replace().read().tolist() == [3, 6, 9, 12, 15, 18]


def test_cuda_apply_fn_lazy():
    # :302-315: Graph<Lazy<Base>>, sin -> cos -> ln recorded, run() succeeds
    with CUDA("Graph", "Lazy", "Base") as dev:
        lhs = dev.buffer([1., 2., 3., 4., 5., 6.])
        out = dev.apply_fn(lhs, lambda x: x.sin())
        out = dev.apply_fn(out, lambda x: x.cos())
        final = dev.apply_fn(out, lambda x: x.ln())
        dev.run()
        want = np.log(np.cos(np.sin(np.arange(1, 7, dtype=np.float32))))
        np.testing.assert_allclose(final.replace().read(), want, rtol=1e-5)


# ------------------------------------------------------------------ src/modules/lazy.rs
def record_clear_then_add(dev):
    out = dev.retrieve(4, (), np.int32)
    dev.clear_op(out)                                        # op 0: add_op(&mut out, |out| out.clear())
    a = dev.buffer(np.array([1, 2, 3, 4], np.int32))
    b = dev.buffer(np.array([1, 2, 3, 4], np.int32))
    dev.add_into(a, b, out)                                  # op 1: add_op((&a, &b, &mut out), ..)
    return out, a, b


def test_lazy_exec_with_range():
    # :723-755: exec_now(1..) runs (and drains) the add; run() then replays only the clear
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        out, a, b = record_clear_then_add(dev)
        dev.exec_now(1, None)
        assert out.replace().read().tolist() == [2, 4, 6, 8]
        dev.run()
        assert out.replace().read().tolist() == [0] * 4


def test_lazy_exec_last_n():
    # :757-789
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        out, a, b = record_clear_then_add(dev)
        dev.exec_last_n(1)
        assert out.replace().read().tolist() == [2, 4, 6, 8]
        dev.run()
        assert out.replace().read().tolist() == [0] * 4


def test_lazy_exec_ub_testing():
    # :791-825: an operand goes out of scope before run() -> run() is an Err, not UB
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        out, a, b = record_clear_then_add(dev)
        b.drop()
        with pytest.raises(CustosError) as ei:
            dev.run()
        assert ei.value.code == N.CB_ERR_INVALID_LAZY_BUF


def test_lazy_loop_add_unary_grad_with_run():
    # :711-730: 100 recorded add_unary_grad ops on the same buffers; the assertions the reference keeps
    # commented out hold here: nothing before run(), [100; 10] after
    with CUDA("Lazy", "Base", dtype=np.int32) as dev:
        lhs = dev.new_buffer(np.int32, 10)
        lhs_grad = lhs.empty_like()
        out_grad = dev.buffer(np.ones(10, np.int32))
        for _ in range(100):
            dev.add_unary_grad(lhs, lhs_grad, out_grad, lambda x: x.add(1))
        assert dev.ops_count() == 100
        assert lhs_grad.read().tolist() == [0] * 10
        dev.run()
        assert lhs_grad.read().tolist() == [100] * 10


def test_lazy_cached_two_producers():
    # :841-853: a retrieve with two parents on Lazy<Cached<Base>>
    with CUDA("Lazy", "Cached", "Base", dtype=np.int32) as dev:
        lhs, rhs = dev.buffer(np.array([1, 2, 3, 4], np.int32)), dev.buffer(np.array([1, 2, 3, 4], np.int32))
        out = dev.retrieve(10, (lhs, rhs), np.int32)
        assert len(out) == 10


# ------------------------------------------------------------------ src/range.rs (cursor ranges of Cached)
def test_cursor_range():
    # :182-201
    with CUDA("Cached", "Base") as dev:
        for _ in dev.range(10):
            assert dev.cursor() == 0
            dev.bump_cursor()
            assert dev.cursor() == 1
            for _ in dev.range(20):
                dev.bump_cursor()
                dev.bump_cursor()
                assert dev.cursor() == 3
            assert dev.cursor() == 3
            dev.bump_cursor()
            assert dev.cursor() == 4


def test_cursor_range_pre_bumped():
    # :203-245
    with CUDA("Cached", "Base") as dev:
        dev.bump_cursor()
        dev.bump_cursor()
        for base in (2, 6):
            for _ in dev.range(10):
                assert dev.cursor() == base
                dev.bump_cursor()
                assert dev.cursor() == base + 1
                for _ in dev.range(20):
                    dev.bump_cursor()
                    dev.bump_cursor()
                    assert dev.cursor() == base + 3
                assert dev.cursor() == base + 3
                dev.bump_cursor()
                assert dev.cursor() == base + 4
            assert dev.cursor() == base + 4


def test_cache_span_resetting():
    # :247-271: span! at a call site restores the cursor that site first saw
    with CUDA("Cached", "Base") as dev:
        spans = {}
        for _ in range(10):
            dev.span(spans)
            dev.bump_cursor()
            assert dev.cursor() == 1
            for _ in range(20):
                dev.span(spans)
                dev.bump_cursor()
                dev.bump_cursor()
                assert dev.cursor() == 3
            dev.bump_cursor()
            assert dev.cursor() == 4
        assert dev.cursor() == 4


def test_cursor_range_forms():
    # :275-335: 5..=10, ..10, 5.., .. — the cursor restarts at 0 each iteration and is left at 1 afterwards
    for args, must_break in (((5, 11), False), ((10,), False), ((5, None), True), ((None,), True)):
        with CUDA("Cached", "Base") as dev:
            for _ in dev.range(*args):
                assert dev.cursor() == 0
                dev.bump_cursor()
                assert dev.cursor() == 1
                if must_break:
                    break
            dev.bump_cursor()
            assert dev.cursor() == 2


def test_retrieves_in_a_range_reuse_their_allocations():
    # src/modules/lazy.rs:553-566 + src/modules/autograd.rs:388-400: cursor ids 0, 1 and, inside a range, the
    # same device memory every iteration
    with CUDA("Cached", "Base") as dev:
        x = dev.buffer(np.arange(8, dtype=np.float32))
        seen = set()
        for _ in dev.range(5):
            a = dev.apply_fn(x, lambda v: v.add(1.0))
            b = dev.apply_fn(a, lambda v: v.mul(2.0))
            seen.add((a.ptr(), b.ptr()))
            assert b.read().tolist() == ((np.arange(8) + 1) * 2).tolist()
        assert len(seen) == 1 and dev.cursor() == 2


# ------------------------------------------------------------------ src/modules/autograd.rs
def test_grad_fn_with_lazy_buffer_source_but_no_true_lazy():
    # :432-453: Autograd<Lazy<Base>> — a grad fn `buf.grad = 5 * out.grad` runs at backward() although Lazy is on the
    # stack (backward executes eagerly); seed ones -> [5; 10].  The reference builds `out` with Buffer::new (allocated at
    # once); here it comes from unary_ew, so the recorded forward op is run first.
    with CUDA("Autograd", "Lazy", "Base") as dev:
        buf = dev.new_buffer(np.float32, 10).require_grad()
        out = dev.unary_ew(buf, lambda x: x.mul(5.0), lambda x: 5.0)
        dev.run()
        assert dev.ops_count() == 1  # backward below does not add to, or replay, the recorded forward ops
        out.backward()
        assert buf.grad().read().tolist() == [5.0] * 10


def test_grad_fn_with_out_of_scope_buffer():
    # :455-476 (#[should_panic]): the differentiated buffer left its scope before backward()
    with CUDA("Autograd", "Lazy", "Base") as dev:
        buf = dev.new_buffer(np.float32, 10).require_grad()
        out = dev.unary_ew(buf, lambda x: x.mul(5.0), lambda x: 5.0)
        dev.run()
        buf.drop()
        with pytest.raises(CustosError) as ei:
            out.backward()
        assert ei.value.code == N.CB_ERR_INVALID_LAZY_BUF


def test_tape_return_with_and_without_autograd():
    # :425-430 (#[should_panic] without Autograd) and :478-493 (grad() allocates a zeroed gradient on first use)
    with CUDA("Base") as dev:
        with pytest.raises(CustosError):
            dev.new_buffer(np.float32, 10).grad()
    with CUDA("Autograd", "Base") as dev:
        buf = dev.new_buffer(np.float32, 10)
        assert buf.grad().read().tolist() == [0.0] * 10


# ------------------------------------------------------------------ tests/cuda/scalar_ops.rs
def test_scalar_op_cuda():
    # :29-39 `lhs + 3.` on [1..5] and :41-59 `+ 1.` on 0..100000 (a buffer op with a scalar = an apply_fn with a literal)
    with CUDA("Base") as dev:
        lhs = dev.buffer(np.array([1., 2., 3., 4., 5.], np.float32))
        assert dev.apply_fn(lhs, lambda x: x.add(3.0)).read().tolist() == [4., 5., 6., 7., 8.]
        big = dev.buffer(np.arange(100000, dtype=np.float32))
        assert np.array_equal(dev.apply_fn(big, lambda x: x.add(1.0)).read(), np.arange(100000, dtype=np.float32) + 1)
