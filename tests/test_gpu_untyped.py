"""GPU tests of the Untyped device and of (de)serialising device buffers.

Each test restates a test of the reference (cited) on the CUDA device through the C ABI.
"""
import struct

import numpy as np
import pytest

from custos_b200 import CustosError
from custos_b200 import _native as N
from custos_b200.device import CUDA, Untyped
from custos_b200.expr import bf16_from_f32
from oracle import oracle as orc
from tests.helpers import assert_bit_exact, random_inputs

pytestmark = pytest.mark.gpu


def test_to_untyped_keeps_the_storage_tag():
    # src/devices/untyped/mod.rs:91-110 (test_to_untype_buf): CudaStorage::F32 holding [1,2,3,4]
    with Untyped() as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float32)).to_untyped()
        assert buf.dtype is None and buf.storage_dtype() == N.F32
        assert buf.read_typed(np.float32).tolist() == [1., 2., 3., 4.]
        with pytest.raises(TypeError):
            buf.read()


def test_add_type_info_to_untyped():
    # mod.rs:112-118, :129-136 (by value and by reference)
    with Untyped() as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float32)).to_untyped()
        typed = buf.to_typed(np.float32)
        assert typed is not None and typed.read().tolist() == [1., 2., 3., 4.]
        assert buf.as_untyped().as_typed(np.float32).read().tolist() == [1., 2., 3., 4.]


def test_add_type_info_type_mismatch_is_none():
    # mod.rs:120-127, :129-134: to_typed::<u32>() / as_typed::<u32>() of an f32 storage -> None
    with Untyped() as dev:
        buf = dev.buffer(np.array([1., 2., 3., 4.], np.float32))
        assert buf.to_untyped().to_typed(np.uint32) is None
        assert buf.as_typed(np.uint32) is None
        assert buf.read_typed(np.uint32) is None
        with pytest.raises(CustosError) as ei:
            N.call("cbm_buffer_matches_type", dev.h, buf.handle, N.U32)
        assert ei.value.code == N.CB_ERR_TYPE_MISMATCH


def test_apply_fn_untyped():
    # src/devices/untyped/ops.rs:167-173: x.add(1.) on [1,2,3,4]
    with Untyped() as dev:
        res = dev.buffer(np.array([1., 2., 3., 4.], np.float64))
        assert dev.apply_fn(res, lambda x: x.add(1.)).read().tolist() == [2., 3., 4., 5.]
        # the same through a type-erased view: the closure is typed by the storage tag
        out = dev.apply_fn(res.to_untyped(), lambda x: x.add(1.))
        assert out.dtype is None and out.read_typed(np.float64).tolist() == [2., 3., 4., 5.]


def test_typed_and_untyped_ew_add():
    # ops.rs:228-246: add of [1,2,3,4] + [1,2,3,4], typed and untyped operands
    with Untyped() as dev:
        lhs, rhs = dev.buffer(np.array([1., 2., 3., 4.], np.float32)), dev.buffer(np.array([1., 2., 3., 4.], np.float32))
        assert dev.add(lhs, rhs).read_typed(np.float32).tolist() == [2., 4., 6., 8.]
        out = dev.add(lhs.to_untyped(), rhs.to_untyped())
        assert out.read_typed(np.float32).tolist() == [2., 4., 6., 8.]
        # storages of different types: `_ => unimplemented!()` (ops.rs:104,143)
        other = dev.buffer(np.array([1, 2, 3, 4], np.uint32))
        with pytest.raises(NotImplementedError):
            dev.add(lhs.to_untyped(), other.to_untyped())


def test_every_astype_storage_dispatches():
    # untyped_binary_op! has an arm per storage type: U8, U32, I64, BF16, F16, F32, F64 (ops.rs:84-140)
    with Untyped() as dev:
        for dt in (N.U8, N.U32, N.I64, N.F16, N.F32, N.F64, N.BF16):
            a, b = random_inputs(dt, 1001, 1), random_inputs(dt, 1001, 2)
            out = dev.add(dev.buffer(a, dtype=dt).to_untyped(), dev.buffer(b, dtype=dt).to_untyped())
            assert out.storage_dtype() == dt
            got, want = out.read_typed(dt), orc.binary(N.BIN_ADD, dt, a, b)
            assert got.tobytes() == want.tobytes(), dt
        for dt in (N.I32, N.I8, N.U64, N.BOOL):  # no AsType impl (matches_type.rs:28-60)
            with pytest.raises(TypeError):
                dev.buffer(np.zeros(4, np.int32), dtype=dt)


# ------------------------------------------------------------------ serde of device buffers
def test_ser_de_of_cuda_ptr_filled():
    # src/devices/cuda/cuda_ptr.rs:170-190: CUDAPtr<i32> holding 1..=10
    with CUDA("Base") as dev:
        buf = dev.buffer(np.arange(1, 11, dtype=np.int32))
        assert buf.to_tokens() == [("Seq", 10)] + [("I32", i) for i in range(1, 11)] + [("SeqEnd",)]
        assert buf.serialize() == b"[1,2,3,4,5,6,7,8,9,10]"
        assert buf.serialize(N.SER_BINCODE) == struct.pack("<Q10i", 10, *range(1, 11))
        back = dev.deserialize(b"[1,2,3,4,5,6,7,8,9,10]", np.int32)
        assert back.read().tolist() == list(range(1, 11)) and back.ptr() != buf.ptr()
        back2 = dev.deserialize(buf.serialize(N.SER_BINCODE), np.int32, N.SER_BINCODE)
        assert back2.read().tolist() == list(range(1, 11))


@pytest.mark.parametrize("fmt", [N.SER_JSON, N.SER_BINCODE])
def test_serialise_results_of_the_hot_path_round_trip(fmt):
    # a computed buffer -> bytes -> a new device buffer holds the same bits (f32, f64, ints, bool)
    n = 100_003
    with CUDA("Base") as dev:
        for dt in (N.F32, N.F64):
            x = random_inputs(dt, n, 5)
            out = dev.apply_fn(dev.buffer(x), lambda v: v.mul(2.0).add(1.0).sin())
            data = out.serialize(fmt)
            assert_bit_exact(dev.deserialize(data, dt, fmt).read(), out.read(), f"round trip dtype {dt}")
        for dt in (N.I8, N.U16, N.I64, N.U64, N.BOOL):
            x = random_inputs(dt, n, 6)
            b = dev.buffer(x, dtype=dt)
            assert np.array_equal(dev.deserialize(b.serialize(fmt), dt, fmt).read(), x)


def test_serde_of_lazy_buffers_and_errors():
    with CUDA("Lazy", "Base") as dev:
        out = dev.apply_fn(dev.buffer(np.array([1., 2., 3.], np.float32)), lambda v: v.mul(2.0))
        dev.run()
        assert out.replace().serialize() == b"[2.0,4.0,6.0]"
    with CUDA("Base") as dev:
        half = dev.buffer(bf16_from_f32(np.ones(4, np.float32)), dtype=N.BF16)
        with pytest.raises(CustosError) as ei:   # half has no serde feature in the reference build
            half.serialize()
        assert ei.value.code == N.CB_ERR_UNSUPPORTED
        with pytest.raises(CustosError) as ei:
            dev.deserialize(b"[1,2", np.int32)
        assert ei.value.code == N.CB_ERR_PARSE
        with pytest.raises(CustosError) as ei:   # CUDAPtr::new(0) fails (api/cuda.rs:69-71)
            dev.deserialize(b"[]", np.int32)
        assert ei.value.code == N.CB_ERR_ZERO_LENGTH
