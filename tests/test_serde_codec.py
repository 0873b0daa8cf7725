"""The buffer serialisation codec on the host (no GPU needed).

Pinned by the reference: `CUDAPtr<i32>` of 1..=10 serialises as `Seq { len: 10 }, I32(1) .. I32(10), SeqEnd`
(src/devices/cuda/cuda_ptr.rs:170-190), i.e. a plain sequence of the elements.  The concrete texts / bytes are
what serde_json (ryu float layout) and bincode 1.x produce for such a sequence; those crates are not in the
tree, so beyond the sequence shape the checks are round trips and the crates' documented layouts.
"""
import json
import struct

import numpy as np
import pytest

from custos_b200 import CustosError
from custos_b200 import _native as N
from custos_b200 import serde


def test_reference_token_test_as_json_and_bincode():
    v = np.arange(1, 11, dtype=np.int32)
    assert serde.encode(v, N.I32) == b"[1,2,3,4,5,6,7,8,9,10]"
    raw = serde.encode(v, N.I32, serde.BINCODE)
    assert raw == struct.pack("<Q10i", 10, *range(1, 11))
    assert serde.decode(raw, N.I32, serde.BINCODE).tolist() == list(range(1, 11))
    assert serde.decode(b" [1, 2,3 ,4,5,6,7,8,9,10 ]\n", N.I32).tolist() == list(range(1, 11))


def test_float_layout_follows_ryu():
    f32 = np.array([1.0, 0.1, 1e16, 1.5e-7, 1e-5, 1e-6, 1e-7, 123456.789, 1e13, 1e12, 3.4028235e38, 0.0, -0.0, 0.3,
                    16777216, 1.17549435e-38, 1e-45, -2.5], np.float32)
    assert serde.encode(f32, N.F32) == (b"[1.0,0.1,1e16,1.5e-7,0.00001,0.000001,1e-7,123456.79,1e13,1000000000000.0,"
                                        b"3.4028235e38,0.0,-0.0,0.3,16777216.0,1.1754944e-38,1e-45,-2.5]")
    f64 = np.array([1.0, 0.1, 1e16, 1e15, 1.5e-7, 1e-5, 1e-6, 123456.789, 1e300, 5e-324, 2 / 3, 1234e7], np.float64)
    assert serde.encode(f64, N.F64) == (b"[1.0,0.1,1e16,1000000000000000.0,1.5e-7,0.00001,1e-6,123456.789,1e300,5e-324,"
                                        b"0.6666666666666666,12340000000.0]")
    # serde_json: non-finite floats are written as null, and a null is not a float when read back
    assert serde.encode(np.array([np.nan, np.inf, -np.inf], np.float32), N.F32) == b"[null,null,null]"
    with pytest.raises(CustosError) as ei:
        serde.decode(b"[null]", N.F32)
    assert ei.value.code == N.CB_ERR_PARSE


@pytest.mark.parametrize("dt", [N.F32, N.F64])
def test_float_text_round_trips_bit_exactly_and_is_valid_json(dt):
    rng = np.random.default_rng(3)
    t = {N.F32: np.float32, N.F64: np.float64}[dt]
    bits = rng.integers(0, 2 ** (8 * np.dtype(t).itemsize), 20000, dtype=np.uint64).astype({N.F32: np.uint32, N.F64: np.uint64}[dt])
    vals = bits.view(t)
    vals = vals[np.isfinite(vals)]
    vals = np.concatenate([vals, rng.standard_normal(5000).astype(t), np.array([0.0, -0.0, 1e-40, 65504.0], t)])
    text = serde.encode(vals, dt)
    back = serde.decode(text, dt)
    assert np.array_equal(back.view(np.uint8), vals.view(np.uint8))          # shortest digits round-trip
    assert np.array_equal(np.array(json.loads(text), t).view(np.uint8), vals.view(np.uint8))  # any JSON reader agrees


@pytest.mark.parametrize("dt,np_t", [(N.I8, np.int8), (N.U8, np.uint8), (N.I16, np.int16), (N.U16, np.uint16),
                                     (N.I32, np.int32), (N.U32, np.uint32), (N.I64, np.int64), (N.U64, np.uint64)])
def test_integers_both_formats(dt, np_t):
    info = np.iinfo(np_t)
    vals = np.array([info.min, info.max, 0, 1, info.max // 3], np_t)
    text = serde.encode(vals, dt)
    assert json.loads(text) == [int(v) for v in vals]
    assert np.array_equal(serde.decode(text, dt), vals)
    raw = serde.encode(vals, dt, serde.BINCODE)
    assert raw[:8] == struct.pack("<Q", vals.size) and raw[8:] == vals.astype(np.dtype(np_t).newbyteorder("<")).tobytes()
    assert np.array_equal(serde.decode(raw, dt, serde.BINCODE), vals)
    for bad in (f"[{int(info.max) + 1}]", f"[{int(info.min) - 1}]", "[1.5]", "[1e3]", "[1,]", "[1 2]", "1,2", "[1]]", "[true]"):
        with pytest.raises(CustosError) as ei:
            serde.decode(bad.encode(), dt)
        assert ei.value.code == N.CB_ERR_PARSE, bad


def test_bool_and_empty_and_errors():
    assert serde.encode(np.array([True, False, True]), N.BOOL) == b"[true,false,true]"
    assert serde.decode(b"[true, false]", N.BOOL).tolist() == [True, False]
    assert serde.encode(np.array([True, False]), N.BOOL, serde.BINCODE) == struct.pack("<Q", 2) + b"\x01\x00"
    with pytest.raises(CustosError):
        serde.decode(struct.pack("<Q", 1) + b"\x02", N.BOOL, serde.BINCODE)  # bincode rejects bool bytes > 1
    assert serde.encode(np.array([], np.float32), N.F32) == b"[]"
    assert serde.decode(b"[]", N.F32).size == 0
    assert serde.encode(np.array([], np.int64), N.I64, serde.BINCODE) == bytes(8)
    for raw in (b"\x01", struct.pack("<Q", 3) + bytes(8), struct.pack("<Q", 2 ** 62) + bytes(4)):
        with pytest.raises(CustosError) as ei:
            serde.decode(raw, N.F32, serde.BINCODE)
        assert ei.value.code == N.CB_ERR_PARSE
    # half is built without its serde feature (Cargo.toml:36): f16 / bf16 buffers are not serialisable
    for dt in (N.F16, N.BF16):
        with pytest.raises(CustosError) as ei:
            serde.encode(np.zeros(2, np.uint16).view(np.float16 if dt == N.F16 else np.uint16), dt)
        assert ei.value.code == N.CB_ERR_UNSUPPORTED


def test_untyped_type_set_is_astype():
    # src/devices/untyped/matches_type.rs:28-71
    lib = N.load()
    assert [dt for dt in range(13) if lib.cbm_untyped_supports(dt)] == sorted([N.U8, N.U32, N.I64, N.BF16, N.F16, N.F32, N.F64])


@pytest.mark.parametrize("dt,np_t", [(N.F32, np.float32), (N.F64, np.float64)])
def test_float_digits_are_the_shortest_round_trip_digits(dt, np_t):
    # ryu prints the shortest digit string that reads back to the same float (closest one on ties); NumPy's
    # `unique=True` formatting implements the same rule independently (Dragon4), so the DIGITS and the decimal
    # exponent must agree value by value — only the layout (where ryu switches to exponent form) is ours to restate.
    rng = np.random.default_rng(11)
    mags = 10.0 ** rng.uniform(-30 if np_t is np.float32 else -250, 30 if np_t is np.float32 else 250, 4000)
    vals = (rng.standard_normal(4000) * mags).astype(np_t)
    vals = vals[np.isfinite(vals) & (vals != 0)]
    text = serde.encode(vals, dt).decode()[1:-1].split(",")
    assert len(text) == vals.size
    for s, v in zip(text, vals):
        want = np.format_float_scientific(v, unique=True, trim="-")          # d.ddde+XX
        wd, we = want.lstrip("-").split("e")
        want_digits, want_exp = wd.replace(".", "").rstrip("0") or "0", int(we)
        body = s.lstrip("-")
        if "e" in body:
            m, e = body.split("e")
            digits = m.replace(".", "")
            exp10 = int(e)
        else:
            ip, fp = body.split(".")
            if ip.strip("0"):
                digits, exp10 = (ip + fp), len(ip.lstrip("0")) - 1
                digits = digits.lstrip("0")
            else:
                stripped = fp.lstrip("0")
                digits, exp10 = stripped, -(len(fp) - len(stripped)) - 1
        assert digits.rstrip("0") == want_digits and exp10 == want_exp, (s, want)
        assert s.startswith("-") == bool(v < 0)


def test_decoder_survives_arbitrary_bytes():
    """The decoder parses untrusted bytes: random garbage and mutated valid documents either decode or fail with
    CB_ERR_PARSE (run under ASan / UBSan by scripts/host_sanitize.sh)."""
    import random
    rng = random.Random(7)
    valid = [serde.encode(np.arange(-3, 4, dtype=np.int32), N.I32), serde.encode(np.array([1.5, -2e-7, 3e30], np.float32), N.F32),
             serde.encode(np.array([True, False]), N.BOOL), serde.encode(np.arange(5, dtype=np.uint8), N.U8, serde.BINCODE),
             serde.encode(np.array([1.0, 2.0]), N.F64, serde.BINCODE)]
    dtypes = [N.I8, N.U8, N.I16, N.U16, N.I32, N.U32, N.I64, N.U64, N.F32, N.F64, N.BOOL]
    outcomes = {"ok": 0, "parse": 0}
    for _ in range(6000):
        doc = bytearray(rng.choice(valid))
        for _ in range(rng.randint(0, 4)):
            roll = rng.random()
            if roll < 0.4 and doc:
                doc[rng.randrange(len(doc))] = rng.randrange(256)
            elif roll < 0.7:
                doc.insert(rng.randrange(len(doc) + 1), rng.choice(b"[],.-+eE0123456789 \x00tf\"{}n"))
            elif doc:
                del doc[rng.randrange(len(doc))]
        if rng.random() < 0.1:
            doc = bytearray(rng.randbytes(rng.randint(0, 40)))
        for fmt in (serde.JSON, serde.BINCODE):
            dt = rng.choice(dtypes)
            try:
                out = serde.decode(bytes(doc), dt, fmt)
                outcomes["ok"] += 1
                assert out.dtype == np.dtype({N.BOOL: np.bool_}.get(dt, out.dtype))
            except CustosError as e:
                assert e.code == N.CB_ERR_PARSE, (bytes(doc), dt, fmt, e)
                outcomes["parse"] += 1
    assert outcomes["ok"] > 200 and outcomes["parse"] > 2000
    with pytest.raises(CustosError):
        serde.decode(b"[1\x00]", N.I32)  # a NUL is not part of a number
    # the JSON number grammar, as serde_json enforces it
    for ok in (b"[1]", b"[-0]", b"[0.5]", b"[1e5]", b"[1E-5]", b"[-1.25e+3]", b"[1e-999]"):
        serde.decode(ok, N.F64)
    for bad in (b"[+1]", b"[01]", b"[1.]", b"[.5]", b"[1e]", b"[--1]", b"[1e999]", b"[0x10]", b"[1_0]", b"[1,,2]", b"[Infinity]", b"[NaN]"):
        with pytest.raises(CustosError) as ei:
            serde.decode(bad, N.F64)
        assert ei.value.code == N.CB_ERR_PARSE, bad
