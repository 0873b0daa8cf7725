"""Shared helpers of the parity tests: ulp distance, edge-case inputs, dtype tables."""
import numpy as np

from custos_b200 import _native as N

from custos_b200.expr import bf16_from_f32, bf16_to_f32  # noqa: E402

# bf16 buffers are uint16 bit patterns on the host (numpy has no bfloat16)
NP = {N.F32: np.float32, N.F64: np.float64, N.F16: np.float16, N.I32: np.int32, N.I64: np.int64,
      N.U32: np.uint32, N.U8: np.uint8, N.BF16: np.uint16, N.I8: np.int8, N.I16: np.int16, N.U16: np.uint16,
      N.U64: np.uint64, N.BOOL: np.bool_}
UINT_OF = {N.F32: np.uint32, N.F64: np.uint64, N.F16: np.uint16}
INT_OF = {N.F32: np.int64, N.F64: np.int64, N.F16: np.int64}


def bits(a: np.ndarray) -> np.ndarray:
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize])


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two float arrays of the same dtype.
    NaN vs NaN counts as 0 (payloads are not compared), NaN vs number as a huge value."""
    assert a.dtype == b.dtype and a.shape == b.shape
    nbits = a.dtype.itemsize * 8
    ua = bits(a).astype(np.int64) if nbits < 64 else bits(a).view(np.int64)
    ub = bits(b).astype(np.int64) if nbits < 64 else bits(b).view(np.int64)
    sign = np.int64(1) << (nbits - 1)
    if nbits < 64:
        oa = np.where(ua & sign, sign - ua, ua)   # map to a monotonic integer line
        ob = np.where(ub & sign, sign - ub, ub)
        d = np.abs(oa - ob).astype(np.float64)
    else:
        oa = np.where(ua < 0, np.int64(-(2 ** 63)) - ua, ua)
        ob = np.where(ub < 0, np.int64(-(2 ** 63)) - ub, ub)
        # exact in 64-bit integers when the signs agree (float64 cannot hold 2^63-sized differences)
        same_sign = (oa < 0) == (ob < 0)
        with np.errstate(over="ignore"):
            exact = np.abs(oa - ob).astype(np.float64)
        d = np.where(same_sign, exact, np.abs(oa.astype(np.float64) - ob.astype(np.float64)))
    na, nb = np.isnan(a), np.isnan(b)
    d = np.where(na & nb, 0.0, d)
    d = np.where(na ^ nb, 1e30, d)
    return d


def assert_bit_exact(got: np.ndarray, want: np.ndarray, what: str = ""):
    if got.dtype.kind == "f":
        both_nan = np.isnan(got) & np.isnan(want)
        same = (bits(got) == bits(want)) | both_nan
    else:
        same = got == want
    if not np.all(same):
        bad = np.flatnonzero(~same)
        i = bad[0]
        raise AssertionError(f"{what}: {bad.size} of {got.size} elements differ; first at {i}: got {got[i]!r} want {want[i]!r}")


def bf16_is_nan(b: np.ndarray) -> np.ndarray:
    return (b & 0x7fff) > 0x7f80


def assert_bf16_bit_exact(got: np.ndarray, want: np.ndarray, what: str = ""):
    """bf16 bit patterns; NaN payloads are not compared (cvt.rn.bf16.f32 canonicalises, `half` keeps bits)."""
    same = (got == want) | (bf16_is_nan(got) & bf16_is_nan(want))
    if not np.all(same):
        bad = np.flatnonzero(~same)
        i = bad[0]
        raise AssertionError(f"{what}: {bad.size} of {got.size} bf16 elements differ; first at {i}: "
                             f"got {int(got[i]):#06x} want {int(want[i]):#06x}")


def bf16_ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ua, ub = a.astype(np.int64), b.astype(np.int64)
    oa = np.where(ua & 0x8000, 0x8000 - ua, ua)
    ob = np.where(ub & 0x8000, 0x8000 - ub, ub)
    d = np.abs(oa - ob).astype(np.float64)
    na, nb = bf16_is_nan(a), bf16_is_nan(b)
    d = np.where(na & nb, 0.0, d)
    return np.where(na ^ nb, 1e30, d)


def assert_same(dt: int, got: np.ndarray, want: np.ndarray, what: str = ""):
    (assert_bf16_bit_exact if dt == N.BF16 else assert_bit_exact)(got, want, what)


def assert_ulp(got: np.ndarray, want: np.ndarray, max_ulp: float, what: str = ""):
    d = ulp_distance(got, want)
    worst = int(np.argmax(d))
    assert d[worst] <= max_ulp, f"{what}: {d[worst]:.0f} ulp at {worst}: got {got[worst]!r} want {want[worst]!r} (limit {max_ulp})"
    return float(d.max()), float(d.mean())


def edge_values(np_dtype) -> np.ndarray:
    """±0, denormals, ±inf, NaN, extremes and a log-spaced sweep of magnitudes."""
    fi = np.finfo(np_dtype)
    special = [0.0, -0.0, fi.tiny, -fi.tiny, fi.smallest_subnormal, -fi.smallest_subnormal, fi.tiny / 4,
               fi.max, -fi.max, np.inf, -np.inf, np.nan, 1.0, -1.0, 0.5, 2.0, np.pi, -np.pi, np.e]
    lo, hi = np.log10(float(fi.tiny)) + 0.5, np.log10(float(fi.max)) - 0.5
    sweep = 10.0 ** np.linspace(lo, hi, 400)
    with np.errstate(over="ignore"):
        return np.concatenate([np.array(special, np.float64), sweep, -sweep]).astype(np_dtype)


def random_inputs(dtype_code: int, n: int, seed: int, lo=-4.0, hi=4.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = NP[dtype_code]
    if dtype_code == N.BF16:
        return bf16_from_f32(rng.uniform(lo, hi, n).astype(np.float32))
    if dtype_code == N.BOOL:
        return rng.integers(0, 2, n).astype(np.bool_)
    if np.dtype(t).kind == "f":
        return rng.uniform(lo, hi, n).astype(t)
    info = np.iinfo(t)
    return rng.integers(max(info.min, -1000), min(info.max, 1000) + 1, n).astype(t)


from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1  # noqa: E402,F401
