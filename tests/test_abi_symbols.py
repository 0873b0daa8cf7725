"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/custos_b200.h declares; compute entry points fail loudly instead of falling back."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from custos_b200 import _native as N
from custos_b200 import CustosError

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "custos_b200.h").read_text()


def declared_functions():
    return sorted(set(re.findall(r"^\s*(?:const char \*|int32_t |size_t )\s*(cbm?_\w+)\s*\(", HEADER, re.M)))


def test_header_declares_a_reasonable_surface():
    names = declared_functions()
    assert len(names) > 90
    for must in ("cb_apply", "cb_unary_grad", "cb_binary", "cb_sum", "cb_clear", "cb_copy", "cb_graph_launch",
                 "cb_comm_sum", "cbm_apply_fn", "cbm_unary_ew", "cbm_unary_fusing", "cbm_backward"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(N.LIB_PATH))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    bound = set(N.SIGNATURES) | set(N._OTHER_RESTYPE)
    assert set(declared_functions()) <= bound, sorted(set(declared_functions()) - bound)
    N.load()  # every signature resolves


def test_no_link_time_dependency_on_the_driver_or_the_oracle():
    out = subprocess.run(["ldd", str(N.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "liboracle" not in out
    assert "libnvrtc" in out


def test_abi_version_and_dtype_sizes():
    lib = N.load()
    assert lib.cb_abi_version() == 1
    assert [lib.cb_dtype_size(i) for i in range(13)] == [4, 8, 2, 4, 8, 4, 1, 2, 1, 2, 2, 8, 1]  # every CDatatype
    assert lib.cb_dtype_size(99) == 0


def test_product_does_not_reference_the_oracle():
    # comments may name the oracle (e.g. where an order is restated); code may not import, include,
    # link or call it
    forbidden = re.compile(r"liboracle|from\s+oracle|import\s+oracle|oracle\.h|oracle\.py|\borc_\w+\s*\(|\borc\.")
    for p in (ROOT / "custos_b200").rglob("*"):
        if p.suffix in (".py", ".cpp", ".cu", ".h", ".cuh") and "build" not in p.parts:
            assert not forbidden.search(p.read_text()), p


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_a_gpu():
    from custos_b200.raw import RawDevice
    with pytest.raises(CustosError) as ei:
        RawDevice(0)
    assert ei.value.code == N.CB_ERR_NO_DEVICE
    from custos_b200.device import CUDA
    with pytest.raises(CustosError):
        CUDA("Base")


def test_sum_plan_and_shard_ranges_host_logic():
    from custos_b200.raw import shard_range, sum_plan
    p = sum_plan(N.F32, 1 << 30)
    assert p["threads"] == 256 and p["vec"] == 4 and p["blocks"] <= 1184 and p["chunk"] % 1024 == 0
    assert p["blocks"] * p["chunk"] >= 1 << 30 > (p["blocks"] - 1) * p["chunk"]
    assert sum_plan(N.F32, 1)["blocks"] == 1
    assert sum_plan(N.F16, 5000)["vec"] == 8 and sum_plan(N.F64, 5000)["vec"] == 2
    for n in (0, 1, 5, 1000, (1 << 28), (1 << 30) + 7):
        for ranks in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(ranks):
                b, e = shard_range(n, 4, ranks, r)
                assert b == prev and e >= b and (b % 4 == 0 or b == n)
                prev = e
            assert prev == n
