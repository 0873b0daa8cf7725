"""The C ABI from plain C: tests/c_abi_smoke.c is compiled with gcc against include/custos_b200.h and
linked to the shared library — no Python, no torch in the process."""
import subprocess
from pathlib import Path

import pytest

from custos_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent


def build(tmp_path) -> Path:
    exe = tmp_path / "c_abi_smoke"
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c_abi_smoke.c"),
           "-o", str(exe), f"-L{N.LIB_PATH.parent}", "-lcustos_b200", f"-Wl,-rpath,{N.LIB_PATH.parent}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_valid_c_and_links(tmp_path):
    exe = build(tmp_path)
    assert exe.exists()


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_c_program_reports_no_device_without_a_gpu(tmp_path):
    r = subprocess.run([str(build(tmp_path))], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_program_runs_on_the_gpu(tmp_path):
    r = subprocess.run([str(build(tmp_path))], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout and "to_cl_source: sin(((x * 2.0) + 1.0))" in r.stdout
