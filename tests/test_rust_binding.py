"""bindings/rust/custos_b200_sys.rs is generated from the header (scripts/gen_rust_ffi.py).  No Rust toolchain exists
in this image, so the file is checked structurally: it is up to date, declares every function of the header, and agrees
with the ctypes binding (which the whole test-suite exercises) on the arity and pointer-ness of every argument."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

from custos_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent
RS = (ROOT / "bindings" / "rust" / "custos_b200_sys.rs").read_text()


def rust_functions():
    out = {}
    for name, args, ret in re.findall(r"pub fn (\w+)\((.*?)\) -> ([^;]+);", RS):
        out[name] = ([a.split(": ", 1)[1] for a in args.split(", ")] if args else [], ret)
    return out


def test_generated_file_is_up_to_date():
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_header_function_is_bound_with_the_same_shape_as_ctypes():
    fns = rust_functions()
    declared = set(N.SIGNATURES) | set(N._OTHER_RESTYPE)
    assert set(fns) == declared, sorted(set(fns) ^ declared)
    for name, argtypes in N.SIGNATURES.items():
        rust_args, ret = fns[name]
        assert ret == "i32", name
        assert len(rust_args) == len(argtypes), name
        for r, c in zip(rust_args, argtypes):
            c_is_pointer = c in (C.c_void_p, C.c_char_p) or hasattr(c, "contents") or getattr(c, "_type_", None) == "P"
            if c is C.c_uint64 and r == "u64":
                continue  # device addresses and buffer handles travel as integers
            assert r.startswith("*") == c_is_pointer, (name, r, c)
    assert fns["cb_last_error"] == ([], "*const c_char")
    assert fns["cb_dtype_size"] == (["i32"], "usize")


def test_type_mapping_spot_checks():
    fns = rust_functions()
    assert fns["cb_expr_compile"][0] == ["*mut cb_device", "i32", "i32", "*const *const cb_node", "*const i32", "i32",
                                         "*mut *mut cb_expr"]
    assert fns["cb_apply"][0] == ["*mut cb_device", "*mut cb_expr", "u64", "u64", "usize"]
    assert fns["cb_comm_create"][0] == ["*mut cb_device", "i32", "i32", "*const u8", "*mut *mut cb_comm"]
    assert fns["cbm_buffer_serialize"][0] == ["*mut cbm_device", "cbm_buf", "i32", "*mut c_void", "usize", "*mut usize"]
    assert "pub const CB_BF16: i32 = 7;" in RS and "pub const CBM_AUTOGRAD: u32 = 8;" in RS
    assert "pub struct cb_node { pub op: i32, pub a: i32, pub b: i32, pub _pad: i32, pub fimm: f64, pub iimm: i64 }" in RS
    assert C.sizeof(N.cb_node) == 32  # 4 x i32 + f64 + i64, the #[repr(C)] layout above
