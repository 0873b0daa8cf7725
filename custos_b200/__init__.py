"""custos_b200 — Blackwell-native CUDA backend for the data-parallel hot path of custos.

The product is the C-ABI library (include/custos_b200.h, built from custos_b200/csrc by
`python -m custos_b200.build`); this package is the thin Python plumbing around it that
mirrors the reference's operator interface for tests and benchmarks.
"""
from . import _native  # noqa: F401
from ._native import CustosError  # noqa: F401
from .expr import Combiner, Resolve  # noqa: F401
