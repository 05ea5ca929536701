"""CPU tests of the drop-in boundary: the library loads and exports every symbol include/fastllm_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fastllm_b200.h")).read()
    return sorted(set(re.findall(r"FL_EXPORT\s+[\w\s\*]+?\b(fl_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fastllm_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding table out of sync with the header"


def test_reference_side_bindings_declare_every_symbol():
    """The Rust `extern "C"` block a maintainer would add and the C++ host mirror name every entry point of
    the header: the three sides of the boundary cannot drift apart silently."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rust = doc[doc.index('extern "C" {'):doc.index("pub fn check(")]
    for name in _declared_symbols():
        assert re.search(r"pub fn %s\(" % name, rust), f"{name} missing from the Rust extern block in INTEGRATION.md"


def test_fl_config_layout_matches_header():
    from fastllm_b200._lib import FlConfig
    # 12 x int32, float, (pad), double, 2 x int32, 6 x int32
    assert ctypes.sizeof(FlConfig) == 96
    assert FlConfig.rope_theta.offset == 56 and FlConfig.norm_eps.offset == 48 and FlConfig.tp_rank.offset == 64


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a device fl_init returns an error and the message says so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fastllm_b200 import _lib
    lib = _lib.load()
    rc = lib.fl_init(0)
    assert rc != 0
    assert b"CUDA" in lib.fl_last_error() or b"cuda" in lib.fl_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastllm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports the oracle"
