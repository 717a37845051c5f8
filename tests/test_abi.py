"""The C-ABI library builds for sm_100a, loads without a GPU and exports every
symbol include/imsim_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from imsim_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    lib = _lib.load()
    assert lib.b2_abi_version() == _abi.B2_ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "imsim_b200.h")).read()
    # drop comments and typedef'd function-free regions, then find prototypes
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", code))
    assert declared, "no prototypes found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)


def test_struct_sizes_match():
    lib = _lib.load()
    for which, cls in enumerate(_abi.SIZEOF_ORDER):
        assert lib.b2_sizeof(which) == C.sizeof(cls), cls.__name__


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.b2_ctx_create(0, None, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.b2_last_error()
    from imsim_b200 import OpticsContext, B2Error

    with pytest.raises(B2Error):
        OpticsContext()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                bad = re.search(r"(import\s+oracle|from\s+oracle|liboracle|\boracle[./]|orc_[a-z_]+\()", src)
                assert bad is None, "%s reaches into the oracle: %r" % (f, bad and bad.group(0))
