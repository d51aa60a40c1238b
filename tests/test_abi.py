"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/cattl3_b200.h
declares, and refuses (loudly) to compute without a CUDA device -- there is no CPU fallback."""
import ctypes
import os
import re

import pytest

from __graft_entry__ import ROOT, load_package


@pytest.fixture(scope="module")
def pkg():
    p = load_package()
    if not os.path.exists(p.LIB_PATH):
        p.build()
    return p


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cattl3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cattl3_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), "declared in cattl3_b200.h but not exported: " + n


def test_python_binding_covers_header(pkg):
    assert set(_declared_symbols()) == set(pkg._SYMBOLS)


def test_abi_version_and_geometry_helpers(pkg):
    L = pkg.lib()
    assert L.cattl3_abi_version() == 1
    g = pkg.ConvGeom(256, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)
    oh, ow = ctypes.c_int32(), ctypes.c_int32()
    assert L.cattl3_conv_output_dims(ctypes.byref(g), 0, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (56, 56)
    # test/gradient_test.cpp:192-204 transposed case {2,3,2}, R=5x3, pad 1x0, stride 1x2, dilation 0x1
    g = pkg.ConvGeom(5, 2, 3, 2, 5, 5, 3, 1, 0, 1, 2, 0, 1)
    assert L.cattl3_conv_output_dims(ctypes.byref(g), 1, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (4, 9)
    bad = pkg.ConvGeom(1, 2, 2, 1, 1, 5, 5, 0, 0, 1, 1, 0, 0)  # receptor larger than input
    assert L.cattl3_conv_output_dims(ctypes.byref(bad), 0, ctypes.byref(oh), ctypes.byref(ow)) == pkg.ERR_INVALID
    assert b"receptor" in L.cattl3_last_error()
    p = pkg.PoolGeom(8, 32, 32, 8, 2, 2, 2, 2)
    assert L.cattl3_pool_output_dims(ctypes.byref(p), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (16, 16)


def test_no_cpu_fallback(pkg):
    """Without a device the context cannot be created; with one, nothing here computes."""
    if pkg.lib().cattl3_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.Cattl3Error) as e:
        pkg.Context(0)
    assert e.value.code == pkg.ERR_NO_DEVICE


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under c-attl3_b200/ or include/ may reference it."""
    for base in ("c-attl3_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) and "test" not in f:
                    assert "oracle" not in open(os.path.join(dp, f), errors="ignore").read().lower(), (dp, f)
