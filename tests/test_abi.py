"""CPU-only: libvsg_cuda.so loads, exports every symbol include/vsg_cuda.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from visual_sgraphs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "vsg_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vsg_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_table_agree():
    assert header_functions() == sorted(_lib._SIGNATURES)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for name in header_functions():
        assert hasattr(L, name), name


def test_keypoint_layout_is_cv_keypoint():
    assert _lib.KEYPOINT_DTYPE.itemsize == 28
    assert [_lib.KEYPOINT_DTYPE.fields[k][1] for k in ("x", "y", "size", "angle", "response", "octave", "class_id")] == \
        [0, 4, 8, 12, 16, 20, 24]


def test_no_cpu_fallback_without_device():
    L = _lib.load()
    if L.vsg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    p = _lib.OrbParams(1000, 1.2, 8, 20, 7)
    assert L.vsg_extractor_create(C.byref(p), 0, 1, C.byref(h)) == _lib.VSG_ERR_CUDA
    assert b"no CPU fallback" in L.vsg_last_error()
    m = C.c_void_p()
    assert L.vsg_matcher_create(0, C.byref(m)) == _lib.VSG_ERR_CUDA


def test_product_never_references_the_oracle():
    """The product package and its CUDA sources must not import, link or call oracle/ (task rule 3)."""
    pkg = os.path.join(ROOT, "visual_sgraphs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src, os.path.join(dirpath, f)
