"""The C-ABI library loads and exports every symbol include/mag.h declares; without a GPU it refuses to work
(no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mag.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mag_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from core_b200._lib import lib, SYMBOLS, LIB_PATH
    names = header_functions()
    assert len(names) >= 25
    assert sorted(SYMBOLS) == names, "core_b200/_lib.py SYMBOLS and include/mag.h disagree"
    raw = C.CDLL(LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libmag.so does not export %s" % n
    lib()


def test_signatures_are_plain_c():
    """extern "C", plain pointers and sizes only: no C++ or torch types in the header."""
    src = open(os.path.join(ROOT, "include", "mag.h")).read()
    assert 'extern "C"' in src
    for bad in ("std::", "torch", "at::", "Tensor", "template", "class "):
        assert bad not in src


def test_header_cites_reference():
    src = open(os.path.join(ROOT, "include", "mag.h")).read()
    for cite in ("maSize.cc", "maQuality.cc", "maAdapt.cc", "maRefine.cc", "maCoarsen.cc", "maShape.cc"):
        assert cite in src


def test_no_cpu_fallback(built):
    """Without a CUDA device mag_create must fail loudly with MAG_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import core_b200 as cb
    with pytest.raises(cb.MagError) as ei:
        cb.Part(0)
    assert ei.value.code == 1
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """core_b200/ must not import, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "core_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "ma_oracle", "libref_oracle", "oracle/_"):
                    assert bad not in txt, "%s mentions %s" % (f, bad)
