"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol
include/snprel_b200.h declares, and refuses to run without a device (no CPU
fallback).  No compute calls."""
import ctypes
import os
import re

import pytest

import snprelate_b200 as S
from snprelate_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "snprel_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snprel_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(S.library_path())
    declared = _declared_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in snprel_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_version_string():
    lib = S.load_library()
    assert b"sm_100a" in lib.snprel_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SNPRelError, match="no CUDA device"):
        S.Context(0)


def test_null_context_is_an_error_not_a_crash():
    lib = S.load_library()
    assert lib.snprel_geno_begin(None, 10, 10) != 0
    assert lib.snprel_kernel_launches(None) == 0
