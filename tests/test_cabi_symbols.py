"""The C-ABI library loads on a CPU-only box and exports every symbol include/fhsim.h declares."""
import ctypes
import os
import re

import pytest

from fhsim import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fhsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 35
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [s for s in syms if not hasattr(handle, s)]
    assert not missing, f"missing exports: {missing}"
    bound = set(_cabi.SIGNATURES) | set(_cabi._SPECIAL)
    assert set(syms) == bound, f"ctypes binding out of sync with header: {set(syms) ^ bound}"


def test_version_and_error_string():
    L = _cabi.lib()
    assert L.fh_version() >= 100
    assert isinstance(L.fh_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from fhsim.backend import Context
    with pytest.raises(_cabi.FhsimError, match="no CPU fallback|no CUDA device"):
        Context(0)
