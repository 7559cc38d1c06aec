"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/wise_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from wise_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wise_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in wise_b200.h but not exported by libwiseb200.so"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature in wise_b200/_capi.py"
    assert sorted(_capi.SIGNATURES) == names


def test_no_torch_or_python_types_in_header():
    src = open(os.path.join(ROOT, "include", "wise_b200.h")).read()
    assert 'extern "C"' in src
    for banned in ("torch", "at::", "PyObject", "std::"):
        assert banned not in src


def test_version_and_error_string():
    L = _capi.lib()
    assert b"sm_100a" in L.wb_version()


def _cuda_available():
    n = ctypes.c_int(0)
    return _capi.lib().wb_device_count(ctypes.byref(n)) == 0 and n.value > 0


@pytest.mark.skipif(_cuda_available(), reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu():
    from wise_b200 import faiss_compat as faiss
    with pytest.raises(RuntimeError) as e:
        faiss.IndexFlatIP(64)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under wise_b200/ (or the measurement scripts) may reference it."""
    for top in ("wise_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
