"""The C-ABI library builds here (nvcc cross-compiles), loads without a GPU, and exports every symbol
include/b200r.h declares.  No compute call is made: without a CUDA device they fail by design."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200r.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200r_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from a_modular_rag_framework_b200 import _native
    assert _declared() == sorted(_native.EXPORTS)


def test_library_builds_loads_and_exports_everything():
    from a_modular_rag_framework_b200 import _native, build
    path = build.build()
    assert os.path.exists(path)
    lib = _native.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.b200r_version() == 100


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    from a_modular_rag_framework_b200 import DenseIndex
    with pytest.raises(ValueError):
        DenseIndex(torch.zeros((8, 8), dtype=torch.bfloat16))
    from a_modular_rag_framework_b200 import _native
    lib = _native.load()
    h = ctypes.c_uint64(0)
    buf = (ctypes.c_uint8 * 256)()
    rc = lib.b200r_index_create(0, ctypes.addressof(buf), 4, 8, _native.BF16, _native.COS, 0, None, None, ctypes.byref(h))
    assert rc != 0 and h.value == 0 and _native.last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "a_modular_rag_framework_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in text and "from oracle" not in text, f
