"""The C-ABI shared library loads and exports every symbol include/zkb200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_what_python_binds():
    from zelana_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 40
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    import zelana_b200
    lib = zelana_b200.load_library()          # raises if the CUDA extension is missing: there is no CPU fallback
    for name in _declared_functions():
        assert getattr(lib, name) is not None, name
    assert b"sm_100a" in lib.zkb_version()
    assert lib.zkb_prof_phase_count() >= 12
    assert lib.zkb_prof_phase_name(2) == b"msm_g1_accumulate"


def test_no_device_is_an_error_not_a_fallback():
    import torch
    import zelana_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = zelana_b200.load_library()
    assert lib.zkb_device_count() == 0
    h = C.c_void_p()
    assert lib.zkb_ctx_create(0, C.byref(h)) == -1   # ZKB_ERR_NO_DEVICE
    with pytest.raises(zelana_b200.ZkbError):
        zelana_b200.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "zelana_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "zkoracle" not in text, f
