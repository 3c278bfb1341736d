"""CPU: the C-ABI library loads and exports every symbol include/kalign_b200.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from kalign_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "kalign_b200.h")).read()
    declared = set(re.findall(r"\b(kb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"kalign_b200" in lib.kb200_version()


def test_params_init_defaults():
    p = _lib.make_params(0)
    assert (p.gpo, p.gpe, p.tgpe, p.nalpha) == (7.0, 1.25, 1.0, 23)
    assert p.subm[0] == 4.0
    p = _lib.make_params(1)
    assert p.nalpha == 5 and abs(p.gpo - 217.0) < 1e-6
    p = _lib.make_params(1, 0)
    assert (p.gpo, p.gpe, p.tgpe) == (8.0, 6.0, 0.0)


def test_no_cpu_fallback():
    lib = _lib.load()
    if lib.kb200_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.kb200_ctx_create(0, C.byref(h)) != 0
    with pytest.raises(RuntimeError):
        _lib.Context(0)
