"""CPU: the C-ABI library loads and exports every symbol include/adas_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import adas_b200  # noqa: F401
from adas_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "adas_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(adas_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    assert os.path.isfile(_capi.LIB_PATH), "libadas_b200.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_capi.SYMBOLS) == declared


def test_error_without_gpu_is_loud(tmp_path):
    import torch
    if torch.cuda.is_available():
        return
    bad = tmp_path / "x.b200w"
    bad.write_bytes(b"B200PLAN" + b"\0" * 300)
    try:
        _capi.Engine(str(bad))
    except Exception as e:          # no silent CPU fallback
        assert str(e)
    else:
        raise AssertionError("engine creation must fail without a GPU / valid plan")
