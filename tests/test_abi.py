"""The C-ABI library loads and exports exactly what include/mfr_b200.h declares (no GPU needed)."""
import ctypes
import re

from helpers import ROOT
from mfr_b200 import lib


def _declared():
    src = open(ROOT + "/include/mfr_b200.h").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in the header but not exported by libmfr_b200.so"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in mfr_b200/lib.py"
    for n in lib.SIGNATURES:
        assert n in names, f"{n} bound in lib.py but not declared in include/mfr_b200.h"


def test_library_loads_without_gpu():
    l = lib.load()
    assert l.mfr_version() >= 100
    assert l.mfr_device_sm_count() >= 0
