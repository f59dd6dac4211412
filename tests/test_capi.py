"""The C-ABI boundary, checked without a GPU: the product library loads, exports every symbol
include/*.h declares, and -- because it has no CPU path -- refuses loudly to serve when no CUDA
device is present."""
import ctypes
import os
import re

import pytest

from _common import ROOT, have_data, model_path

LIB = os.path.join(ROOT, "blingfire_b200", "lib", "libblingfiretokdll.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="product library not built (run __graft_entry__.build())")


def declared_symbols():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"^[A-Za-z_][A-Za-z0-9_\s\*]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    # the north_star's four plus the additive batch calls must be among them
    for must in ("LoadModel", "FreeModel", "TextToIds", "TextToWords", "TextToIdsBatch", "TextToIdsBatchCsr",
                 "TextToIdsBatchDevice", "GetBlingFireTokVersion", "SetModel"):
        assert must in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"


def test_version_matches_reference():
    lib = ctypes.CDLL(LIB)
    lib.GetBlingFireTokVersion.restype = ctypes.c_int
    assert lib.GetBlingFireTokVersion() == 18000   # blingfiretokdll.cpp:107-111


def test_python_binding_loads():
    import blingfire_b200 as bf
    assert bf.get_blingfiretok_version() == 18000


def test_null_and_degenerate_arguments_do_not_crash():
    import blingfire_b200 as bf
    L = bf.lib()
    assert L.FreeModel(None) == 0                                   # blingfiretokdll.cpp:1656-1658
    buf = (ctypes.c_int32 * 4)()
    assert L.TextToIds(None, b"abc", 3, buf, 4, 0) == 0             # :1629
    assert not L.LoadModel(b"/nonexistent/model.bin")               # reference would std::terminate
    assert bf.last_error() != ""
    assert not L.SetModel(None, 0)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device LoadModel must fail with a CUDA error, not serve from the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if not have_data():
        pytest.skip("data/ not staged")
    import blingfire_b200 as bf
    h = bf.lib().LoadModel(model_path("bert_base_tok.bin").encode())
    assert not h
    assert "cuda" in bf.last_error().lower()
