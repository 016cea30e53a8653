"""The C-ABI library loads and exports every symbol include/azb200.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

from azb200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(headers=("azb200.h",)):
    names = set()
    for h in headers:
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(azb_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libazb200.so lacks {missing}"


def test_nn_header_symbols_are_exported():
    """include/azb200_nn.h: the fused leaf evaluators (tcgen05 and mma.sync) and the pinned-tensor upload."""
    text = open(os.path.join(ROOT, "include", "azb200_nn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(azb_[a-z0-9_]+)\s*\(", text)))
    assert {"azb_nn_forward", "azb_nn_forward_tc", "azb_nn_forward_tc_debug", "azb_upload_pinned", "azb_nng_forward",
            "azb_nng_layout"} <= set(names)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libazb200.so lacks {missing}"


def test_binding_table_matches_header():
    """every symbol of both headers is bound in ONE table (no ad-hoc restype / argtypes elsewhere)"""
    assert sorted(_capi.SYMBOLS) == _declared(("azb200.h", "azb200_nn.h"))
    pkg = os.path.join(ROOT, "alphazero-general_b200", "azb200")
    for f in os.listdir(pkg):
        if f.endswith(".py") and f != "_capi.py":
            assert ".argtypes" not in open(os.path.join(pkg, f)).read(), f


def test_library_loads_and_reports_abi():
    lib = _capi.load()
    assert lib.azb_abi_version() == _capi.ABI_VERSION
    assert lib.azb_last_error() is not None


def test_config_struct_layout_matches_header():
    # the field order of azb_config in the header and in the ctypes mirror
    text = open(os.path.join(ROOT, "include", "azb200.h")).read()
    body = text[text.index("typedef struct azb_config {"):text.index("} azb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:const\s+)?(?:u?int\d+_t|float|double)\s*\*?\s*(\w+)\s*;", body)
    assert fields == [n for n, _ in _capi.AzbConfig._fields_]
    body = text[text.index("typedef struct azb_stats {"):text.index("} azb_stats;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\bint64_t\s+(\w+)\s*;", body)
    assert fields == [n for n, _ in _capi.AzbStats._fields_]


def test_null_engine_calls_fail_cleanly():
    lib = _capi.load()
    assert lib.azb_select(None, 0, 0, None) == -7          # AZB_ERR_BAD_ARGUMENT, no crash
    assert b"null" in lib.azb_last_error()
    assert lib.azb_destroy(None) == 0
