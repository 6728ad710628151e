"""The C-ABI shared library loads and exports every symbol include/tgtk_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tgtk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tgtk_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from tumorgrowthtoolkit_b200 import _native
    return ctypes.CDLL(_native.LIB_PATH)


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for must in ("tgtk_sim_create", "tgtk_sim_run", "tgtk_sim_upload", "tgtk_sim_download", "tgtk_fk_update_host",
                 "tgtk_fk2c_update_host", "tgtk_faces_host", "tgtk_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    for name in _declared_symbols():
        assert hasattr(built_lib, name), f"libtgtk_b200.so lacks {name}"


def test_abi_version_and_error_string(built_lib):
    assert built_lib.tgtk_abi_version() == 1
    built_lib.tgtk_last_error.restype = ctypes.c_char_p
    assert isinstance(built_lib.tgtk_last_error(), bytes)


def test_param_struct_layout_matches_header():
    from tumorgrowthtoolkit_b200 import _native
    # 6 int32 + 15 doubles (h[3] + 12 scalars), no padding surprises
    assert ctypes.sizeof(_native.Params) == 6 * 4 + 15 * 8
    assert ctypes.sizeof(_native.Status) == 4 * 4 + 2 * 8 + 8


def test_missing_library_fails_loudly(monkeypatch):
    from tumorgrowthtoolkit_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libtgtk_b200.so")
    with pytest.raises(_native.NativeError):
        _native.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tumorgrowthtoolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"
