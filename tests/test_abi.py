"""C-ABI checks that need no GPU: libpdab.so loads and exports exactly what include/pdab.h declares."""
import ctypes
import re

from conftest import ROOT

from pdanet_b200 import _lib


def header_symbols():
    text = (ROOT / "include" / "pdab.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdab_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_every_python_binding():
    assert sorted(_lib.EXPORTED_SYMBOLS) == header_symbols()


def test_library_exports_every_declared_symbol():
    _lib.build()
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in header_symbols():
        assert hasattr(handle, name), f"{name} declared in include/pdab.h but not exported by libpdab.so"


def test_version_and_error_strings():
    lib = _lib.lib()
    assert b"sm_100a" in lib.pdab_version()
    assert lib.pdab_error_string(0) == b"success"
    assert b"invalid argument" in lib.pdab_error_string(-1)
    assert b"range" in lib.pdab_error_string(-2)


def test_argument_validation_happens_before_any_cuda_call():
    lib = _lib.lib()
    # null pointers / negative sizes -> PDAB_EINVAL, empty problems -> success; no launch in either case
    assert lib.pdab_fps(1, 16, 4, None, None, None, None) == -1
    assert lib.pdab_ball_query(1, 16, 4, 0.5, 0, None, None, None, None) == -1
    assert lib.pdab_group_points(-1, 1, 1, 1, 1, None, None, None, None) == -1
    assert lib.pdab_topk_ctr(1, 8, 3, 9, None, None, None) == -1
    assert lib.pdab_nms_workspace_bytes(256) == 256 * 4 * 8
    assert lib.pdab_nms_workspace_bytes(0) == 0
    assert lib.pdab_nms_host(None, 0, 0.1, None, 0, None) == 0
