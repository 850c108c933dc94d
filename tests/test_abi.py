"""The C-ABI library loads on a CPU-only box and exports every symbol include/abl_b200.h declares."""
import ctypes
import os
import re

from ab_line_classifier_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "abl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abl_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for s in ["abl_conv3x3_igemm", "abl_conv3x3_wgrad", "abl_conv_first_fwd", "abl_head_fwd", "abl_adam_flat",
              "abl_rmsprop_flat", "abl_gradcam", "abl_last_error"]:
        assert s in syms


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/abl_b200.h but not exported"


def test_ctypes_binding_covers_every_header_symbol():
    assert sorted(_native.SIGNATURES) == _header_symbols()


def test_loads_without_a_gpu_and_reports_version():
    lib = _native.load()
    assert lib.abl_abi_version() == 8
    assert lib.abl_last_error() is not None


def test_no_link_time_dependency_on_libcuda():
    import subprocess
    out = subprocess.run(["ldd", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out
