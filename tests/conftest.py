import ctypes
import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
REF_DIR = ROOT / "oracle" / "_ref"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_ref_iou3d():
    """The reference's own iou3d_nms_cuda pybind module, built unmodified by oracle/build_ref.py."""
    path = REF_DIR / "iou3d_nms_cuda.so"
    if not path.exists():
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("iou3d_nms_cuda", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_pointnet2():
    """ctypes handle on the reference pointnet2_batch kernels behind oracle/ref_binding.cu."""
    path = REF_DIR / "libpdanet_ref_pointnet2.so"
    if not path.exists():
        return None
    lib = ctypes.CDLL(str(path))
    return lib


@pytest.fixture(scope="session")
def ref_iou3d():
    mod = load_ref_iou3d()
    if mod is None:
        pytest.skip("oracle/_ref/iou3d_nms_cuda.so not built (reference tree unavailable)")
    return mod


@pytest.fixture(scope="session")
def ref_pointnet2():
    lib = load_ref_pointnet2()
    if lib is None:
        pytest.skip("oracle/_ref/libpdanet_ref_pointnet2.so not built (reference tree unavailable)")
    return lib
