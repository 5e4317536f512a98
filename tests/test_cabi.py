"""The C-ABI library loads and exports every symbol include/pml.h declares; argument validation
happens before any CUDA call, so it is checkable without a GPU."""
import ctypes
import os
import re

import pytest

import common  # noqa: F401  (sys.path)
from ssde_b200 import _cabi

HEADER = os.path.join(common.ROOT, "include", "pml.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_cabi.DEFAULT_LIB):
        import __graft_entry__ as ge
        ge.build()
    return _cabi.Library(_cabi.DEFAULT_LIB)


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pml_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_cabi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib.dll, name), name
    assert lib.dll.pml_abi_version() == _cabi.PML_ABI_VERSION
    assert lib.dll.pml_strerror(0) == b"ok"
    assert b"workspace" in lib.dll.pml_strerror(-3)


def test_struct_layout_matches_header():
    # offsets implied by include/pml.h on LP64
    assert ctypes.sizeof(_cabi.PmlPass) == 16 + 9 * 8
    assert _cabi.PmlProblem.seed.offset == 40
    assert _cabi.PmlProblem.target.offset == 48
    assert _cabi.PmlProblem.passes.offset == 48 + 8 * (1 + 8 + 2 + 8)
    assert ctypes.sizeof(_cabi.PmlProblem) == _cabi.PmlProblem.passes.offset + 8 * ctypes.sizeof(_cabi.PmlPass) + 48


def test_argument_validation_without_gpu(lib):
    p = _cabi.PmlProblem()
    assert lib.dll.pml_workspace_bytes(ctypes.byref(p)) == 0
    assert lib.dll.pml_loss_forward(ctypes.byref(p), None, 0, None) == -1          # PML_ERR_INVALID
    p.B, p.H, p.W, p.S, p.n_pass = 1, 32, 64, 9, 1
    assert lib.dll.pml_loss_forward(ctypes.byref(p), None, 0, None) == -2          # too many sources
    assert lib.dll.pml_ssim_fwd(None, None, None, 1, 8, 8, None) == -1
    assert lib.dll.pml_pose_fwd(None, None, None, 1, 0, None) == -1
    assert lib.dll.pml_smooth_fwd(None, None, None, None, 0, 1, 3, 8, 8, None) == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.PmlError, match="no CPU or PyTorch fallback"):
        _cabi.Library(str(tmp_path / "nope.so"))
