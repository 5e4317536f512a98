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
    assert ctypes.sizeof(_cabi.PmlProblem) == _cabi.PmlProblem.passes.offset + 8 * ctypes.sizeof(_cabi.PmlPass) + 48 + 32


def test_argument_validation_without_gpu(lib):
    p = _cabi.PmlProblem()
    assert lib.dll.pml_workspace_bytes(ctypes.byref(p)) == 0
    assert lib.dll.pml_loss_forward(ctypes.byref(p), None, 0, None) == -1          # PML_ERR_INVALID
    p.B, p.H, p.W, p.S, p.n_pass = 1, 32, 64, 9, 1
    assert lib.dll.pml_loss_forward(ctypes.byref(p), None, 0, None) == -2          # too many sources
    assert lib.dll.pml_ssim_fwd(None, None, None, 1, 8, 8, None) == -1
    assert lib.dll.pml_pose_fwd(None, None, None, 1, 0, None) == -1
    assert lib.dll.pml_smooth_fwd(None, None, None, None, 0, 1, 3, 8, 8, None) == -1
    # entry points added for the predictive mask and the input pyramid
    assert lib.dll.pml_upsample_fwd(None, None, 1, 4, 4, 8, 8, None) == -1
    assert lib.dll.pml_upsample_bwd(None, None, 1, 4, 4, 8, 8, None) == -1
    assert lib.dll.pml_bce_ones_fwd(None, 16, None, None, 0, None) == -1
    assert lib.dll.pml_bce_ones_bwd(None, None, None, 16, None) == -1
    assert lib.dll.pml_bce_workspace_bytes() > 0
    assert lib.dll.pml_pyramid_u8(None, 1, 32, 64, 4, None, None, 0, None) == -1
    assert lib.dll.pml_pyramid_workspace_bytes(2, 32, 64, 4) >= 2 * (16 * 32 + 8 * 16) * 3
    # a frame weight (predictive mask) without PML_FLAG_NO_AUTOMASK is refused (trainer.py:556 / :571)
    buf = (ctypes.c_float * 4)()
    q = _cabi.PmlProblem()
    q.B, q.H, q.W, q.S, q.n_pass = 1, 32, 64, 2, 1
    addr = ctypes.addressof(buf)
    q.target = q.K = q.inv_K = q.losses = addr
    q.sources[0] = q.sources[1] = q.T[0] = q.T[1] = addr
    q.passes[0].hd, q.passes[0].wd = 32, 64
    q.passes[0].disp = q.passes[0].smooth_color = q.passes[0].frame_weight = addr
    assert lib.dll.pml_workspace_bytes(ctypes.byref(q)) == 0
    q.flags = _cabi.PML_FLAG_NO_AUTOMASK
    assert lib.dll.pml_workspace_bytes(ctypes.byref(q)) > 0


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.PmlError, match="no CPU or PyTorch fallback"):
        _cabi.Library(str(tmp_path / "nope.so"))
