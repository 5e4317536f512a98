"""ctypes binding of libpml.so (include/pml.h).  This is the only place the package touches
native code.  There is no fallback: if the shared library is missing the first call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint32, c_uint64, c_void_p)

PML_ABI_VERSION = 4
PML_MAX_SOURCES = 8
PML_MAX_PASSES = 8
PML_MAX_SEGMENTS = 16
PML_FLAG_NO_SSIM = 1
PML_FLAG_NO_AUTOMASK = 2
PML_FLAG_AVG_REPROJ = 4
PML_FLAG_KERNEL_CTA = 256

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libpml.so")


class PmlPass(Structure):
    _fields_ = [("hd", c_int32), ("wd", c_int32), ("smooth_weight", c_float), ("reserved", c_int32),
                ("disp", c_void_p), ("smooth_color", c_void_p), ("noise", c_void_p),
                ("argmin", c_void_p), ("depth", c_void_p), ("warped", c_void_p), ("grad_disp", c_void_p),
                ("frame_weight", c_void_p), ("grad_frame_weight", c_void_p)]


class PmlSegments(Structure):
    _fields_ = [("n_seg", c_int32), ("seg_size", c_int32),
                ("target", c_void_p * PML_MAX_SEGMENTS),
                ("sources", (c_void_p * PML_MAX_SEGMENTS) * PML_MAX_SOURCES),
                ("K", c_void_p * PML_MAX_SEGMENTS), ("inv_K", c_void_p * PML_MAX_SEGMENTS),
                ("smooth_color", (c_void_p * PML_MAX_SEGMENTS) * PML_MAX_PASSES)]


class PmlProblem(Structure):
    _fields_ = [("B", c_int32), ("H", c_int32), ("W", c_int32), ("S", c_int32), ("n_pass", c_int32),
                ("flags", c_uint32), ("min_depth", c_float), ("max_depth", c_float), ("eps", c_float),
                ("reserved", c_int32), ("seed", c_uint64),
                ("target", c_void_p), ("sources", c_void_p * PML_MAX_SOURCES),
                ("K", c_void_p), ("inv_K", c_void_p), ("T", c_void_p * PML_MAX_SOURCES),
                ("passes", PmlPass * PML_MAX_PASSES),
                ("losses", c_void_p), ("grad_T", c_void_p), ("grad_disp_const", c_void_p),
                ("prof_start", c_void_p), ("prof_stop", c_void_p), ("loss_vector", c_void_p),
                ("loss_total", c_void_p), ("loss_total_div", c_float), ("reserved2", c_int32),
                ("segments", POINTER(PmlSegments)), ("seed_device", c_void_p)]


class PmlError(RuntimeError):
    pass


_SIGNATURES = {
    "pml_abi_version": (c_int, []),
    "pml_strerror": (c_char_p, [c_int]),
    "pml_workspace_bytes": (c_size_t, [POINTER(PmlProblem)]),
    "pml_loss_forward": (c_int, [POINTER(PmlProblem), c_void_p, c_size_t, c_void_p]),
    "pml_loss_forward_backward": (c_int, [POINTER(PmlProblem), c_void_p, c_size_t, c_void_p]),
    "pml_scale_grads": (c_int, [c_int32, c_int32, c_int32, POINTER(c_int32), POINTER(c_int32),
                                POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "pml_selection_masks": (c_int, [c_int32, c_int64, POINTER(c_void_p), c_int32, POINTER(c_void_p), c_void_p]),
    "pml_disp_to_depth_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "pml_disp_to_depth_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "pml_backproject_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pml_backproject_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pml_project_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "pml_project_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_int32, c_int32, c_int32, c_float, c_void_p]),
    "pml_project_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "pml_ssim_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pml_ssim_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pml_smooth_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pml_smooth_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pml_smooth_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "pml_pose_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "pml_pose_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "pml_upsample_fwd": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pml_upsample_bwd": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pml_bce_workspace_bytes": (c_size_t, []),
    "pml_bce_ones_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pml_bce_ones_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pml_disp_head_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pml_disp_head_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "pml_disp_head_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pml_pyramid_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "pml_pyramid_u8": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    "pml_depth_metrics_workspace_bytes": (c_size_t, []),
    "pml_depth_metrics_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int32] * 9 +
                                  [c_float, c_float, c_void_p]),
    "pml_depth_metrics_median_ratio": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pml_depth_metrics_reduce": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                         c_void_p, c_size_t, c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class Library:
    """A loaded libpml.so.  ``emulator=True`` is set only by the test-suite when it loads the
    host-thread emulation build of the same kernel sources (tests/emu); the product never does."""

    def __init__(self, path: str = DEFAULT_LIB, emulator: bool = False):
        if not os.path.isfile(path):
            raise PmlError(
                "libpml.so not found at %s -- build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback." % path)
        self.path = path
        self.emulator = emulator
        self.dll = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
        got = self.dll.pml_abi_version()
        if got != PML_ABI_VERSION:
            raise PmlError("libpml.so ABI %d != binding ABI %d" % (got, PML_ABI_VERSION))

    def check(self, status: int, what: str):
        if status != 0:
            msg = self.dll.pml_strerror(status)
            raise PmlError("%s failed: %s (status %d)" % (what, msg.decode() if msg else "?", status))

    def __getattr__(self, name):
        return getattr(self.dll, name)


_LIB = None


def get_library() -> Library:
    global _LIB
    if _LIB is None:
        _LIB = Library(os.environ.get("PML_LIBRARY", DEFAULT_LIB))
    return _LIB


def set_library_for_testing(lib):
    """Test hook: swap the library handle (used with the tests/emu build).  Returns the old one."""
    global _LIB
    old, _LIB = _LIB, lib
    return old
