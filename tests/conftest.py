import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_lib():
    """Host-thread emulation build of the kernel sources, loaded through the product's ctypes
    binding.  Test infrastructure: exercises kernel logic + host code without a GPU."""
    import __graft_entry__ as ge
    from ssde_b200 import _cabi
    path = ge.build_emu()
    lib = _cabi.Library(path, emulator=True)
    old = _cabi.set_library_for_testing(lib)
    yield lib
    _cabi.set_library_for_testing(old)


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ssde_b200 import _cabi
    _cabi.set_library_for_testing(None)
    return _cabi.get_library()
