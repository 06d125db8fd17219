import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dropin')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def built():
    """Build the CUDA library and the CPU oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.fixture
def emulated_backend(monkeypatch):
    """TEST ONLY.  Route log_b200's ctypes calls to the CPU SIMT emulation of the real kernel source (tests/emu: the
    kernels and lgr_capi.cu compiled for the host, one fiber per CUDA thread) and let the host code accept CPU
    tensors, so that the host-side classes and the kernels' logic can be exercised under `-m "not gpu"`.  The product
    itself has no such path: log_b200 loads only its CUDA library and rejects CPU tensors."""
    import ctypes
    emu_dir = os.path.join(ROOT, 'tests', 'emu')
    if emu_dir not in sys.path:
        sys.path.insert(0, emu_dir)
    import build_emu
    import util
    from log_b200 import _capi
    lib = _capi.bind(ctypes.CDLL(build_emu.build()))
    monkeypatch.setattr(_capi, '_lib', lib)
    monkeypatch.setattr(_capi, 'current_stream', lambda device=None: None)
    monkeypatch.setattr(_capi, 'require_cuda', lambda t, name: None)
    monkeypatch.setattr(util, 'DEVICE', ['cpu'])
    return lib
