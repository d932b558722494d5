"""Shared test plumbing.  `-m "not gpu"` runs on the CPU-only dev container
(oracle vs golden vectors, host logic, C-ABI surface); `-m gpu` tests are the
parity tests proper and call the CUDA library through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


@pytest.fixture(scope='session')
def mel_emul_lib():
    """CPU replay of the mel kernel arithmetic (tests/csrc/mel_emul.cpp)."""
    import ctypes
    out = os.path.join(ROOT, 'tests', '_build', 'libmel_emul.so')
    src = os.path.join(ROOT, 'tests', 'csrc', 'mel_emul.cpp')
    hdr = os.path.join(ROOT, 'ppgs_b200', 'csrc', 'mel_math.cuh')
    if (not os.path.exists(out) or
            os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call([
            'g++', '-O2', '-ffp-contract=off', '-Wno-unknown-pragmas', '-shared', '-fPIC',
            '-o', out, src])
    return ctypes.CDLL(out)


@pytest.fixture(scope='session')
def library():
    """The built product library, via the package's own ctypes binding."""
    path = os.path.join(ROOT, 'ppgs_b200', 'lib', 'libppgs_b200.so')
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    from ppgs_b200 import _lib
    return _lib
