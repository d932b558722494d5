"""The mel kernel's per-frame arithmetic (ppgs_b200/csrc/mel_math.cuh, shared
verbatim by the CUDA kernel) replayed on the CPU against the reference's golden
mels.  Tolerance: <= 1 fp16 ulp, mismatch rate <= 1e-3 (the kernel runs its own
radix-8 Stockham FFT, torch.stft runs pocketfft: same values, different rounding
order)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ppg_oracle as O
from test_oracle_golden import MEL, mel_case_audio


def kernel_tables():
    """Host tables exactly as ppgs_b200/csrc/mel.cu:build_mel_tables builds them."""
    basis = O.mel_basis()
    window = torch.hann_window(1024, dtype=torch.float32).numpy()
    a = -2 * np.pi * np.arange(512) / 512
    tw512 = np.stack([np.cos(a), np.sin(a)], 1).astype(np.float32)
    a = -2 * np.pi * np.arange(513) / 1024
    tw1024 = np.stack([np.cos(a), np.sin(a)], 1).astype(np.float32)
    return window, tw512, tw1024, np.ascontiguousarray(basis, dtype=np.float32)


def emulate(lib, audio):
    batch, _, samples = audio.shape
    flat = np.ascontiguousarray(audio.numpy().reshape(batch, samples))
    out = np.zeros((batch, 80, samples // 160), np.uint16)
    tables = kernel_tables()
    ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib.mel_emul(ptr(flat), batch, ctypes.c_long(samples), *[ptr(t) for t in tables], ptr(out))
    return out.view(np.float16)


def ulp_distance(a, b):
    to_ordered = lambda x: np.where(x < 0, -(x & 0x7fff), x).astype(np.int32)  # noqa: E731
    return np.abs(to_ordered(a.view(np.int16)) - to_ordered(b.view(np.int16)))


@pytest.mark.parametrize('name', MEL)
def test_kernel_arithmetic_vs_reference_mel(mel_emul_lib, name):
    g = golden(name)
    mel = emulate(mel_emul_lib, mel_case_audio(g))
    dist = ulp_distance(mel, g['mel'])
    assert dist.max() <= 1
    assert (dist > 0).mean() <= 1e-3


def test_slaney_filterbank_fits_kernel_table(mel_emul_lib):
    """The Slaney basis cut into pieces of 7 bins: 6 rounds of 32 pieces (kFbMaxRounds = 9 in
    mel_math.cuh); a dense basis is refused."""
    *_, basis = kernel_tables()
    ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert 0 < mel_emul_lib.mel_filterbank_rounds(ptr(basis)) <= 9
    assert mel_emul_lib.mel_filterbank_rounds(ptr(np.ones((80, 513), np.float32))) == -1
