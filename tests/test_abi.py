"""C-ABI surface of libppgs_b200.so: the library loads on a CPU-only box, exports
every symbol include/ppgs_b200.h declares, and reports errors through status
codes + ppgs_last_error (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'ppgs_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ppgs_[a-z0-9_]+)\s*\(', text)))


def test_header_compiles_as_c():
    """The boundary is plain C: no C++ / torch types in the signatures."""
    subprocess.check_call(
        ['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', HEADER])


def test_exports_every_declared_symbol(library):
    lib = ctypes.CDLL(library.LIBRARY_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(names) == set(library.SYMBOLS), 'python binding and header disagree'


def test_default_config_matches_reference_constants(library):
    cfg = library.ModelConfig()
    library.lib.ppgs_default_config(ctypes.byref(cfg))
    # ppgs/config/defaults.py:127-161 + torch TransformerEncoderLayer defaults
    assert (cfg.input_channels, cfg.hidden_channels, cfg.num_layers, cfg.num_heads) == (80, 256, 5, 2)
    assert (cfg.ffn_channels, cfg.output_channels, cfg.kernel_size) == (2048, 40, 5)
    assert (cfg.chunk_length, cfg.chunk_overlap, cfg.max_len, cfg.is_causal) == (500, 50, 5000, 0)
    assert abs(cfg.layer_norm_eps - 1e-5) < 1e-12
    assert library.lib.ppgs_abi_version() == 1


def test_errors_are_status_codes_not_aborts(library):
    import torch
    cfg = library.ModelConfig()
    library.lib.ppgs_default_config(ctypes.byref(cfg))
    handle = ctypes.c_void_p()
    assert library.lib.ppgs_engine_create(None, 0, ctypes.byref(handle)) == library.E_INVALID
    assert 'NULL' in library.last_error()
    cfg.kernel_size = 4
    assert library.lib.ppgs_engine_create(ctypes.byref(cfg), 0, ctypes.byref(handle)) == library.E_INVALID
    cfg.kernel_size = 5
    if not torch.cuda.is_available():
        # fails loudly without a GPU: no CPU fallback exists
        assert library.lib.ppgs_engine_create(ctypes.byref(cfg), 0, ctypes.byref(handle)) == library.E_CUDA
        with pytest.raises(RuntimeError):
            library.check(library.E_CUDA)
    assert library.lib.ppgs_engine_finalize(None) == library.E_INVALID
    with pytest.raises(ValueError):
        library.check(library.E_INVALID)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ppgs_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'ppgs_b200')):
        for name in files:
            if name.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                text = open(os.path.join(dirpath, name)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, name
                assert '/root/reference' not in text, name


def test_no_cpu_fallback_in_python_host():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    import ppgs_b200
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ppgs_b200.from_audio(torch.zeros(1, 1, 16000), 16000)


def test_plain_c_consumer_links_and_runs(library, tmp_path):
    """tests/csrc/abi_smoke.c: a C program built against include/ppgs_b200.h and linked to the
    shared library runs the host-side entry points (no GPU) — the boundary is usable from C."""
    import torch
    from test_host_logic import write_wav
    lib_dir = os.path.dirname(library.LIBRARY_PATH)
    exe = tmp_path / 'abi_smoke'
    subprocess.check_call([
        'gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
        os.path.join(ROOT, 'tests', 'csrc', 'abi_smoke.c'), '-o', str(exe),
        '-L', lib_dir, '-lppgs_b200', f'-Wl,-rpath,{lib_dir}'])
    wav, out = tmp_path / 'a.wav', tmp_path / 'a.pt'
    raw = write_wav(wav, 6400 * 40, seed=3)
    result = subprocess.run([str(exe), str(wav), str(out)], check=True, capture_output=True, text=True)
    assert result.stdout.split() == ['256000', '16000', '1', '16', '0', '475', '160', '17', '1600']
    tensor = torch.load(out)
    assert tensor.shape == (40, 1600)
    expected = torch.from_numpy(raw[:, 0].astype('float32') / 32768.0)[:64000].reshape(40, 1600)
    assert torch.equal(tensor, expected)
