"""Host half of the ingest / egress C ABI (no GPU): the WAVE probe and decoder against
scipy's reader with torchaudio.load's normalisation, the `.pt` writer against torch.load,
and the resampler's filter table against torchaudio's own (bit-exact)."""
import ctypes
import math
import struct
import zipfile

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from test_host_logic import write_wav


def probe(lib, path):
    frames, rate = ctypes.c_int64(), ctypes.c_int()
    channels, bits, is_float = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    code = lib.lib.ppgs_wav_info(str(path).encode(), frames, rate, channels, bits, is_float)
    return code, (frames.value, rate.value, channels.value, bits.value, is_float.value)


def decode(lib, path, capacity):
    buffer = np.empty(capacity, np.float32)
    frames, rate = ctypes.c_int64(), ctypes.c_int()
    lib.check(lib.lib.ppgs_wav_read_f32(
        str(path).encode(), buffer.ctypes.data, capacity, frames, rate))
    return buffer[:frames.value], rate.value


def normalised(data):
    """torchaudio.load(normalize=True) of integer PCM."""
    if data.dtype == np.int16:
        return data.astype(np.float32) / 32768.0
    if data.dtype == np.int32:
        return data.astype(np.float32) / 2147483648.0
    if data.dtype == np.uint8:
        return (data.astype(np.float32) - 128.0) / 128.0
    return data.astype(np.float32)


@pytest.mark.parametrize('kind,rate', [
    ('int16', 16000), ('int16-stereo', 44100), ('int32', 8000), ('uint8', 16000),
    ('float32', 22050), ('float64', 16000)])
def test_wav_probe_and_decode_match_scipy(library, tmp_path, kind, rate):
    rng = np.random.default_rng(3)
    n = 1237
    data = {
        'int16': lambda: rng.integers(-32768, 32768, n).astype(np.int16),
        'int16-stereo': lambda: rng.integers(-32768, 32768, (n, 2)).astype(np.int16),
        'int32': lambda: rng.integers(-2**31, 2**31, n).astype(np.int32),
        'uint8': lambda: rng.integers(0, 256, n).astype(np.uint8),
        'float32': lambda: rng.standard_normal(n).astype(np.float32),
        'float64': lambda: rng.standard_normal(n),
    }[kind]()
    path = tmp_path / f'{kind}.wav'
    wavfile.write(path, rate, data)
    code, (frames, got_rate, channels, bits, is_float) = probe(library, path)
    assert code == 0
    assert (frames, got_rate) == (n, rate)
    assert channels == (2 if data.ndim == 2 else 1)
    assert bits == data.dtype.itemsize * 8 and is_float == int(data.dtype.kind == 'f')
    audio, got_rate = decode(library, path, n)
    channel0 = data if data.ndim == 1 else data[:, 0]
    assert got_rate == rate and np.array_equal(audio, normalised(channel0))


def test_wav_24bit_extensible_and_extra_chunks(library, tmp_path):
    """24-bit PCM in a WAVE_FORMAT_EXTENSIBLE header, a LIST chunk of odd size before
    `data` and trailing bytes after it."""
    rng = np.random.default_rng(5)
    values = rng.integers(-2**23, 2**23, 333)
    payload = b''.join(struct.pack('<i', int(v))[:3] for v in values)
    fmt = struct.pack('<HHIIHH', 0xFFFE, 1, 16000, 48000, 3, 24) + struct.pack('<HHI', 22, 24, 4)
    fmt += struct.pack('<H', 1) + b'\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71'
    body = b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt
    body += b'LIST' + struct.pack('<I', 5) + b'hello' + b'\x00'
    body += b'data' + struct.pack('<I', len(payload)) + payload + b'\x00tail'
    path = tmp_path / 'x.wav'
    path.write_bytes(b'RIFF' + struct.pack('<I', len(body)) + body)
    code, (frames, rate, channels, bits, is_float) = probe(library, path)
    assert (code, frames, rate, channels, bits, is_float) == (0, 333, 16000, 1, 24, 0)
    audio, _ = decode(library, path, 400)
    assert np.array_equal(audio, (values * 256).astype(np.float32) / 2147483648.0)


def test_wav_errors(library, tmp_path):
    assert probe(library, tmp_path / 'missing.wav')[0] == library.E_INVALID
    assert 'cannot open' in library.last_error()
    other = tmp_path / 'not.wav'
    other.write_bytes(b'ID3\x03' + bytes(64))
    assert probe(library, other)[0] == library.E_UNSUPPORTED
    short = tmp_path / 'short.wav'
    write_wav(short, 100)
    with pytest.raises(ValueError, match='do not fit'):
        decode(library, short, 50)
    from ppgs_b200 import load
    assert load.wav_info(other) is None
    assert load.wav_info(short) == {
        'samples': 100, 'sample_rate': 16000, 'channels': 1, 'bits': 16, 'is_float': False}


@pytest.mark.parametrize('cols', [0, 1, 255, 256, 1000, 1234])
def test_pt_writer_is_torch_loadable(library, tmp_path, cols):
    """ppgs_pt_write_f32 == torch.save(tensor[..., :length].clone()) for what torch.load
    returns (ppgs/preprocess/core.py:219-221), also with weights_only and mmap."""
    x = torch.randn(40, 1234)
    path = tmp_path / f'ppg-{cols}.pt'
    library.check(library.lib.ppgs_pt_write_f32(str(path).encode(), x.data_ptr(), 40, cols, 1234))
    for kwargs in ({}, {'weights_only': True}, {'weights_only': False}, {'mmap': True}):
        y = torch.load(path, **kwargs)
        assert y.dtype == torch.float32 and y.shape == (40, cols) and y.is_contiguous()
        assert torch.equal(y, x[:, :cols])
    archive = zipfile.ZipFile(path)
    assert archive.testzip() is None
    names = archive.namelist()
    assert names == [f'ppg-{cols}/data.pkl', f'ppg-{cols}/byteorder', f'ppg-{cols}/data/0',
                     f'ppg-{cols}/version']
    info = archive.getinfo(f'ppg-{cols}/data/0')
    # storage starts 64-byte aligned, like torch's own writer (mmap-friendly)
    with open(path, 'rb') as f:
        f.seek(info.header_offset + 26)
        name_len, extra_len = struct.unpack('<HH', f.read(4))
    assert (info.header_offset + 30 + name_len + extra_len) % 64 == 0
    # and a tensor saved by torch round-trips to the same values
    reference = tmp_path / 'reference.pt'
    torch.save(x[..., :cols].clone(), reference)
    assert torch.equal(torch.load(reference), torch.load(path))


def test_pt_writer_large_and_errors(library, tmp_path):
    big = torch.randn(40, 70000)
    path = tmp_path / 'big.pt'
    library.check(library.lib.ppgs_pt_write_f32(str(path).encode(), big.data_ptr(), 40, 70000, 70000))
    assert torch.equal(torch.load(path), big)
    with pytest.raises(ValueError, match='bad shape'):
        library.check(library.lib.ppgs_pt_write_f32(str(path).encode(), big.data_ptr(), 40, 10, 5))
    with pytest.raises(ValueError, match='cannot open'):
        library.check(library.lib.ppgs_pt_write_f32(
            str(tmp_path / 'no' / 'dir.pt').encode(), big.data_ptr(), 40, 10, 10))


@pytest.mark.parametrize('orig,new', [
    (44100, 16000), (22050, 16000), (8000, 16000), (48000, 16000), (11025, 16000),
    (32000, 16000), (24000, 16000), (16001, 16000)])
def test_resample_taps_equal_torchaudio(library, orig, new):
    """The filter bank of torchaudio.transforms.Resample(orig, new) (defaults), bit for bit."""
    import torchaudio.functional as F
    kernel, width = F.functional._get_sinc_resample_kernel(orig, new, math.gcd(orig, new))
    ntaps, phases, got_width = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    library.check(library.lib.ppgs_resample_taps(orig, new, None, 0, ntaps, phases, got_width))
    assert (phases.value, ntaps.value, got_width.value) == (kernel.shape[0], kernel.shape[2], width)
    taps = np.empty((ntaps.value, phases.value), np.float32)
    library.check(library.lib.ppgs_resample_taps(
        orig, new, taps.ctypes.data, taps.size, ntaps, phases, got_width))
    assert np.array_equal(taps, kernel[:, 0, :].numpy().T)
    for samples in (0, 1, 159, 12345, 160000):
        expected = math.ceil(new * samples / orig)
        assert library.lib.ppgs_resample_length(samples, orig, new) == expected


def test_metadata_flags_native_eligibility(tmp_path):
    from ppgs_b200 import data
    a, b, c = tmp_path / 'a.wav', tmp_path / 'b.wav', tmp_path / 'c.wav'
    write_wav(a, 16000)
    write_wav(b, 8000, channels=2)
    assert data.Metadata([a, b]).native
    assert data.Metadata([a, b]).samples == {a: 16000, b: 8000}
    write_wav(c, 22050, rate=22050)
    metadata = data.Metadata([a, c])
    assert not metadata.native and metadata.lengths == [100, 100]
    wavfile.write(c, 16000, np.zeros(1600, np.float32))
    assert not data.Metadata([a, c]).native


def test_wav_info_many_matches_single_probe(library, tmp_path):
    from ppgs_b200 import load
    files = []
    for i in range(130):      # > 64: the threaded path
        files.append(tmp_path / f'{i}.wav')
        write_wav(files[-1], 100 + i, rate=16000 if i % 3 else 8000, channels=1 + i % 2)
    other = tmp_path / 'x.mp3'
    other.write_bytes(b'ID3' + bytes(40))
    files.insert(7, other)
    infos = load.wav_info_many(files, threads=8)
    assert infos[7] is None
    for file, info in zip(files, infos):
        assert info == load.wav_info(file)
    with pytest.raises(ValueError, match='cannot open'):
        load.wav_info_many(files + [tmp_path / 'missing.wav'])
    assert load.wav_info_many([]) == []


def test_pt_writer_fp16_and_save_masked(library, tmp_path):
    """fp16 feature tensors (torch.HalfStorage) and the save_masked wrapper."""
    from ppgs_b200 import preprocess
    x = torch.randn(80, 333).half()
    for cols in (0, 1, 100, 333):
        path = tmp_path / f'mel-{cols}.pt'
        library.check(library.lib.ppgs_pt_write_f16(str(path).encode(), x.data_ptr(), 80, cols, 333))
        for kwargs in ({}, {'weights_only': True}):
            y = torch.load(path, **kwargs)
            assert y.dtype == torch.float16 and y.shape == (80, cols) and torch.equal(y, x[:, :cols])
    for tensor in (x, x.float(), torch.randn(2, 3, 10), torch.arange(12).reshape(3, 4)):
        path = tmp_path / 'masked.pt'
        preprocess.save_masked(tensor, path, 3)         # native for 2-D fp16 / fp32, torch.save otherwise
        assert torch.equal(torch.load(path), tensor[..., :3])


def test_native_pt_reader_roundtrip_and_torch_files(tmp_path):
    """ppgs_pt_info / ppgs_pt_read (the feature-cache reader, ppgs/data/dataset.py:98-101): files
    written by torch.save itself (fp16 / fp32, 1-D .. 3-D) and by the library's own writer come
    back bit for bit, also into a row of a padded batch; anything else reports 'unsupported'
    through tensor_info() -> None so that callers fall back to torch.load."""
    import torch
    from ppgs_b200 import load, preprocess
    g = torch.Generator().manual_seed(0)
    cases = {
        'mel': torch.randn(80, 333, generator=g).half(),
        'ppg': torch.rand(40, 1000, generator=g),
        'vec': torch.randn(17, generator=g),
        'cube': torch.randn(2, 3, 5, generator=g).half(),
    }
    for name, tensor in cases.items():
        file = tmp_path / f'{name}.pt'
        torch.save(tensor, file)
        shape, dtype = load.tensor_info(file)
        assert shape == tuple(tensor.shape) and dtype == tensor.dtype
        assert torch.equal(load.features(file), tensor)
    # the library's own writer (save_masked crops a padded row)
    padded = torch.zeros(80, 400, dtype=torch.float16)
    padded[:, :333] = cases['mel']
    preprocess.save_masked(padded, tmp_path / 'own.pt', 333)
    assert load.tensor_info(tmp_path / 'own.pt') == ((80, 333), torch.float16)
    assert torch.equal(load.features(tmp_path / 'own.pt'), cases['mel'])
    # straight into a row of a padded batch
    batch = torch.zeros(3, 80, 512, dtype=torch.float16)
    view = load.features(tmp_path / 'mel.pt', out=batch[1])
    assert view.shape == (80, 333) and torch.equal(batch[1, :, :333], cases['mel'])
    assert not batch[1, :, 333:].any() and not batch[0].any() and not batch[2].any()
    # not covered natively: views, other dtypes, dicts -> None, and features() still loads them
    torch.save(cases['ppg'][:, 10:20], tmp_path / 'view.pt')     # storage larger than the tensor
    torch.save(torch.arange(6).reshape(2, 3), tmp_path / 'int.pt')
    torch.save({'model': cases['vec']}, tmp_path / 'dict.pt')
    for name in ('view', 'int', 'dict'):
        assert load.tensor_info(tmp_path / f'{name}.pt') is None
    assert torch.equal(load.features(tmp_path / 'view.pt'), cases['ppg'][:, 10:20])


def test_damaged_wav_and_pt_files_never_crash_the_readers(tmp_path):
    """Untrusted files: random damage / truncation / insertions into valid WAVE and `.pt` files end in
    a status code (or a successful read of in-range data), never in a crash or a write past the buffer."""
    import ctypes
    import io
    import random
    from scipy.io import wavfile
    from ppgs_b200 import _lib, load
    random.seed(1)

    def damaged(pool, head=None):
        data = bytearray(random.choice(pool))
        span = min(len(data), head) if head else len(data)
        mode = random.random()
        if mode < 0.6:
            for _ in range(random.randint(1, 4)):
                data[random.randrange(span)] = random.randrange(256)
        elif mode < 0.8:
            data = data[:random.randrange(1, len(data))]
        else:
            at = random.randrange(span)
            data[at:at] = bytes(random.randrange(256) for _ in range(random.randint(1, 8)))
        return bytes(data)

    wavs = []
    for array, rate in ((np.random.default_rng(0).integers(-3000, 3000, 5000).astype(np.int16), 16000),
                        (np.random.default_rng(1).standard_normal((3000, 2)).astype(np.float32), 22050)):
        buffer = io.BytesIO()
        wavfile.write(buffer, rate, array)
        wavs.append(buffer.getvalue())
    out = torch.full((2, 6000), 7.0)                 # row 1 is the guard
    path = tmp_path / 'fuzz.wav'
    for _ in range(800):
        path.write_bytes(damaged(wavs, head=80))
        frames, rate = ctypes.c_int64(), ctypes.c_int()
        _lib.lib.ppgs_wav_read_f32(str(path).encode(), ctypes.c_void_p(out.data_ptr()), 6000,
                                   ctypes.byref(frames), ctypes.byref(rate))
        assert bool((out[1] == 7.0).all())

    tensors = []
    for tensor in (torch.randn(80, 137).half(), torch.randn(40, 55), torch.randn(2, 3, 17).half()):
        buffer = io.BytesIO()
        torch.save(tensor, buffer)
        tensors.append(buffer.getvalue())
    path = tmp_path / 'fuzz.pt'
    for _ in range(800):
        path.write_bytes(damaged(tensors))
        info = load.tensor_info(path)
        if info is None or len(info[0]) < 2:
            continue
        shape, dtype = info
        rows, cols = int(np.prod(shape[:-1])), int(shape[-1])
        if not 0 < rows * cols <= 1 << 16:
            continue
        buffer = torch.full((rows * cols + 64,), 7.0, dtype=dtype)
        _lib.lib.ppgs_pt_read(str(path).encode(), ctypes.c_void_p(buffer.data_ptr()), rows, cols,
                              2 if dtype == torch.float16 else 4, cols)
        assert bool((buffer[rows * cols:] == 7.0).all())
