"""FLAC ingest of ppgs.load.audio (ppgs/load.py:17-30): the native decoder
against (1) the format specification's own example stream (RFC 9639, appendix D.1 — an external
known answer: its frame CRCs and STREAMINFO MD5 were produced by the reference encoder) and
(2) streams written by the test-side encoder tests/flac_writer.py on seeded signals, one case
per decoder branch.  Bit-exact: sample / 2^(bits-1) is exact in fp32 up to 24 bits."""
import ctypes

import numpy as np
import pytest
import torch

import flac_writer as F
from ppgs_b200 import _lib, load

RFC_EXAMPLE_1 = bytes.fromhex(
    '664c614380000022100010000000' '0f00000f0ac442f0000000013e84'
    'b41807dc690307586a3dad1a2e0f' 'fff869180000bf0358fd03128baa9a')


def decode(path, expect_ok=True):
    info = load.flac_info(path)
    out = torch.empty(info['channels'], info['samples'], dtype=torch.float32)
    frames, rate, channels = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib.ppgs_flac_read_f32(
        str(path).encode(), ctypes.c_void_p(out.data_ptr()), info['samples'],
        ctypes.byref(frames), ctypes.byref(rate), ctypes.byref(channels)))
    assert (frames.value, rate.value, channels.value) == (info['samples'], info['sample_rate'], info['channels'])
    return out.numpy(), info


def signal(channels, count, bits, seed, kind='speech'):
    rng = np.random.default_rng(seed)
    t = np.arange(count)
    peak = (1 << (bits - 1)) - 1
    rows = []
    for c in range(channels):
        if kind == 'noise':
            x = rng.integers(-peak - 1, peak + 1, count)
        else:
            x = 0.4 * np.sin(2 * np.pi * (110 + 37 * c) * t / 16000) + 0.2 * np.sin(2 * np.pi * 1234 * t / 16000)
            x = x + 0.02 * rng.standard_normal(count) + (0.15 * rows[0] / peak if rows else 0)
            x = np.clip(np.round(x * peak), -peak - 1, peak)
        rows.append(x.astype(np.int64))
    return np.stack(rows)


def roundtrip(tmp_path, samples, rate, bits, **options):
    path = tmp_path / 'x.flac'
    path.write_bytes(F.encode(samples, rate, bits, **options))
    out, info = decode(path)
    assert info == {'samples': samples.shape[1], 'sample_rate': rate, 'channels': samples.shape[0], 'bits': bits}
    expect = (samples.astype(np.float64) / float(1 << (bits - 1))).astype(np.float32)
    assert np.array_equal(out, expect)
    return path


def test_rfc9639_example_stream(tmp_path):
    path = tmp_path / 'rfc.flac'
    path.write_bytes(RFC_EXAMPLE_1)
    out, info = decode(path)
    assert info == {'samples': 1, 'sample_rate': 44100, 'channels': 2, 'bits': 16}
    assert (out * 32768).tolist() == [[25588.0], [10416.0]]


SUBFRAMES = {
    'verbatim': [dict(kind='verbatim')],
    'fixed0': [dict(kind='fixed', order=0)],
    'fixed1': [dict(kind='fixed', order=1, partition_order=2)],
    'fixed2': [dict(kind='fixed', order=2, partition_order=3)],
    'fixed3': [dict(kind='fixed', order=3, method=1)],
    'fixed4': [dict(kind='fixed', order=4, partition_order=4, method=1)],
    'lpc1': [dict(kind='lpc', order=1, precision=5)],
    'lpc8': [dict(kind='lpc', order=8, precision=12, partition_order=3)],
    'lpc12_15bit': [dict(kind='lpc', order=12, precision=15, partition_order=2)],
    'lpc32': [dict(kind='lpc', order=32, precision=14, partition_order=1)],
    'escaped': [dict(kind='fixed', order=2, partition_order=2, escape=(0, 2))],
    'escaped_method1': [dict(kind='lpc', order=4, partition_order=1, method=1, escape=(1,))],
    'mixed_per_channel': [dict(kind='lpc', order=6), dict(kind='fixed', order=3, partition_order=1)],
}


@pytest.mark.parametrize('name', sorted(SUBFRAMES))
def test_subframe_types(tmp_path, name):
    roundtrip(tmp_path, signal(2, 3 * 1024 + 100, 16, 1), 16000, 16, blocks=1024, specs=SUBFRAMES[name])


@pytest.mark.parametrize('stereo', ['left_side', 'right_side', 'mid_side'])
@pytest.mark.parametrize('bits', [16, 24])
def test_stereo_decorrelation(tmp_path, stereo, bits):
    roundtrip(tmp_path, signal(2, 5000, bits, 2), 44100, bits, blocks=4096, stereo=stereo,
              specs=[dict(kind='lpc', order=8, partition_order=2)])


@pytest.mark.parametrize('bits,kind', [(8, 'speech'), (12, 'speech'), (20, 'speech'), (24, 'noise'),
                                        (32, 'speech'), (16, 'noise')])
def test_sample_sizes(tmp_path, bits, kind):
    if bits == 32:      # fp32 cannot hold 32-bit samples exactly: compare with fp32 rounding
        samples = signal(1, 2000, 32, 3)
        path = tmp_path / 'x.flac'
        path.write_bytes(F.encode(samples, 48000, 32, blocks=576, specs=[dict(kind='fixed', order=2)]))
        out, _ = decode(path)
        assert np.array_equal(out, samples.astype(np.float32) * np.float32(2.0 ** -31))
        return
    roundtrip(tmp_path, signal(2, 2500, bits, 3, kind), 48000, bits, blocks=1152,
              specs=[dict(kind='fixed', order=1 if kind == 'noise' else 3, method=1)])


def test_32_bit_side_channel(tmp_path):
    samples = signal(2, 700, 32, 4)
    samples[0, :10] = (1 << 31) - 1
    samples[1, :10] = -(1 << 31)          # side needs the 33rd bit
    path = tmp_path / 'x.flac'
    path.write_bytes(F.encode(samples, 48000, 32, blocks=256, stereo='left_side', specs=[dict(kind='verbatim')]))
    out, _ = decode(path)
    assert np.array_equal(out, samples.astype(np.float32) * np.float32(2.0 ** -31))


def test_constant_and_wasted_bits(tmp_path):
    samples = signal(2, 2048, 16, 5)
    samples[0, :1024] = -1234                       # CONSTANT subframe in frame 0
    samples[1] = (samples[1] >> 3) << 3             # three wasted bits

    def specs(index):
        first = dict(kind='constant') if index == 0 else dict(kind='fixed', order=2)
        return [first, dict(kind='lpc', order=4, wasted=True)]
    roundtrip(tmp_path, samples, 16000, 16, blocks=1024, specs=specs)
    silence = np.zeros((1, 4000), dtype=np.int64)
    roundtrip(tmp_path, silence, 16000, 16, blocks=4096, specs=[dict(kind='constant', wasted=True)])


@pytest.mark.parametrize('channels', [1, 3, 8])
def test_channel_counts(tmp_path, channels):
    roundtrip(tmp_path, signal(channels, 1500, 16, 6), 22050, 16, blocks=512)


def test_block_size_and_rate_codes(tmp_path):
    samples = signal(1, 9000, 16, 7)
    for blocks in (192, 576, 255, 256, 257, 4608, 1000):
        roundtrip(tmp_path, samples, 16000, 16, blocks=blocks)
    roundtrip(tmp_path, samples, 16000, 16, blocks=1024, explicit_block=True)
    for rate, code in ((16000, 12), (16000, 13), (16000, 14), (16000, 0), (11025, 13), (11025, None),
                       (352800, 14), (655350, None)):
        roundtrip(tmp_path, samples[:, :3000], rate, 16, blocks=1024, explicit_rate=code)
    roundtrip(tmp_path, samples[:, :3000], 16000, 16, blocks=1024, streaminfo_size=True)


def test_variable_blocking_and_long_streams(tmp_path):
    samples = signal(2, 40000, 16, 8)
    roundtrip(tmp_path, samples, 16000, 16, blocks=[192, 4096, 1000, 16, 2304], variable=True,
              specs=lambda i: [dict(kind='fixed', order=i % 5)])
    # frame numbers past one coded byte (fixed blocking, > 127 frames): multi-byte coded numbers
    roundtrip(tmp_path, samples[:1], 16000, 16, blocks=16, specs=[dict(kind='fixed', order=1)])


def test_stream_wrappers(tmp_path):
    samples = signal(1, 3000, 16, 9)
    roundtrip(tmp_path, samples, 16000, 16, blocks=1024, extra_metadata=False)
    roundtrip(tmp_path, samples, 16000, 16, blocks=1024, id3=True)
    roundtrip(tmp_path, samples, 16000, 16, blocks=1024, md5=False)
    roundtrip(tmp_path, samples, 16000, 16, blocks=1024, total_known=False)        # streamed encoder
    path = tmp_path / 'tagged.flac'
    path.write_bytes(F.encode(samples, 16000, 16, blocks=1024, total_known=False) + b'TAG' + bytes(125))
    out, _ = decode(path)
    assert out.shape == (1, 3000)


def test_damage_is_detected(tmp_path):
    samples = signal(2, 6000, 16, 10)
    stream = bytearray(F.encode(samples, 16000, 16, blocks=1024, specs=[dict(kind='lpc', order=8)]))
    path = tmp_path / 'bad.flac'

    def failure(data):
        path.write_bytes(bytes(data))
        with pytest.raises((ValueError, RuntimeError)) as info:
            decode(path)
        return str(info.value)

    flipped = bytearray(stream)
    flipped[len(stream) // 2] ^= 0x10
    assert 'CRC' in failure(flipped) or 'sync' in failure(flipped) or 'residual' in failure(flipped)
    header = bytearray(stream)
    first = stream.index(b'\xff\xf8', 42)
    header[first + 2] ^= 0x10                       # block-size code of the first frame header
    assert 'CRC' in failure(header)
    assert 'truncated' in failure(stream[:-700]) or 'sync' in failure(stream[:-700]) or 'decoded' in failure(stream[:-700])
    wrong_md5 = bytearray(stream)
    wrong_md5[4 + 4 + 18] ^= 1
    assert 'MD5' in failure(wrong_md5)
    # a damaged sample that keeps both CRCs intact is exactly what the MD5 is for
    samples2 = samples.copy()
    samples2[0, 100] += 1
    forged = bytearray(F.encode(samples2, 16000, 16, blocks=1024, specs=[dict(kind='lpc', order=8)]))
    forged[8 + 18:8 + 34] = stream[8 + 18:8 + 34]
    assert 'MD5' in failure(forged)
    path.write_bytes(b'RIFF' + bytes(100))
    assert load.flac_info(path) is None
    with pytest.raises(ValueError, match='cannot open'):
        load.flac_info(tmp_path / 'missing.flac')
    small = torch.empty(2, 10)
    path.write_bytes(bytes(stream))
    code = _lib.lib.ppgs_flac_read_f32(str(path).encode(), ctypes.c_void_p(small.data_ptr()), 10, None, None, None)
    assert code != 0 and 'too small' in _lib.last_error()


def test_load_audio_reads_flac(tmp_path):
    """ppgs.load.audio on a 16 kHz FLAC file: (channels, samples) fp32, no resampling needed."""
    samples = signal(2, 16000, 16, 11)
    path = tmp_path / 'utterance.flac'
    path.write_bytes(F.encode(samples, 16000, 16, blocks=4096, stereo='mid_side',
                              specs=[dict(kind='lpc', order=8, partition_order=3)]))
    audio = load.audio(path)
    assert audio.shape == (2, 16000) and audio.dtype == torch.float32
    assert np.array_equal(audio.numpy(), (samples / 32768.0).astype(np.float32))
    assert load.wav_num_frames(path) == (16000, 16000)


def test_damaged_streams_never_crash_the_decoder(tmp_path):
    """Files are untrusted input: random byte damage, truncation and insertions into valid streams
    must end in a status code (almost always an error: CRC-8 / CRC-16 / MD5 / structure checks),
    never in a crash or a write outside the caller's buffer (guard rows stay untouched)."""
    import random
    random.seed(0)
    streams = [
        F.encode(signal(2, 3000, 16, 1), 16000, 16, blocks=1024, stereo='mid_side',
                 specs=[dict(kind='lpc', order=8, partition_order=3)]),
        F.encode(signal(1, 2500, 24, 2), 48000, 24, blocks=576,
                 specs=[dict(kind='fixed', order=3, method=1, partition_order=2, escape=(1,))]),
        F.encode(signal(3, 1500, 8, 3), 8000, 8, blocks=[192, 16, 1000], variable=True, specs=[dict(kind='verbatim')]),
        F.encode(signal(2, 900, 32, 4), 48000, 32, blocks=256, stereo='left_side',
                 specs=[dict(kind='lpc', order=32, precision=14)]),
    ]
    capacity = 4000
    out = torch.full((9, capacity), 7.0)            # 8 channels at most + one guard row
    path = tmp_path / 'fuzz.flac'
    rejected = 0
    for _ in range(1500):
        data = bytearray(random.choice(streams))
        mode = random.random()
        if mode < 0.6:
            for _ in range(random.randint(1, 4)):
                data[random.randrange(len(data))] = random.randrange(256)
        elif mode < 0.8:
            data = data[:random.randrange(1, len(data))]
        else:
            at = random.randrange(len(data))
            data[at:at] = bytes(random.randrange(256) for _ in range(random.randint(1, 8)))
        path.write_bytes(bytes(data))
        frames, rate, channels = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
        _lib.lib.ppgs_flac_info(str(path).encode(), ctypes.byref(frames), ctypes.byref(rate), ctypes.byref(channels), None)
        code = _lib.lib.ppgs_flac_read_f32(str(path).encode(), ctypes.c_void_p(out.data_ptr()), capacity,
                                           ctypes.byref(frames), ctypes.byref(rate), ctypes.byref(channels))
        rejected += code != 0
        assert bool((out[8] == 7.0).all())
    assert rejected >= 1400
