"""Ingest / egress on the GPU through the C ABI: PCM decode and the sinc resampler against
torchaudio (the reference's `ppgs.resample`, ppgs/core.py:599-608), and the native file
pipeline `ppgs_files_to_files` against the per-batch API on the same batches (bitwise) and
against the oracle (<= 1e-4)."""
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from oracle import ppg_oracle as O
from test_host_logic import write_wav

pytestmark = pytest.mark.gpu

# fp32 dot products of <= 16 015 taps in a different summation order than conv1d
RESAMPLE_TOL = 2e-6


@pytest.fixture(scope='module')
def ppgs_b200():
    import ppgs_b200
    return ppgs_b200


@pytest.fixture(scope='module')
def engine(ppgs_b200):
    engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(5, peaky=True))
    engine.precision = 'f16x2'
    return engine


@pytest.mark.parametrize('count', [1, 7, 8, 9, 4096, 160000 * 3 + 5])
def test_pcm16_to_f32_is_exact(engine, count):
    pcm = torch.randint(-32768, 32768, (count,), dtype=torch.int16)
    pcm[0] = -32768
    out = engine.pcm16_to_f32(pcm.cuda()).cpu()
    assert torch.equal(out, pcm.float() / 32768.0)


@pytest.mark.parametrize('orig,samples', [
    (44100, 44100), (22050, 33333), (8000, 8000), (48000, 100001), (11025, 5000),
    (32000, 433), (24000, 1)])
def test_resample_matches_torchaudio(engine, orig, samples):
    import torchaudio
    audio = O.synthetic_audio(3, samples, orig)   # (3, 1, samples)
    reference = torchaudio.transforms.Resample(orig, 16000)(audio)
    got = engine.resample(audio.cuda(), orig).cpu()
    assert got.shape == reference.shape
    assert (got - reference).abs().max() <= RESAMPLE_TOL


def test_public_resample_and_from_audio_at_other_rates(ppgs_b200, tmp_path):
    """`ppgs.resample` keeps the tensor where it was; `from_audio(audio, 22050)` equals the
    oracle on torchaudio-resampled audio."""
    import torchaudio
    sd = O.random_state_dict(5, peaky=True)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    audio = O.synthetic_audio(2, 30000, 9)
    reference = torchaudio.transforms.Resample(22050, 16000)(audio)
    on_cpu = ppgs_b200.resample(audio, 22050)
    assert on_cpu.device.type == 'cpu' and (on_cpu - reference).abs().max() <= RESAMPLE_TOL
    assert ppgs_b200.resample(audio.cuda(), 22050).is_cuda
    assert ppgs_b200.resample(audio, 16000) is audio
    out = ppgs_b200.from_audio(audio, 22050, checkpoint=checkpoint, gpu=0).cpu()
    expected = O.from_audio(sd, reference)
    assert out.shape == expected.shape
    assert (out - expected).abs().max() <= 1e-4
    # a 22.05 kHz file goes through the native decoder + the GPU resampler
    path = tmp_path / 'a.wav'
    wavfile.write(path, 22050, audio[0, 0].numpy())
    from_file = ppgs_b200.from_file(path, checkpoint=checkpoint, gpu=0).cpu()
    assert (from_file - expected[0]).abs().max() <= 1e-4
    with pytest.raises(ValueError, match='integer'):
        ppgs_b200.resample(audio.cuda(), 22050.5)


def make_files(tmp_path, lengths, channels=None):
    files = []
    for i, n in enumerate(lengths):
        files.append(str(tmp_path / f'{i}.wav'))
        write_wav(files[-1], n, seed=40 + i, channels=(channels or {}).get(i, 1))
    return files


def test_native_file_pipeline_equals_per_batch_api(ppgs_b200, tmp_path, monkeypatch):
    """Same batches through ppgs_files_to_files (reader / writer threads in the library,
    int16 over PCIe) and through the Python reader threads + from_dataloader: bitwise equal
    `.pt` files, cropped like save_masked; <= 1e-4 from the oracle run file by file... the
    oracle batch is the same zero-padded batch, so compare on the batch."""
    sd = O.random_state_dict(6, peaky=True)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    lengths = [16000, 40000, 90000, 8000, 16161, 32000, 64000, 433, 159999, 81000, 80160, 16000]
    files = make_files(tmp_path, lengths, channels={2: 2, 7: 2})
    native = [str(tmp_path / f'{i}-native.pt') for i in range(len(files))]
    python = [str(tmp_path / f'{i}-python.pt') for i in range(len(files))]
    calls = []
    original = ppgs_b200.Engine.files_to_files
    monkeypatch.setattr(ppgs_b200.Engine, 'files_to_files',
                        lambda self, *a, **k: calls.append(1) or original(self, *a, **k))
    ppgs_b200.from_files_to_files(files, native, checkpoint=checkpoint, num_workers=6, gpu=0,
                                  max_frames=1500)
    assert calls, 'the native pipeline was not used'
    monkeypatch.setenv('PPGS_B200_NATIVE_FILES', '0')
    ppgs_b200.from_files_to_files(files, python, checkpoint=checkpoint, num_workers=6, gpu=0,
                                  max_frames=1500)
    assert len(calls) == 1
    for a, b, n in zip(native, python, lengths):
        x, y = torch.load(a), torch.load(b)
        assert x.shape == (40, n // 160) and x.dtype == torch.float32
        assert torch.equal(x, y)
    # one batch against the oracle: the batch that holds file 0
    from ppgs_b200 import data
    loader = data.loader(files, num_workers=0, max_frames=1500)
    for audio, sample_lengths, names in loader:
        if files[0] in names:
            break
    expected = O.from_audio(sd, audio, lengths=sample_lengths)
    for row, name, n in zip(expected, names, sample_lengths.tolist()):
        got = torch.load(native[files.index(name)])
        assert (got - row[:, :n // 160]).abs().max() <= 1e-4


def test_native_file_pipeline_takes_flac(ppgs_b200, tmp_path, monkeypatch):
    """A corpus of 16 kHz FLAC files (16-bit mono, 16-bit mid-side stereo, 12-bit) next to WAVE
    files runs inside ppgs_files_to_files (FLAC decoded by the reader threads straight to the int16
    batch): bitwise the Python reader path, which decodes through ppgs_flac_read_f32."""
    import flac_writer
    sd = O.random_state_dict(6, peaky=True)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    rng = np.random.default_rng(0)
    files, lengths = [], [16000, 24000, 8000, 40000, 12345]
    for i, n in enumerate(lengths):
        pcm = (O.synthetic_audio(1, n, 60 + i)[0, 0].numpy() * 32768).round().clip(-32768, 32767).astype(np.int64)
        path = tmp_path / f'{i}.flac'
        if i == 1:      # stereo: channel 0 is what the pipeline keeps
            other = rng.integers(-2000, 2000, n)
            path.write_bytes(flac_writer.encode(np.stack([pcm, other]), 16000, 16, blocks=4096, stereo='mid_side',
                                                specs=[dict(kind='lpc', order=8, partition_order=3)]))
        elif i == 2:    # 12-bit samples
            path.write_bytes(flac_writer.encode((pcm >> 4)[None], 16000, 12, blocks=1152,
                                                specs=[dict(kind='fixed', order=2, partition_order=2)]))
        elif i == 3:    # a WAVE file in the same list
            path = tmp_path / f'{i}.wav'
            wavfile.write(path, 16000, pcm.astype(np.int16))
        else:
            path.write_bytes(flac_writer.encode(pcm[None], 16000, 16, blocks=4096,
                                                specs=[dict(kind='lpc', order=12, partition_order=4)]))
        files.append(str(path))
    native = [str(tmp_path / f'{i}-native.pt') for i in range(len(files))]
    python = [str(tmp_path / f'{i}-python.pt') for i in range(len(files))]
    calls = []
    original = ppgs_b200.Engine.files_to_files
    monkeypatch.setattr(ppgs_b200.Engine, 'files_to_files',
                        lambda self, *a, **k: calls.append(1) or original(self, *a, **k))
    ppgs_b200.from_files_to_files(files, native, checkpoint=checkpoint, num_workers=4, gpu=0, max_frames=600)
    assert calls, 'the native pipeline was not used'
    monkeypatch.setenv('PPGS_B200_NATIVE_FILES', '0')
    ppgs_b200.from_files_to_files(files, python, checkpoint=checkpoint, num_workers=4, gpu=0, max_frames=600)
    assert len(calls) == 1
    for a, b, n in zip(native, python, lengths):
        x, y = torch.load(a), torch.load(b)
        assert x.shape == (40, n // 160)
        assert torch.equal(x, y)
    # a damaged FLAC file fails the call with the decoder's message
    data = bytearray(open(files[0], 'rb').read())
    data[len(data) // 2] ^= 0x55
    open(files[0], 'wb').write(bytes(data))
    monkeypatch.delenv('PPGS_B200_NATIVE_FILES')
    with pytest.raises((RuntimeError, ValueError), match='CRC|sync|residual|MD5|subframe'):
        ppgs_b200.from_files_to_files(files, native, checkpoint=checkpoint, num_workers=4, gpu=0, max_frames=600)


def test_native_file_pipeline_many_batches(ppgs_b200, tmp_path):
    """More batches than staging slots, one reader / one writer thread and many."""
    sd = O.random_state_dict(6)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    rng = np.random.default_rng(0)
    lengths = [int(n) for n in rng.integers(3000, 30000, 60)]
    files = make_files(tmp_path, lengths)
    outputs = {}
    for workers in (2, 16):
        outputs[workers] = [str(tmp_path / f'{i}-{workers}.pt') for i in range(len(files))]
        ppgs_b200.from_files_to_files(files, outputs[workers], checkpoint=checkpoint,
                                      num_workers=workers, gpu=0, max_frames=400)
    engine = ppgs_b200.load.model(checkpoint, 'mel', 0)
    for i, n in enumerate(lengths):
        a, b = torch.load(outputs[2][i]), torch.load(outputs[16][i])
        assert a.shape == (40, n // 160) and torch.equal(a, b)
    engine.check()


def test_native_file_pipeline_reports_errors(ppgs_b200, tmp_path):
    sd = O.random_state_dict(6)
    engine = ppgs_b200.Engine(0).load_state_dict(sd)
    files = make_files(tmp_path, [16000, 16000, 16000])
    outputs = {file: file + '.pt' for file in files}
    samples = {file: 16000 for file in files}
    os.remove(files[1])
    with pytest.raises(ValueError, match='cannot open'):
        engine.files_to_files([[files[0]], [files[1]], [files[2]]], outputs, samples)
    # header / announced length mismatch
    write_wav(files[1], 8000)
    with pytest.raises(RuntimeError, match='announced length'):
        engine.files_to_files([[files[0], files[1]], [files[2]]], outputs, samples)
    # unwritable output
    write_wav(files[1], 16000)
    outputs[files[2]] = str(tmp_path / 'missing-dir' / 'x.pt')
    with pytest.raises(ValueError, match='cannot open for writing'):
        engine.files_to_files([[files[0], files[1]], [files[2]]], outputs, samples)
    # too short for the reflection padding
    write_wav(files[0], 300)
    with pytest.raises(ValueError, match='at least 433'):
        engine.files_to_files([[files[0]]], outputs, {files[0]: 300})
    # and the engine still works afterwards
    assert engine.files_to_files([[files[1]]], outputs, samples) == 100


def test_non_native_files_take_the_python_path(ppgs_b200, tmp_path):
    """A float32 or 22.05 kHz file in the list: batches come from the Python readers
    (native decode + GPU resampler per file), results still match the oracle."""
    import torchaudio
    sd = O.random_state_dict(6, peaky=True)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    a = O.synthetic_audio(1, 22050, 1)[0, 0]
    b = O.synthetic_audio(1, 16000, 2)[0, 0]
    fa, fb = str(tmp_path / 'a.wav'), str(tmp_path / 'b.wav')
    wavfile.write(fa, 22050, a.numpy())
    wavfile.write(fb, 16000, b.numpy())
    outs = [str(tmp_path / 'a.pt'), str(tmp_path / 'b.pt')]
    ppgs_b200.from_files_to_files([fa, fb], outs, checkpoint=checkpoint, num_workers=2, gpu=0)
    resampled = torchaudio.transforms.Resample(22050, 16000)(a[None])
    assert resampled.shape[-1] == 16000
    expected = O.from_audio(sd, torch.stack([resampled, b[None]]))
    for out, row in zip(outs, expected):
        assert (torch.load(out) - row).abs().max() <= 1e-4


def test_flac_files_through_the_file_api(ppgs_b200, tmp_path):
    """FLAC corpora (f2): a 16 kHz mono and a 22.05 kHz stereo FLAC file through
    from_files_to_files (native verifying decoder -> GPU resampler -> engine) against the
    oracle on the same samples; channel 0 is what ppgs/data/collate.py:27 keeps."""
    import torchaudio
    import flac_writer
    sd = O.random_state_dict(6, peaky=True)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    a = (O.synthetic_audio(1, 22050, 3)[0, 0] * 32768).round().clamp(-32768, 32767)
    b = (O.synthetic_audio(1, 16000, 4)[0, 0] * 32768).round().clamp(-32768, 32767)
    fa, fb = tmp_path / 'a.flac', tmp_path / 'b.flac'
    stereo = np.stack([a.numpy(), -a.numpy().clip(-32767, 32767)]).astype(np.int64)
    fa.write_bytes(flac_writer.encode(stereo, 22050, 16, blocks=4096, stereo='mid_side',
                                      specs=[dict(kind='lpc', order=8, partition_order=3)]))
    fb.write_bytes(flac_writer.encode(b.numpy()[None].astype(np.int64), 16000, 16, blocks=1152,
                                      specs=[dict(kind='fixed', order=2, partition_order=2)]))
    outs = [str(tmp_path / 'a.pt'), str(tmp_path / 'b.pt')]
    ppgs_b200.from_files_to_files([str(fa), str(fb)], outs, checkpoint=checkpoint, num_workers=2, gpu=0)
    resampled = torchaudio.transforms.Resample(22050, 16000)(a[None] / 32768)
    expected = O.from_audio(sd, torch.stack([resampled, b[None] / 32768]))
    for out, row in zip(outs, expected):
        assert (torch.load(out) - row).abs().max() <= 1e-4


def test_files_sharded_over_a_gpu_list(ppgs_b200, tmp_path):
    """`gpu=[...]`: one pipeline thread per listed device over batches i::n of the same batch
    list (here the same device twice = two engines sharing a blob copy): results equal the
    single-pipeline run bit for bit."""
    sd = O.random_state_dict(8)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    rng = np.random.default_rng(1)
    lengths = [int(n) for n in rng.integers(3000, 40000, 30)]
    files = make_files(tmp_path, lengths)
    single = [str(tmp_path / f'{i}-single.pt') for i in range(len(files))]
    sharded = [str(tmp_path / f'{i}-sharded.pt') for i in range(len(files))]
    ppgs_b200.from_files_to_files(files, single, checkpoint=checkpoint, num_workers=4, gpu=0,
                                  max_frames=500)
    devices = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    ppgs_b200.from_files_to_files(files, sharded, checkpoint=checkpoint, num_workers=4, gpu=devices,
                                  max_frames=500)
    for a, b, n in zip(single, sharded, lengths):
        x, y = torch.load(a), torch.load(b)
        assert x.shape == (40, n // 160) and torch.equal(x, y)


def test_native_file_pipeline_uniform_batches_replay_graphs(ppgs_b200, tmp_path, monkeypatch):
    """Equal-length files: every batch has the same plan, so the pipeline's forwards are
    replayed from the engine's CUDA-graph cache — the files still equal the per-batch API's."""
    sd = O.random_state_dict(10, peaky=True)
    checkpoint = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, checkpoint)
    files = make_files(tmp_path, [16000] * 50)
    native = [str(tmp_path / f'{i}-native.pt') for i in range(len(files))]
    python = [str(tmp_path / f'{i}-python.pt') for i in range(len(files))]
    engine = ppgs_b200.load.model(checkpoint, 'mel', 0)
    before = engine.graph_replays
    ppgs_b200.from_files_to_files(files, native, checkpoint=checkpoint, num_workers=4, gpu=0,
                                  max_frames=500)      # 5 files per batch, 10 equal batches
    assert engine.graph_replays - before == 9
    monkeypatch.setenv('PPGS_B200_NATIVE_FILES', '0')
    engine.set_graphs(False)
    try:
        ppgs_b200.from_files_to_files(files, python, checkpoint=checkpoint, num_workers=4, gpu=0,
                                      max_frames=500)
    finally:
        engine.set_graphs(True)
    for a, b in zip(native, python):
        assert torch.equal(torch.load(a), torch.load(b))


def test_preprocess_from_files_to_files(ppgs_b200, tmp_path):
    """ppgs.preprocess.from_files_to_files (ppgs/preprocess/core.py:63-97): fp16 feature files,
    cropped per file, equal to the batch front-end's rows and within 1 fp16 ulp of the oracle."""
    from test_mel_emulation import ulp_distance
    from ppgs_b200 import data, preprocess
    lengths = [16000, 40000, 24321, 8000, 16161, 32000]
    files = make_files(tmp_path, lengths)
    outputs = [str(tmp_path / f'{i}-{{}}.pt') for i in range(len(files))]
    preprocess.from_files_to_files(files, outputs, ['mel'], num_workers=4, gpu=0)
    loader = data.loader(files, num_workers=0, max_frames=ppgs_b200.config.MAX_PREPROCESS_FRAMES)
    seen = 0
    for audio, sample_lengths, names in loader:
        expected = ppgs_b200.preprocess.mel.from_audios(audio.cuda(), sample_lengths, gpu=0).cpu()
        oracle = O.mel_from_audios(audio)
        for row, reference, name, n in zip(expected, oracle, names, sample_lengths.tolist()):
            got = torch.load(str(tmp_path / f'{files.index(name)}-mel.pt'))
            assert got.dtype == torch.float16 and got.shape == (80, n // 160)
            assert torch.equal(got, row[:, :n // 160])
            assert ulp_distance(got.numpy(), reference[:, :n // 160].numpy()).max() <= 1
            seen += 1
    assert seen == len(files)
    with pytest.raises(ValueError, match='not supported'):
        preprocess.from_files_to_files(files, outputs, ['bottleneck'], gpu=0)


def test_from_feature_files_to_files_and_container(ppgs_b200, tmp_path):
    """cached `<stem>-mel.pt` features (what `python -m ppgs.preprocess` writes and
    ppgs/data/dataset.py:98-101 loads) -> posteriorgrams, read by the native `.pt` reader into
    padded batches; one `.pt` per input, or one container file per batch."""
    import torch
    from oracle import ppg_oracle as O
    sd = O.random_state_dict(5)
    ckpt = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, ckpt)
    frames = [400, 333, 1000, 57]
    feature_files, output_files, cached = [], [], []
    for i, n in enumerate(frames):
        feats = O.mel_from_audios(O.synthetic_audio(1, n * 160, 20 + i))[0]
        file = tmp_path / f'utt{i}-mel.pt'
        torch.save(feats, file)                                  # torch's own writer
        feature_files.append(file)
        output_files.append(tmp_path / f'utt{i}-ppg.pt')
        cached.append(feats)
    ppgs_b200.from_feature_files_to_files(feature_files, output_files, checkpoint=ckpt, gpu=0, num_workers=4,
                                          max_frames=1500)
    # the reference's semantics depend on the batch (chunking is decided by the padded length,
    # SURVEY §3.2): compare batch against batch, composed by the same sampler
    for batch in ppgs_b200.data.frame_budget_batches(frames, 1500):
        longest = max(frames[i] for i in batch)
        padded = torch.zeros(len(batch), 80, longest, dtype=torch.float16)
        for row, i in enumerate(batch):
            padded[row, :, :frames[i]] = cached[i]
        ref = O.from_features(sd, padded, torch.tensor([frames[i] for i in batch]))
        for row, i in enumerate(batch):
            out = torch.load(output_files[i])
            assert out.shape == (40, frames[i])
            assert (out - ref[row, :, :frames[i]]).abs().max() <= 1e-4
    ppgs_b200.from_feature_files_to_files(feature_files, output_files, checkpoint=ckpt, gpu=0, max_frames=1500,
                                          container=tmp_path / 'shard')
    seen = {}
    for file in sorted(tmp_path.glob('shard.*.pt')):
        shard = torch.load(file)
        for name, n, ppg in zip(shard['files'], shard['lengths'].tolist(), shard['ppgs']):
            seen[name] = ppg[:, :n]
    assert len(seen) == len(frames)
    for file, n in zip(output_files, frames):
        assert torch.equal(seen[str(file)], torch.load(file))
