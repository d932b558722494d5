"""Host-side logic that needs no GPU: the batch scheduler against the oracle's
restatement of the reference sampler / collate, the WAV reader, the CLI flags,
and the multi-process sharding (world_size 2 over gloo)."""
import os
import socket
import subprocess
import sys
import wave

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import ppg_oracle as O


def write_wav(path, samples, rate=16000, seed=0, channels=1):
    rng = np.random.default_rng(seed)
    data = (rng.uniform(-0.5, 0.5, size=(samples, channels)) * 32767).astype(np.int16)
    with wave.open(str(path), 'wb') as f:
        f.setnchannels(channels)
        f.setsampwidth(2)
        f.setframerate(rate)
        f.writeframes(data.tobytes())
    return data


@pytest.mark.parametrize('max_frames', [1000, 4000, 64000])
def test_frame_budget_batches_equal_reference_sampler(max_frames):
    from ppgs_b200 import data
    rng = np.random.default_rng(3)
    lengths = rng.integers(10, 1000, size=257).tolist()
    mine = data.frame_budget_batches(lengths, max_frames)
    assert mine == O.sampler_batches(lengths, max_frames)
    assert sorted(i for b in mine for i in b) == list(range(257))
    for batch in mine:
        assert len(batch) == 1 or len(batch) * max(lengths[i] for i in batch) <= max_frames


def test_collate_zero_pads_like_reference():
    from ppgs_b200 import data
    audios = [torch.randn(1, n) for n in (1600, 4801, 320)]
    padded, lengths = data.collate(audios)
    ref_padded, ref_lengths = O.collate(audios)
    assert torch.equal(padded, ref_padded) and torch.equal(lengths, ref_lengths)


def test_wav_reader_and_loader(tmp_path):
    from ppgs_b200 import data, load
    files, lengths = [], [16000, 4000, 32000, 8000, 8000]
    for i, n in enumerate(lengths):
        files.append(tmp_path / f'{i}.wav')
        raw = write_wav(files[-1], n, seed=i)
        audio = load.audio(files[-1])
        assert audio.shape == (1, n) and audio.dtype == torch.float32
        assert np.array_equal(audio[0].numpy(), raw[:, 0].astype(np.float32) / 32768.0)
        assert load.wav_num_frames(files[-1]) == (n, 16000)
    loader = data.loader(files, num_workers=2, max_frames=250)
    seen = []
    for audio, sample_lengths, names in loader:
        assert audio.shape == (len(names), 1, int(sample_lengths.max()))
        for row, n, name in zip(audio, sample_lengths, names):
            assert torch.equal(row[0, :n], load.audio(name)[0])
            assert not row[0, n:].any()
        seen.extend(names)
    assert sorted(seen) == sorted(files)
    with pytest.warns(UserWarning, match='exceeds max_frames'):
        assert len(data.Metadata(files, max_frames=100)) == 4
    with pytest.raises(ValueError):
        data.loader(files, features=['phonemes'])


def test_shard_by_frames_balances():
    from ppgs_b200 import parallel
    lengths = [1000] * 7 + [250] * 9 + [40] * 30
    parts = parallel.shard_by_frames(lengths, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 1000
    assert parallel.shard(list(range(10)), rank=1, world=4) == [1, 5, 9]


def test_cli_flags_match_reference():
    from ppgs_b200.__main__ import parse_args
    args = parse_args(['--audio_files', 'a.wav', 'b.wav', '--output_files', 'a.pt', 'b.pt',
                       '--num-workers', '4', '--gpu', '1', '--max-frames', '64000',
                       '--legacy-mode', '--checkpoint', 'x.pt'])
    assert [str(p) for p in args.audio_files] == ['a.wav', 'b.wav']
    assert args.num_workers == 4 and args.gpu == 1 and args.legacy_mode
    assert args.representation == 'mel' and args.max_frames == 64000
    sharded = parse_args(['--audio_files', 'a.wav', '--output_files', 'a.pt', '--gpu', '0', '1', '3'])
    assert sharded.gpu == [0, 1, 3]


def test_unknown_representation_raises_value_error():
    from ppgs_b200 import load, preprocess
    with pytest.raises(ValueError):
        preprocess.get('bottleneck')
    with pytest.raises(ValueError):
        load.model_kwargs('encodec')
    assert load.model_kwargs('w2v2fb') == {'hidden_channels': 512, 'input_channels': 768}


def test_engine_mel_basis_equals_oracle():
    from ppgs_b200 import engine
    assert np.array_equal(engine.mel_basis(), O.mel_basis())


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_two_rank_gloo_sharding(tmp_path):
    """N>1 host path on CPU: two ranks over gloo build the same batch list, take
    disjoint shards that cover every file, and receive rank 0's blob bytes."""
    files = []
    for i, n in enumerate([16000, 4000, 32000, 8000, 8000, 12000, 24000]):
        files.append(str(tmp_path / f'{i}.wav'))
        write_wav(files[-1], n, seed=i)
    out = tmp_path / 'out'
    out.mkdir()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()),
           os.path.join(ROOT, 'tests', '_dist_worker.py'), str(out)] + files
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS='1')
    subprocess.run(cmd, check=True, timeout=240, env=env, cwd=ROOT)
    shards = [open(out / f'rank{r}.txt').read().split() for r in range(2)]
    assert sorted(shards[0] + shards[1]) == sorted(files)
    assert not set(shards[0]) & set(shards[1])
    blobs = [np.load(out / f'blob{r}.npy') for r in range(2)]
    assert np.array_equal(blobs[0], blobs[1]) and blobs[0].sum() > 0


def test_configure_applies_yapecs_style_overrides(tmp_path):
    import ppgs_b200
    file = tmp_path / 'causal_transformer.py'
    file.write_text("MODULE = 'ppgs'\nCONFIG = 'causal_transformer'\nIS_CAUSAL = True\nlower = 1\n")
    try:
        applied = ppgs_b200.configure(file)
        assert applied == {'MODULE': 'ppgs', 'CONFIG': 'causal_transformer', 'IS_CAUSAL': True}
        assert ppgs_b200.IS_CAUSAL is True and ppgs_b200.config.IS_CAUSAL is True
        assert ppgs_b200.load.cache_key.__defaults__ is not None
    finally:
        ppgs_b200.configure({'IS_CAUSAL': False})
    assert ppgs_b200.config.IS_CAUSAL is False


def test_bounded_max_frames_keeps_finite_budgets():
    from ppgs_b200 import data
    assert data.bounded_max_frames(64000) == 64000
    assert data.bounded_max_frames(1) == 1
    if not torch.cuda.is_available():      # nothing to derive a budget from: unchanged
        assert data.bounded_max_frames(float('inf')) == float('inf')


def test_loader_reuses_a_header_probe(tmp_path):
    from ppgs_b200 import data
    files = []
    for i, n in enumerate([16000, 8000, 12000, 4000]):
        files.append(tmp_path / f'{i}.wav')
        write_wav(files[-1], n, seed=i)
    first = data.loader(files, num_workers=0, max_frames=120, shard=(0, 2))
    second = data.loader(files, num_workers=0, max_frames=120, shard=(1, 2), dataset=first.dataset)
    assert second.dataset is first.dataset
    together = sorted(i for batch in first.batches + second.batches for i in batch)
    assert together == [0, 1, 2, 3]
    everything = data.loader(files, num_workers=0, max_frames=120)
    assert everything.batches[0::2] == first.batches and everything.batches[1::2] == second.batches


def test_configuration_is_read_at_call_time(tmp_path):
    """`configure()` / `--config` after import must reach argument defaults (they
    are `config.live(...)` placeholders resolved inside the functions), the CLI's own defaults
    and the engine cache key — not values frozen when the modules were imported."""
    import inspect
    import ppgs_b200
    from ppgs_b200 import config, core, data, engine, __main__ as cli
    saved = {k: getattr(config, k) for k in ('REPRESENTATION', 'IS_CAUSAL', 'HIDDEN_CHANNELS', 'MAX_INFERENCE_FRAMES')}
    try:
        default = inspect.signature(core.from_audio).parameters['representation'].default
        assert isinstance(default, config.Live) and config.resolve(default) == 'mel'
        cfg = tmp_path / 'causal.py'
        cfg.write_text("IS_CAUSAL = True\nREPRESENTATION = 'w2v2fb'\nMAX_INFERENCE_FRAMES = 12345\n")
        args = cli.parse_args(['--audio_files', 'a.wav', '--output_files', 'a.pt', '--config', str(cfg)])
        assert args.representation == 'w2v2fb' and args.max_frames == 12345   # defaults read AFTER --config
        assert config.resolve(default) == 'w2v2fb' and ppgs_b200.IS_CAUSAL is True
        for fn, name in ((core.from_features, 'representation'), (core.from_files_to_files, 'max_frames'),
                         (data.loader, 'max_frames'), (engine.Engine.__init__, 'is_causal'),
                         (engine.Engine.__init__, 'hidden_channels')):
            assert isinstance(inspect.signature(fn).parameters[name].default, config.Live), (fn, name)
        assert config.resolve(inspect.signature(engine.Engine.__init__).parameters['is_causal'].default) is True
    finally:
        config.configure(saved)
