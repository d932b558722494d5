"""Parity tests proper: the CUDA library, called through the C ABI
(ppgs_b200._lib / Engine), against the oracle and against the committed golden
outputs of the reference's own modules.

Tolerances
* mel features: <= 1 fp16 ulp from the reference mel, mismatch rate <= 1e-3 (same
  bound against the CPU replay of the kernel arithmetic);
* posteriorgrams: <= 1e-4 max-abs (BASELINE.json north_star) against the
  reference modules evaluated in fp32 with autocast off (oracle mode O3).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ppg_oracle as O
from test_mel_emulation import emulate, ulp_distance
from test_oracle_golden import MEL, PPG, mel_case_audio, ppg_case_inputs

pytestmark = pytest.mark.gpu

PPG_TOL = 1e-4
PRECISIONS = ['fp32', 'f16x2']


@pytest.fixture(scope='module')
def ppgs_b200():
    import ppgs_b200
    return ppgs_b200


def make_engine(ppgs_b200, sd, precision, **kwargs):
    engine = ppgs_b200.Engine(0, **kwargs)
    engine.load_state_dict(sd)
    try:
        engine.precision = precision
    except RuntimeError as error:
        if 'not available' in str(error):
            pytest.skip(f'precision {precision} not built')
        raise
    return engine


@pytest.fixture(scope='module')
def frontend(ppgs_b200):
    return make_engine(ppgs_b200, O.random_state_dict(0), 'fp32')


# --------------------------------------------------------------------------- mel
@pytest.mark.parametrize('name', MEL)
def test_mel_vs_reference_golden(frontend, mel_emul_lib, name):
    g = golden(name)
    audio = mel_case_audio(g)
    mel = frontend.mel(audio.cuda()).cpu().numpy()
    assert mel.shape == g['mel'].shape and mel.dtype == np.float16
    dist = ulp_distance(mel, g['mel'])
    assert dist.max() <= 1
    assert (dist > 0).mean() <= 1e-3
    # the kernel and its CPU replay execute the same fp32 operation sequence; only
    # logf differs (CUDA's is faithfully, glibc's correctly rounded)
    replay = ulp_distance(mel, emulate(mel_emul_lib, audio))
    assert replay.max() <= 1 and (replay > 0).mean() <= 1e-3


@pytest.mark.parametrize('batch,samples', [(1, 433), (3, 160 * 33 + 7), (2, 160 * 64), (5, 48000)])
def test_mel_edges_vs_oracle(frontend, batch, samples):
    """Short inputs (everything is reflect padding), ragged tile tails, odd frame
    counts (scalar store path) and unaligned rows (non-TMA load path)."""
    audio = O.synthetic_audio(batch, samples, seed=batch)
    ref = O.mel_from_audios(audio).numpy()
    mel = frontend.mel(audio.cuda()).cpu().numpy()
    assert ulp_distance(mel, ref).max() <= 1
    # a view whose rows start at a 4-byte (not 16-byte) aligned address
    wide = torch.zeros(batch, samples + 3, device='cuda')
    wide[:, 1:samples + 1] = audio[:, 0].cuda()
    mel2 = frontend.mel(wide[:, 1:samples + 1]).cpu().numpy()
    assert np.array_equal(mel2.view(np.uint16), mel.view(np.uint16))


def test_mel_rejects_too_short(frontend):
    with pytest.raises(ValueError):
        frontend.mel(torch.zeros(1, 1, 432, device='cuda'))


# ------------------------------------------------------------------- transformer
@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('name', PPG)
def test_transformer_vs_reference_golden(ppgs_b200, name, precision):
    g = golden(name)
    sd, audio, lengths = ppg_case_inputs(g)
    feats = O.mel_from_audios(audio)
    engine = make_engine(ppgs_b200, sd, precision, is_causal=bool(g['causal']))
    ppg = engine.transformer(feats.cuda(), lengths).cpu().numpy()
    assert ppg.shape == g['ppg'].shape
    err = np.abs(ppg - g['ppg']).max()
    assert err <= PPG_TOL, f'{name} {precision}: max-abs {err:.3e}'
    logits = engine.transformer(feats.cuda(), lengths, softmax=False).cpu().numpy()
    assert np.abs(logits - g['logits']).max() <= 2e-3
    engine.check()      # no kernel reported a pipeline time-out


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('frames,lengths', [
    (160, [160]), (400, [400, 399, 3]), (500, [500]), (501, [501, 500]),
    (900, [900, 450, 451, 50, 49]), (1000, [1000] * 3), (1234, [1234, 2]),
    (3, [3]), (5, [5, 4, 1]), (4801, [4801])])
def test_from_audio_end_to_end_vs_oracle(ppgs_b200, frames, lengths, precision):
    """T3: audio in, posteriors out, batched-vs-batched, equal and ragged lengths
    around the 500/400/50 chunk boundaries."""
    sd = O.random_state_dict(3, peaky=True)
    audio = O.synthetic_audio(len(lengths), frames * 160, seed=frames)
    sample_lengths = torch.tensor(lengths) * 160
    ref = O.from_audio(sd, audio, lengths=sample_lengths).numpy()
    engine = make_engine(ppgs_b200, sd, precision)
    out = engine.from_audio(audio.cuda(), lengths=sample_lengths).cpu().numpy()
    assert out.shape == ref.shape
    for row, n in enumerate(lengths):      # valid frames (save_masked crops the rest)
        err = np.abs(out[row, :, :n] - ref[row, :, :n]).max()
        assert err <= PPG_TOL, f'row {row}: {err:.3e}'
    pinned = audio.pin_memory()
    host = engine.from_audio_host(pinned, lengths=sample_lengths).numpy()
    assert np.array_equal(host, out)
    # pipelined submit / wait: three requests through the two slots
    outs = [engine.from_audio_host(pinned, lengths=sample_lengths, wait=False) for _ in range(3)]
    engine.wait()
    for other in outs:
        assert np.array_equal(other.numpy(), out)


@pytest.mark.parametrize('precision', PRECISIONS)
def test_legacy_mode_and_logits(ppgs_b200, precision):
    sd = O.random_state_dict(5)
    feats = O.mel_from_audios(O.synthetic_audio(2, 700 * 160, 9))
    lengths = torch.tensor([700, 512])
    engine = make_engine(ppgs_b200, sd, precision)
    ref = O.from_features(sd, feats, lengths, legacy_mode=True).numpy()
    out = engine.transformer(feats.cuda(), lengths, legacy_mode=True).cpu().numpy()
    assert np.abs(out[0] - ref[0]).max() <= PPG_TOL
    assert np.abs(out[1, :, :512] - ref[1, :, :512]).max() <= PPG_TOL
    chunked = engine.transformer(feats.cuda(), lengths).cpu().numpy()
    assert np.abs(chunked - out).max() > 1e-3      # chunking is semantics


def test_single_pass_f16_mode_is_the_autocast_numerics_class(ppgs_b200):
    """PPGS_PRECISION_F16 (one MMA pass, fp16 operands) is NOT the parity mode: it sits in
    the error class of the reference's own CUDA autocast (~1e-3), an order of
    magnitude above the 1e-4 that the split-fp16 mode meets."""
    sd = O.random_state_dict(1, peaky=True)
    audio = O.synthetic_audio(2, 400 * 160, 2)
    ref = O.from_audio(sd, audio).numpy()
    err = {}
    for precision in ('f16', 'f16x2'):
        engine = make_engine(ppgs_b200, sd, precision)
        out = torch.full((2, 40, 400), float('nan'), device='cuda')   # never a recycled result
        out.copy_(engine.from_audio(audio.cuda()))
        engine.check()
        err[precision] = np.abs(out.cpu().numpy() - ref).max()
    assert err['f16x2'] <= PPG_TOL < err['f16'] <= 2e-2


@pytest.mark.parametrize('switch', ['PPGS_B200_FUSED_FFN=0', 'PPGS_B200_FUSED_FFN=2', 'PPGS_B200_PAIR=0', 'PPGS_B200_ATTENTION=0',
                                    'PPGS_B200_ATTN_DUAL=0', 'PPGS_B200_ATTN_QK_PLANES=2', 'PPGS_B200_PROJ_LN=0',
                                    'PPGS_B200_MEL_ROWS=0', 'PPGS_B200_SERPENTINE=0', 'PPGS_B200_L2_HINTS=1'])
def test_alternative_kernel_paths(ppgs_b200, monkeypatch, switch):
    """The optional kernels stay parity-checked: linear1 / linear2 as two GEMMs instead of the
    fused FFN kernel, the single-CTA (cta_group::1) GEMMs, the CUDA-core attention kernel, the
    one-tile-per-CTA attention kernel, split-plane Q / K and the (B, 80, T) mel tensor + fold pass
    instead of the mel kernel writing the operand rows (read at engine creation)."""
    name, value = switch.split('=')
    monkeypatch.setenv(name, value)
    sd = O.random_state_dict(6, peaky=True)
    engine = make_engine(ppgs_b200, sd, 'f16x2')
    for frames, lengths in ((400, [400, 250, 31]), (1000, [1000, 640])):
        audio = O.synthetic_audio(len(lengths), frames * 160, seed=frames + 1)
        sample_lengths = torch.tensor(lengths) * 160
        ref = O.from_audio(sd, audio, lengths=sample_lengths).numpy()
        out = engine.from_audio(audio.cuda(), lengths=sample_lengths).cpu().numpy()
        for row, n in enumerate(lengths):
            assert np.abs(out[row, :, :n] - ref[row, :, :n]).max() <= PPG_TOL
    engine.check()


@pytest.mark.parametrize('frames,lengths,legacy', [
    (400, [400, 250, 31], False), (500, [500], False), (501, [501, 77], False), (1000, [1000, 640], False),
    (1237, [1237, 1200, 801, 399, 5], False), (2003, [2003], False), (33, [33, 1], False),
    (4000, [4000, 123], True)])
def test_mel_operand_rows_equal_the_fold_pass(ppgs_b200, frames, lengths, legacy):
    """from_audio's mel kernel writes the input convolution's operand rows itself (chunk
    overlap, replicate padding of chunk 0, zero rows behind every chunk tensor); the result must
    be BITWISE that of ppgs_mel_forward -> (B, 80, T) -> fold pass -> Transformer, for chunked,
    unchunked, ragged, odd-sized and legacy (one long sequence) batches."""
    sd = O.random_state_dict(3, peaky=True)
    engine = make_engine(ppgs_b200, sd, 'f16x2')
    audio = O.synthetic_audio(len(lengths), frames * 160 + 37, seed=frames).cuda()
    sample_lengths = torch.tensor(lengths) * 160 + 37
    direct = engine.from_audio(audio, lengths=sample_lengths, legacy_mode=legacy).clone()
    engine.check()
    feats = engine.mel(audio)
    folded = engine.transformer(feats, torch.tensor(lengths), legacy_mode=legacy)
    assert torch.equal(direct, folded)
    engine.set_profiling(True)
    engine.from_audio(audio, lengths=sample_lengths, legacy_mode=legacy)
    names = set(engine.kernel_stats())
    engine.set_profiling(False)
    assert 'mel_stft_fbank_rows' in names and 'fold_chunks' not in names


def test_broadcast_engine_follows_live_configuration(ppgs_b200):
    """after configure({'IS_CAUSAL': True}) the torchrun load path must build causal
    engines (it used to capture IS_CAUSAL at import and cache a non-causal engine under the
    causal key)."""
    from ppgs_b200 import config, load, parallel
    sd = O.random_state_dict(0, peaky=True)
    feats = O.mel_from_audios(O.synthetic_audio(1, 160 * 160, 3))
    lengths = torch.tensor([160])
    try:
        config.configure({'IS_CAUSAL': True})
        engine = parallel.broadcast_engine(sd, 'mel', 0)
        assert engine.cfg.is_causal == 1
        assert load.cache_key('mel', None, 0)[3] is True
        ref = O.from_features(sd, feats, lengths, is_causal=True)
        assert (engine.transformer(feats.cuda(), lengths).cpu() - ref).abs().max() <= PPG_TOL
    finally:
        config.configure({'IS_CAUSAL': False})
    assert parallel.broadcast_engine(sd, 'mel', 0).cfg.is_causal == 0


def test_error_conventions(ppgs_b200):
    engine = make_engine(ppgs_b200, O.random_state_dict(0), 'fp32')
    feats = torch.zeros(2, 80, 100, dtype=torch.float16, device='cuda')
    with pytest.raises(ValueError):      # max(lengths) != T: reference fails at `* mask`
        engine.transformer(feats, torch.tensor([90, 80]))
    with pytest.raises(ValueError):      # F4: one length per row
        engine.transformer(feats, torch.tensor([100]))
    with pytest.raises(ValueError, match='size is too large'):
        engine.transformer(torch.zeros(1, 80, 5001, dtype=torch.float16, device='cuda'),
                           torch.tensor([5001]), legacy_mode=True)
    bad = dict(O.random_state_dict(0))
    bad.pop('output_layer.bias')
    with pytest.raises(RuntimeError, match='missing key'):
        ppgs_b200.Engine(0).load_state_dict(bad)
    bad = dict(O.random_state_dict(0))
    bad['input_layer.weight'] = torch.zeros(256, 81, 5)
    with pytest.raises(ValueError, match='size mismatch'):
        ppgs_b200.Engine(0).load_state_dict(bad)


# ---------------------------------------------------------------- public API
def test_public_api_and_files(ppgs_b200, tmp_path):
    """from_audio / from_features / from_file / from_files_to_files with
    `checkpoint=` (T4: .pt outputs cropped like save_masked)."""
    from test_host_logic import write_wav
    sd = O.random_state_dict(2)
    ckpt = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, ckpt)
    audio = O.synthetic_audio(2, 64000, 1)
    ref = O.from_audio(sd, audio).numpy()
    out = ppgs_b200.from_audio(audio, 16000, checkpoint=ckpt, gpu=0)
    assert out.is_cuda and out.dtype == torch.float32 and out.shape == (2, 40, 400)
    assert np.abs(out.cpu().numpy() - ref).max() <= PPG_TOL
    feats = ppgs_b200.preprocess.from_audio(audio, 'mel', 16000, gpu=0)
    assert feats.dtype == torch.float16 and feats.shape == (2, 80, 400)
    out2 = ppgs_b200.from_features(feats, torch.tensor([400, 400]), checkpoint=ckpt, gpu=0)
    assert torch.equal(out, out2)
    with pytest.raises(ValueError):
        ppgs_b200.from_audio(audio, 16000, representation='dac', checkpoint=ckpt, gpu=0)

    files, outputs, lengths = [], [], [16000, 40000, 90000, 8000, 16161]
    for i, n in enumerate(lengths):
        files.append(tmp_path / f'{i}.wav')
        outputs.append(tmp_path / f'{i}-ppg.pt')
        write_wav(files[-1], n, seed=10 + i)
    ppgs_b200.from_files_to_files(files, outputs, checkpoint=ckpt, gpu=0)
    single = [torch.load(o) for o in outputs]
    for ppg, file, n in zip(single, files, lengths):
        assert ppg.shape == (40, n // 160) and ppg.dtype == torch.float32
        expect = O.from_audio(sd, ppgs_b200.load.audio(file)[None])[0].numpy()
        assert np.abs(ppg.numpy() - expect).max() <= PPG_TOL
    # batched path: same batches as the reference sampler -> batched-vs-batched parity
    batched_out = [tmp_path / f'{i}-b.pt' for i in range(len(files))]
    ppgs_b200.from_files_to_files(files, batched_out, checkpoint=ckpt, num_workers=4, gpu=0,
                                  max_frames=1200)
    for batch in ppgs_b200.data.frame_budget_batches([n // 160 for n in lengths], 1200):
        audios = [ppgs_b200.load.audio(files[i]) for i in batch]
        padded, sample_lengths = O.collate(audios)
        expect = O.from_audio(sd, padded, lengths=sample_lengths).numpy()
        for row, i in enumerate(batch):
            got = torch.load(batched_out[i]).numpy()
            assert got.shape == (40, lengths[i] // 160)
            assert np.abs(got - expect[row, :, :got.shape[1]]).max() <= PPG_TOL


# ------------------------------------------------------- full-size properties
@pytest.mark.parametrize('precision', PRECISIONS)
def test_config2_full_size_properties(ppgs_b200, precision):
    """BASELINE config 2 (64 x 10 s): size-independent properties + an oracle
    check of two rows (rows are independent when all lengths are equal)."""
    sd = O.random_state_dict(0, peaky=True)
    engine = make_engine(ppgs_b200, sd, precision)
    audio = O.synthetic_audio(64, 160000, 0)
    out = engine.from_audio(audio.cuda())
    assert out.shape == (64, 40, 1000)
    assert torch.isfinite(out).all() and (out >= 0).all()
    assert (out.sum(1) - 1).abs().max() <= 1e-5
    # batch-composition independence for equal-length rows
    sub = engine.from_audio(audio[5:7].cuda())
    assert (sub - out[5:7]).abs().max() <= 1e-6
    ref = O.from_audio(sd, audio[5:7])
    assert (out[5:7].cpu() - ref).abs().max() <= PPG_TOL
    engine.check()
    # launch accounting used by bench.py
    before = engine.launches
    engine.from_audio(audio.cuda())
    assert engine.launches > before


@pytest.mark.parametrize('precision', ['fp32', 'f16x2'])
def test_weight_blob_roundtrip(ppgs_b200, precision):
    """The multi-GPU load path on one device: a second engine adopts the first
    engine's packed blob (what the NCCL broadcast delivers) and agrees bitwise at the
    same arithmetic mode."""
    sd = O.random_state_dict(4)
    a = make_engine(ppgs_b200, sd, precision)
    b = ppgs_b200.Engine(0)
    b.blob().copy_(a.blob())
    torch.cuda.synchronize()
    b.adopt_blob()
    b.precision = precision
    audio = O.synthetic_audio(2, 32000, 3).cuda()
    assert torch.equal(a.from_audio(audio), b.from_audio(audio))


def test_config4_causal_windows(ppgs_b200):
    """BASELINE config 4: config/causal_transformer.py (IS_CAUSAL=True), 256 windows of 160
    frames.  The reference has no streaming state: 256 independent windows with
    a square subsequent mask + key padding."""
    sd = O.random_state_dict(2, peaky=True)
    engine = make_engine(ppgs_b200, sd, 'f16x2', is_causal=True)
    audio = O.synthetic_audio(256, 160 * 160, 11)
    out = engine.from_audio(audio.cuda())
    engine.check()
    assert out.shape == (256, 40, 160) and torch.isfinite(out).all()
    assert (out.sum(1) - 1).abs().max() <= 1e-5
    rows = [0, 97, 255]
    ref = O.from_audio(sd, audio[rows], is_causal=True)
    assert (out[rows].cpu() - ref).abs().max() <= PPG_TOL
    # causality: perturbing the last second of a window leaves the early posteriors alone,
    # up to the 2 + 2 frames of look-ahead of the k=5 convolutions and the STFT window
    changed = audio.clone()
    changed[:, :, -16000:] = 0
    early = engine.from_audio(changed.cuda())[:, :, :50]
    assert (early - out[:, :, :50]).abs().max() <= 1e-6
    non_causal = make_engine(ppgs_b200, sd, 'f16x2')
    a = non_causal.from_audio(audio[:4].cuda())[:, :, :50]
    b = non_causal.from_audio(changed[:4].cuda())[:, :, :50]
    assert (a - b).abs().max() > 1e-4


def test_config5_sharded_files_two_ranks(ppgs_b200, tmp_path):
    """BASELINE config 5 in miniature: `from_files_to_files` sharded over two ranks under
    torchrun (both on cuda:0 here, gloo for the one weight-blob broadcast): every file is
    written once and equals the single-process batched result."""
    import subprocess
    import sys
    from conftest import ROOT
    from test_host_logic import free_port, write_wav
    sd = O.random_state_dict(3)
    ckpt = tmp_path / 'ckpt.pt'
    torch.save({'model': sd}, ckpt)
    lengths = [16000, 40000, 90000, 8000, 16161, 32000, 64000]
    files = []
    for i, n in enumerate(lengths):
        files.append(str(tmp_path / f'{i}.wav'))
        write_wav(files[-1], n, seed=20 + i)
    single = [str(tmp_path / f'{i}-single.pt') for i in range(len(files))]
    ppgs_b200.from_files_to_files(files, single, checkpoint=ckpt, num_workers=2, gpu=0, max_frames=900)
    sharded = [str(tmp_path / f'{i}-sharded.pt') for i in range(len(files))]
    worker = tmp_path / 'worker.py'
    worker.write_text(
        'import sys, json\n'
        'from ppgs_b200 import parallel\n'
        'args = json.loads(sys.argv[1])\n'
        'parallel.from_files_to_files(args["files"], args["out"], checkpoint=args["ckpt"],\n'
        '                             num_workers=2, max_frames=900)\n')
    import json
    payload = json.dumps({'files': files, 'out': sharded, 'ckpt': str(ckpt)})
    env = dict(os.environ, PYTHONPATH=ROOT, PPGS_B200_DIST_BACKEND='gloo', LOCAL_RANK_GPU='0')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()), str(worker), payload]
    subprocess.run(cmd, check=True, timeout=300, env=env, cwd=ROOT)
    for a, b, n in zip(single, sharded, lengths):
        x, y = torch.load(a), torch.load(b)
        assert x.shape == (40, n // 160)
        assert torch.equal(x, y)


def test_cuda_graph_replay_is_bitwise_the_launch_path(ppgs_b200):
    """Steady-state loops replay the forward as one CUDA graph: same bits as 29 launches, the
    replay stops as soon as anything else touches the workspace or the inputs change, and
    launches keep being counted."""
    sd = O.random_state_dict(9, peaky=True)
    engine = make_engine(ppgs_b200, sd, 'f16x2')
    audio = O.synthetic_audio(3, 160 * 620, 21).cuda()
    other = O.synthetic_audio(3, 160 * 620, 22).cuda()
    engine.set_graphs(False)
    want, want_other = engine.from_audio(audio), engine.from_audio(other)
    per_forward = engine.launches
    engine.from_audio(audio)
    per_forward = engine.launches - per_forward
    engine.set_graphs(True)
    out = torch.empty_like(want)

    def forward(x):
        # fixed output buffer, like a serving loop (the public wrapper allocates per call)
        from ppgs_b200 import _lib
        _lib.check(_lib.lib.ppgs_from_audio(
            engine._handle, ctypes.c_void_p(x.data_ptr()), 3, x.shape[-1], x.shape[-1], None, 1, 0,
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out.clone()

    x = audio.squeeze(1).contiguous()
    y = other.squeeze(1).contiguous()
    assert torch.equal(forward(x), want)            # launch path, makes the plan resident
    assert engine.graph_replays == 0
    before = engine.launches
    assert torch.equal(forward(x), want)            # captured + launched as a graph
    assert torch.equal(forward(x), want)            # replayed
    assert engine.graph_replays == 2 and engine.launches - before == 2 * per_forward
    x.copy_(y)                                      # same buffers, new contents: still a replay
    assert torch.equal(forward(x), want_other)
    assert engine.graph_replays == 3
    assert torch.equal(forward(y), want_other)      # other input buffer: its own graph
    # something else uses the workspace (other shape): the next call must not replay blindly
    engine.from_audio(O.synthetic_audio(1, 160 * 100, 3).cuda())
    replays = engine.graph_replays
    assert torch.equal(forward(y), want_other) and engine.graph_replays == replays
    assert torch.equal(forward(y), want_other) and engine.graph_replays == replays + 1
    # profiling and precision changes drop back to launches
    engine.set_profiling(True)
    assert torch.equal(forward(y), want_other) and engine.graph_replays == replays + 1
    engine.set_profiling(False)
    engine.precision = 'f16'
    loose = forward(y)
    assert (loose - want_other).abs().max() > 0
    engine.precision = 'f16x2'
    assert torch.equal(forward(y), want_other)
    engine.check()


@pytest.mark.parametrize('seed', range(8))
def test_from_audio_random_batches_vs_oracle(ppgs_b200, seed):
    """Seeded random batch shapes (1-7 utterances, ragged lengths from 3 frames to 3.2 chunks,
    sample counts that are not multiples of the hop) against the oracle, device and host
    entry points, repeated so that the CUDA-graph cache is exercised on odd shapes too."""
    rng = np.random.default_rng(1000 + seed)
    batch = int(rng.integers(1, 8))
    longest = int(rng.choice([437, 160 * 37 + 5, 160 * 400, 160 * 500 + 159, 160 * 777 + 80, 160 * 1600]))
    lengths = [longest] + [int(rng.integers(433, longest + 1)) for _ in range(batch - 1)]
    order = rng.permutation(batch)
    lengths = [lengths[i] for i in order]
    sd = O.random_state_dict(50 + seed, peaky=bool(seed % 2))
    engine = make_engine(ppgs_b200, sd, 'f16x2')
    audio = O.synthetic_audio(batch, longest, 70 + seed)
    for row, n in enumerate(lengths):
        audio[row, :, n:] = 0                      # zero padding, like the reference collate
    sample_lengths = torch.tensor(lengths)
    ref = O.from_audio(sd, audio, lengths=sample_lengths)
    frames = [n // 160 for n in lengths]
    pinned = audio.squeeze(1).contiguous().pin_memory()
    for repeat in range(3):
        out = engine.from_audio(audio.cuda(), lengths=sample_lengths).cpu()
        host = engine.from_audio_host(pinned, lengths=sample_lengths)
        assert out.shape == ref.shape == host.shape
        for row, n in enumerate(frames):
            assert (out[row, :, :n] - ref[row, :, :n]).abs().max() <= PPG_TOL
            assert torch.equal(host[row, :, :n], out[row, :, :n])
    engine.check()
