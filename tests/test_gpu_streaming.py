"""Stateful streaming decoder (ppgs_stream_*) against the oracle: every
frame a session emits equals the reference's un-chunked causal forward
(`legacy_mode=True`, IS_CAUSAL=True; ppgs/model/transformer.py:65-81) of the WHOLE
utterance, whatever the push sizes — to the north-star tolerance of 1e-4."""
import pytest
import torch

from oracle import ppg_oracle as O

pytestmark = pytest.mark.gpu

PPG_TOL = 1e-4


@pytest.fixture(scope='module')
def ppgs_b200():
    import ppgs_b200
    return ppgs_b200


@pytest.fixture(scope='module')
def state():
    return O.random_state_dict(7, peaky=True)


@pytest.fixture(scope='module')
def engine(ppgs_b200, state):
    engine = ppgs_b200.Engine(0, is_causal=True).load_state_dict(state)
    engine.precision = 'f16x2'
    return engine


def features_and_reference(state, streams, frames, seed):
    audio = O.synthetic_audio(streams, frames * 160, seed)
    features = O.mel_from_audios(audio)
    lengths = torch.full((streams,), frames, dtype=torch.long)
    reference = O.from_features(state, features, lengths, is_causal=True, legacy_mode=True)
    return audio, features, reference


@pytest.mark.parametrize('streams,pushes', [
    (4, [160, 160, 160]),                       # BASELINE config 4: 160-frame chunks
    (2, [128, 128, 128, 126]),                  # tile-aligned pushes up to the capacity
    (3, [1, 2, 3, 4, 5, 37, 200, 1, 130, 127]),  # ragged pushes, odd stream count
    (1, [510]),                                 # one push
    (2, [300, 0, 0, 100]),                      # empty pushes
])
def test_streaming_equals_full_causal_forward(ppgs_b200, engine, state, streams, pushes):
    total = sum(pushes)
    _, features, reference = features_and_reference(state, streams, total, seed=total + streams)
    streamer = ppgs_b200.Streamer(engine, streams)
    assert streamer.capacity == 510
    at, pieces = 0, []
    for i, n in enumerate(pushes):
        final = i + 1 == len(pushes)
        out = streamer.push(features[..., at:at + n].cuda(), final=final)
        at += n
        expected_end = at if final else max(at - 4, 0)
        assert streamer.length == at and streamer.emitted == expected_end
        begin = expected_end - out.shape[-1]
        # every partial result is already the whole-utterance value (causality + look-ahead)
        if out.shape[-1]:
            assert (out.cpu() - reference[..., begin:expected_end]).abs().max() <= PPG_TOL
        pieces.append(out)
    result = torch.cat(pieces, dim=-1).cpu()
    assert result.shape == reference.shape
    assert (result - reference).abs().max() <= PPG_TOL
    assert (result.sum(1) - 1).abs().max() <= 1e-5
    engine.check()


def test_streaming_flush_reset_and_errors(ppgs_b200, engine, state):
    _, features, reference = features_and_reference(state, 2, 200, seed=3)
    streamer = ppgs_b200.Streamer(engine, 2)
    first = streamer.push(features[..., :200].cuda())
    assert first.shape[-1] == 196
    tail = streamer.push(None, final=True)            # flush: the last 4 frames
    assert tail.shape[-1] == 4
    result = torch.cat((first, tail), dim=-1).cpu()
    assert (result - reference).abs().max() <= PPG_TOL
    with pytest.raises(RuntimeError, match='finalised'):
        streamer.push(features[..., :10].cuda())
    # the same session object serves the next utterance after reset, bit for bit
    streamer.reset()
    again = torch.cat((streamer.push(features.cuda()), streamer.push(None, final=True)), dim=-1).cpu()
    assert torch.equal(again, result)
    # logits instead of posteriors
    streamer.reset()
    logits = streamer.push(features.cuda(), final=True, softmax=False).cpu()
    expected = O.from_features(state, features, torch.tensor([200, 200]), softmax=False,
                               is_causal=True, legacy_mode=True)
    assert (logits - expected).abs().max() <= 2e-3
    assert (torch.softmax(logits, 1) - reference).abs().max() <= PPG_TOL
    # capacity: ValueError('size is too large') like ppgs/model/transformer.py:103-104
    streamer.reset()
    streamer.push(torch.zeros(2, 80, 500, dtype=torch.float16).cuda())
    with pytest.raises(ValueError, match='size is too large'):
        streamer.push(torch.zeros(2, 80, 11, dtype=torch.float16).cuda())
    with pytest.raises(ValueError, match='expected features'):
        streamer.push(torch.zeros(3, 80, 1, dtype=torch.float16).cuda())
    # streaming needs a causal model
    plain = ppgs_b200.Engine(0).load_state_dict(state)
    with pytest.raises(ValueError, match='causal'):
        ppgs_b200.Streamer(plain, 2)


def test_streaming_is_incremental(ppgs_b200, engine, state):
    """A push recomputes only the row tiles from 4 frames before the previous end: the
    launch count of a late push equals an early push's, and sessions do not disturb the
    engine's batch API."""
    _, features, reference = features_and_reference(state, 2, 480, seed=5)
    streamer = ppgs_b200.Streamer(engine, 2)
    engine.set_profiling(True)
    streamer.push(features[..., :160].cuda())
    streamer.push(features[..., 160:320].cuda())
    batch = engine.transformer(features.cuda(), torch.tensor([480, 480]), legacy_mode=True).cpu()
    out = streamer.push(features[..., 320:].cuda(), final=True).cpu()
    stats = engine.kernel_stats()
    engine.set_profiling(False)
    assert (out - reference[..., 316:]).abs().max() <= PPG_TOL
    assert (batch - reference).abs().max() <= PPG_TOL
    assert stats['stream_append'][1] == 3 and stats['tc_conv_out_softmax'][1] == 4


def test_streaming_from_audio(ppgs_b200, engine, state):
    """push_audio: mel frames are produced as their 1024-sample windows complete; the
    concatenated result equals the oracle's from_audio of the whole utterance."""
    audio, features, reference = features_and_reference(state, 2, 300, seed=11)
    streamer = ppgs_b200.Streamer(engine, 2)
    cuts = [0, 100, 1000, 1700, 16000, 16001, 30000, 47999, 48000]
    pieces = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        pieces.append(streamer.push_audio(audio[..., a:b].cuda(), final=i + 2 == len(cuts)))
    result = torch.cat(pieces, dim=-1).cpu()
    assert result.shape == reference.shape
    assert (result - reference).abs().max() <= PPG_TOL
    # and the streamed features are the batch front-end's, bit for bit
    assert streamer.length == 300


@pytest.mark.parametrize('total,pushes', [
    (501, [501]),
    (801, [160] * 5 + [1]),
    (1234, [400, 54, 46, 300, 434]),
    (1700, [37] * 45 + [35]),
])
def test_long_streamer_equals_chunked_reference(ppgs_b200, engine, state, total, pushes):
    """Unbounded streaming = the reference's chunked inference (500 / 400 / 50,
    ppgs/model/transformer.py:49-64) computed incrementally by two alternating sessions."""
    assert sum(pushes) == total
    audio = O.synthetic_audio(2, total * 160, total)
    features = O.mel_from_audios(audio)
    lengths = torch.full((2,), total, dtype=torch.long)
    reference = O.from_features(state, features, lengths, is_causal=True)   # chunked
    streamer = ppgs_b200.LongStreamer(engine, 2)
    at, pieces = 0, []
    for i, n in enumerate(pushes):
        out = streamer.push(features[..., at:at + n].cuda(), final=i + 1 == len(pushes))
        begin = streamer.emitted - out.shape[-1]
        if out.shape[-1]:
            assert (out.cpu() - reference[..., begin:streamer.emitted]).abs().max() <= PPG_TOL
        at += n
        pieces.append(out)
    result = torch.cat(pieces, dim=-1).cpu()
    assert result.shape == reference.shape
    assert (result - reference).abs().max() <= PPG_TOL
    with pytest.raises(RuntimeError, match='finalised'):
        streamer.push(features[..., :1].cuda())
    # the batch API on the same features agrees too (same chunking, folded into the batch axis)
    batch = engine.transformer(features.cuda(), lengths).cpu()
    assert (batch - result).abs().max() <= 2e-5
    engine.check()


def test_independent_streams_join_finish_and_restart(ppgs_b200, engine, state):
    """Serving: four streams with their own utterances, push sizes, end points and restarts.
    Whatever the interleaving, each utterance's frames equal its own whole-utterance causal
    forward."""
    import random
    random.seed(4)
    streams = 4
    streamer = ppgs_b200.Streamer(engine, streams)
    totals = [[300, 77], [510], [40, 40, 200], [129, 256]]          # utterances per stream
    features, references = {}, {}
    for b, lengths in enumerate(totals):
        for u, total in enumerate(lengths):
            _, f, r = features_and_reference(state, 1, total, seed=100 * b + u)
            features[b, u], references[b, u] = f[0], r[0]
    utterance = [0] * streams                     # current utterance of each stream
    position = [0] * streams
    collected = {key: [] for key in features}
    steps = 0
    while any(utterance[b] < len(totals[b]) for b in range(streams)):
        steps += 1
        counts, finals = [], []
        for b in range(streams):
            if utterance[b] >= len(totals[b]):
                counts.append(0)
                finals.append(False)
                continue
            total = totals[b][utterance[b]]
            n = min(random.choice([0, 1, 7, 33, 64, 128, 160]), total - position[b])
            counts.append(n)
            finals.append(position[b] + n == total)
        width = max(max(counts), 1)
        chunk = torch.zeros(streams, 80, width, dtype=torch.float16)
        for b, n in enumerate(counts):
            if n:
                chunk[b, :, :n] = features[b, utterance[b]][:, position[b]:position[b] + n]
        out, produced = streamer.push(chunk.cuda(), final=finals, lengths=counts)
        lengths_now, emitted_now = streamer.state()
        finished = []
        for b in range(streams):
            if utterance[b] >= len(totals[b]):
                assert produced[b] == 0
                continue
            position[b] += counts[b]
            assert lengths_now[b] == position[b]
            collected[b, utterance[b]].append(out[b, :, :produced[b]].cpu())
            if finals[b]:
                assert emitted_now[b] == totals[b][utterance[b]]
                finished.append(b)
        if finished:                                # those streams start their next utterance
            streamer.reset(streams=finished)
            for b in finished:
                utterance[b] += 1
                position[b] = 0
    assert steps > 10
    for key, pieces in collected.items():
        result = torch.cat(pieces, dim=-1)
        assert result.shape == references[key].shape
        assert (result - references[key]).abs().max() <= PPG_TOL, key
    with pytest.raises(ValueError, match='brings'):
        streamer.push(torch.zeros(streams, 80, 2, dtype=torch.float16).cuda(), lengths=[3, 0, 0, 0])
    engine.check()


def test_stream_reset_isolates_utterances(ppgs_b200, engine, state):
    """Non-finite features in one utterance neither reach the other stream nor the next
    utterance of the same stream after its reset."""
    _, features, reference = features_and_reference(state, 2, 200, seed=17)
    streamer = ppgs_b200.Streamer(engine, 2)
    poisoned = features.clone()
    poisoned[0, :, 60:90] = float('nan')
    out, produced = streamer.push(poisoned.cuda(), final=[True, True], lengths=[200, 200])
    assert produced == [200, 200]
    assert not torch.isfinite(out[0]).all()
    assert (out[1].cpu() - reference[1]).abs().max() <= PPG_TOL      # the neighbour is untouched
    streamer.reset(streams=[0])
    out, produced = streamer.push(features[:, :, :130].cuda(), final=[True, False], lengths=[130, 0])
    assert produced == [130, 0]
    _, _, short = features_and_reference(state, 2, 200, seed=17)
    expected = O.from_features(state, features[:1, :, :130], torch.tensor([130]), is_causal=True,
                               legacy_mode=True)
    assert (out[0, :, :130].cpu() - expected[0]).abs().max() <= PPG_TOL
    engine.check()
