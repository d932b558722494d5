"""Bookkeeping of `LongStreamer` without a GPU: with the device sessions replaced by a
recorder, every un-padded frame g must come out exactly once, in order, produced by chunk
g // 400 at local position g % 400 + 50 — the plan of ppgs/model/transformer.py:49-64 —
whatever the push sizes."""
import random

import pytest
import torch

from ppgs_b200 import streaming


class FakeEngine:
    device = torch.device('cpu')

    class cfg:
        input_channels = 1
        output_channels = 1

    def _on_device(self, tensor, dtype):
        return tensor.to(dtype)


class RecordingSession:
    """Stands in for `Streamer`: emits (frame id * 1000 + local index) for the frames that
    became final under the 4-frame look-ahead rule."""
    LOOKAHEAD = 4

    def __init__(self, engine, streams):
        self.reset()

    def reset(self):
        self.frames, self.emitted, self.final = [], 0, False

    @property
    def length(self):
        return len(self.frames)

    def push(self, features=None, final=False, softmax=True):
        assert not self.final
        if features is not None:
            self.frames += features[0, 0].tolist()
        assert len(self.frames) <= 510, 'session capacity'
        end = len(self.frames) if final else max(len(self.frames) - 4, self.emitted)
        out = torch.tensor([[[self.frames[i] * 1000 + i for i in range(self.emitted, end)]]],
                           dtype=torch.float32)
        self.emitted, self.final = end, final
        return out


@pytest.mark.parametrize('total', [1, 3, 400, 401, 450, 454, 500, 501, 799, 800, 801, 850, 851,
                                   1234, 1654, 2000, 2047])
def test_long_streamer_chunk_plan(monkeypatch, total):
    monkeypatch.setattr(streaming, 'Streamer', RecordingSession)
    for trial in range(6):
        random.seed(total * 10 + trial)
        streamer = streaming.LongStreamer(FakeEngine(), 1)
        x = torch.arange(1, total + 1, dtype=torch.float32).reshape(1, 1, total)
        at, outs = 0, []
        while at < total:
            n = random.choice([1, 2, 5, 50, 54, 100, 160, 399, 400, 401, 460, 1000]) if trial else total
            n = min(n, total - at)
            outs.append(streamer.push(x[..., at:at + n], final=at + n == total and trial % 2 == 0))
            at += n
        if not streamer.closed:
            outs.append(streamer.push(None, final=True))
        result = torch.cat(outs, -1)[0, 0].tolist()
        assert len(result) == total == streamer.emitted
        for g, value in enumerate(result):
            frame, local = divmod(int(value), 1000)
            assert frame == g + 1 and local == g % 400 + 50
