"""BASELINE config 4 (config/causal_transformer.py, 160-frame chunks, batch = 256) two ways:
(a) the reference's semantics — 256 independent 160-frame windows, no state;
(b) the stateful streaming decoder — 256 sessions, three 160-frame pushes each (480 frames
    of context at the end), every frame equal to the whole-utterance causal forward.
Device-resident features; CUDA events; prints JSON lines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

streams, chunk, pushes = 256, 160, 3
engine = ppgs_b200.Engine(0, is_causal=True).load_state_dict(O.random_state_dict(2, peaky=True))
engine.precision = 'f16x2'
features = O.mel_from_audios(O.synthetic_audio(8, chunk * pushes * 160, 0)).cuda()
features = features.repeat(streams // 8, 1, 1)                     # (256, 80, 480)
features = features + 0                                             # own storage
lengths = torch.full((streams,), chunk)


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(steps):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / steps


windows = [features[..., i * chunk:(i + 1) * chunk].contiguous() for i in range(pushes)]
ms = timed(lambda: [engine.transformer(w, lengths) for w in windows])
print(json.dumps({'arm': 'stateless windows (reference semantics)', 'streams': streams, 'chunk': chunk,
                  'ms_per_480_frames': round(ms, 3),
                  'frames_per_sec': round(streams * chunk * pushes / ms * 1e3)}))

streamer = ppgs_b200.Streamer(engine, streams)


def session():
    streamer.reset()
    for i, w in enumerate(windows):
        streamer.push(w, final=i + 1 == pushes)


ms = timed(session)
engine.set_profiling(True)
session()
stats = engine.kernel_stats()
engine.set_profiling(False)
print(json.dumps({'arm': 'stateful streaming (480-frame context)', 'streams': streams, 'chunk': chunk,
                  'ms_per_480_frames': round(ms, 3),
                  'frames_per_sec': round(streams * chunk * pushes / ms * 1e3),
                  'ms_per_push': round(ms / pushes, 3),
                  'kernels_ms': {k: round(v[0], 3) for k, v in sorted(stats.items())}}))
full = timed(lambda: engine.transformer(features, torch.full((streams,), chunk * pushes), legacy_mode=True))
print(json.dumps({'arm': 'one un-chunked causal forward of 480 frames (offline lower bound)',
                  'ms_per_480_frames': round(full, 3),
                  'frames_per_sec': round(streams * chunk * pushes / full * 1e3)}))
