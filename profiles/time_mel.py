"""Mel front-end alone at BASELINE config 2 (64 x 10 s): ms per launch of `ppgs_mel_forward`
(CUDA events around 50 launches over rotating inputs larger than L2 in total) and, through the
engine's per-kernel events, the mel / fold launches inside `from_audio`."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0))
audio = [torch.randn(64, 1, 160000, device='cuda') * 0.1 for _ in range(4)]
for i in range(8):
    engine.mel(audio[i % 4])
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
start.record()
for i in range(50):
    engine.mel(audio[i % 4])
stop.record()
torch.cuda.synchronize()
mel_ms = start.elapsed_time(stop) / 50
engine.set_profiling(True)
for i in range(10):
    engine.from_audio(audio[i % 4])
torch.cuda.synchronize()
stats = {k: round(v[0] / 10, 4) for k, v in engine.kernel_stats().items()}
print(json.dumps({'mel_forward_ms': round(mel_ms, 4), 'audio_gb_per_s': round(64 * 160000 * 4 / mel_ms / 1e6, 1),
                  'from_audio_kernels_ms': stats}))
