"""BASELINE config 1 (1 x 4 s utterance): latency of one `from_audio` through the C ABI with
pinned host buffers (H2D + 29 kernels + D2H, synchronous), CUDA-graph replay vs launches."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0, peaky=True))
engine.precision = 'f16x2'
for batch, samples in ((1, 64000), (1, 160000), (8, 64000)):
    audio = O.synthetic_audio(batch, samples, 0).squeeze(1).pin_memory()
    out = torch.empty(batch, 40, samples // 160).pin_memory()
    line = {'batch': batch, 'seconds': samples / 16000}
    for graphs in (False, True):
        engine.set_graphs(graphs)
        for _ in range(20):
            engine.from_audio_host(audio, out=out)
        torch.cuda.synchronize()
        times = []
        for _ in range(200):
            t0 = time.perf_counter()
            engine.from_audio_host(audio, out=out)      # returns after the D2H copy has landed
            times.append(time.perf_counter() - t0)
        times.sort()
        line['graph_replay_us' if graphs else 'launches_us'] = {
            'median': round(times[100] * 1e6, 1), 'p90': round(times[180] * 1e6, 1)}
    ref = O.from_audio(O.random_state_dict(0, peaky=True), audio.unsqueeze(1))
    line['max_abs_vs_oracle'] = (out - ref).abs().max().item()
    print(json.dumps(line), flush=True)
