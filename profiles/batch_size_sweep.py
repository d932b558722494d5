"""ms per forward and per utterance against the batch size (10 s utterances, 5 row-tile pairs each):
batches whose tile-pair count fills whole rounds of the 74 CTA pairs (29 -> 145 = 1.96 rounds, 44 -> 220 =
2.97, 59 -> 295 = 3.99) against the bench's 64 (320 = 4.32 rounds).  Smaller batches also have smaller
kernel-to-kernel hand-overs (more L2 hits): per-utterance time tells whether running the forward on
L2-sized, wave-aligned parts of the batch would pay."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0, peaky=True))
pool = [O.synthetic_audio(64, 160000, 7 + i).squeeze(1).cuda() for i in range(3)]
for batch in (64, 59, 44, 32, 29, 14, 64, 29):
    audio = [a[:batch].contiguous() for a in pool]
    for i in range(8):
        engine.from_audio(audio[i % 3])
    times = []
    for _ in range(5):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for i in range(20):
            engine.from_audio(audio[i % 3])
        stop.record()
        torch.cuda.synchronize()
        times.append(start.elapsed_time(stop) / 20)
    ms = sorted(times)[2]
    print(json.dumps({'batch': batch, 'tile_pairs': 5 * batch, 'rounds_of_74': round(5 * batch / 74, 2),
                      'ms_per_forward': round(ms, 4), 'us_per_utterance': round(1000 * ms / batch, 2)}), flush=True)
