"""BASELINE config 3 (w2v2fb, batch = 32 x 10 s) timing of the CUDA front-end + d=512 PPG
head, per-kernel — context for DESIGN.md, not the headline (see `bench.py --workload w2v2fb`)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402
from oracle import w2v2_oracle as W  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
front = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0))
front.load_w2v2_state_dict(W.random_state_dict(0))
head = ppgs_b200.Engine(0, input_channels=768, hidden_channels=512).load_state_dict(
    O.random_state_dict(1, input_channels=768, hidden_channels=512))
audio = O.synthetic_audio(batch, 160000, 0).cuda()
lengths = torch.full((batch,), 1000)


def step():
    feats = front.w2v2fb(audio)
    return head.transformer(feats, lengths)


for _ in range(2):
    step()
front.set_profiling(True)
head.set_profiling(True)
torch.cuda.synchronize()
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
steps = 3
for _ in range(steps):
    out = step()
stop.record()
torch.cuda.synchronize()
ms = start.elapsed_time(stop) / steps
print(f'w2v2fb {batch}x10s: {ms:.1f} ms/step, {batch * 1000 / ms * 1e3:.0f} frames/s, '
      f'algorithmic {batch * 198.6e9 / ms / 1e9:.1f} TFLOP/s')
for name, engine in (('front-end', front), ('ppg head', head)):
    for kernel, (total, launches) in sorted(engine.kernel_stats().items(), key=lambda kv: -kv[1][0]):
        print(f'  {name:9s} {kernel:24s} {total / steps:8.2f} ms/step  ({launches // steps} launches)')
