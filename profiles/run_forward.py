"""Profiling driver: `forwards` passes of the hot path at BASELINE config 2 (64 x 10 s),
for `ncu` launch lists / full captures (never a bench number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402  (synthetic inputs only)

forwards = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0, peaky=True))
engine.precision = os.environ.get('PPGS_B200_PRECISION', 'f16x2')
audio = O.synthetic_audio(batch, 160000, 0).cuda()
for _ in range(forwards):
    out = engine.from_audio(audio)
torch.cuda.synchronize()
print('ok', tuple(out.shape), engine.launches)
