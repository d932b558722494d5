"""Where does the posterior error come from?  GPU (f16x2 / fp32 modes) vs the oracle in
fp32 and in fp64 on BASELINE config-2 shaped rows (peaky model)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count())
for seed in (0, 1, 2, 3):
    sd = O.random_state_dict(seed, peaky=True)
    audio = O.synthetic_audio(2, 160000, 100 + seed)
    ref32 = O.from_audio(sd, audio)
    ref64 = O.from_audio(sd, audio, dtype=torch.float64).float()
    line = f'seed {seed}: oracle32 vs oracle64 {(ref32 - ref64).abs().max():.2e}'
    for precision in ('f16x2', 'fp32', 'f16'):
        engine = ppgs_b200.Engine(0).load_state_dict(sd)
        engine.precision = precision
        out = engine.from_audio(audio.cuda()).cpu()
        line += f' | {precision}: vs32 {(out - ref32).abs().max():.2e} vs64 {(out - ref64).abs().max():.2e}'
    print(line, flush=True)
