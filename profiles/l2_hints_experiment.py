"""L2 eviction hints (PPGS_B200_L2_HINTS): TMA loads of operands that die with the kernel (attention's Q / K / V,
the projection's A operand, residual rows) are issued evict-first.  ms/step and per-kernel times at BASELINE
config 2 (64 x 10 s), alternating the setting in one process; outputs must be bitwise equal."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

audio = [O.synthetic_audio(64, 160000, 7 + i).squeeze(1).cuda() for i in range(3)]
sd = O.random_state_dict(0, peaky=True)
reference = None
for setting in (1, 0, 1, 0, 1, 0):
    os.environ['PPGS_B200_L2_HINTS'] = str(setting)
    engine = ppgs_b200.Engine(0).load_state_dict(sd)
    out = engine.from_audio(audio[0]).clone()
    if reference is None:
        reference = out
    same = bool(torch.equal(out, reference))
    for i in range(10):
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
    engine.set_profiling(True)
    for i in range(10):
        engine.from_audio(audio[i % 3])
    torch.cuda.synchronize()
    stats = {k: round(v[0] / 10, 4) for k, v in engine.kernel_stats().items()}
    engine.set_profiling(False)
    print(json.dumps({'l2_hints': setting, 'bitwise_equal': same, 'ms_per_step_median': sorted(times)[2],
                      'ms_per_step_min': min(times), 'kernels_ms': stats}), flush=True)
