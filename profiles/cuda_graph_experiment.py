"""A/B in one process: the fused forward through the engine's CUDA-graph cache
(ppgs_engine_set_graphs) vs the 29-launch path, alternating, 64 x 10 s.  The GPU runs under
its power cap, so only interleaved measurements are comparable."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from ppgs_b200 import _lib  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0, peaky=True))
engine.precision = 'f16x2'
audio = [O.synthetic_audio(64, 160000, i).squeeze(1).cuda() for i in range(4)]
out = torch.empty(64, 40, 1000, device='cuda')


def forward(i):
    x = audio[i % 4]
    _lib.check(_lib.lib.ppgs_from_audio(
        engine._handle, ctypes.c_void_p(x.data_ptr()), 64, 160000, 160000, None, 1, 0,
        ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def timed(n=40):
    for i in range(8):
        forward(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        forward(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for round_ in range(4):
    engine.set_graphs(False)
    launches = timed()
    engine.set_graphs(True)
    before = engine.graph_replays
    graphs = timed()
    print(f'round {round_}: launches {launches:.3f} ms   graph replay {graphs:.3f} ms '
          f'({engine.graph_replays - before} replays)')
