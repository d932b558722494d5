"""Experiment: 64 utterances as one batch vs two 32-utterance half-batches on two engines / two
streams (to fill the tails of the persistent kernels).  B200 result: 4.35 ms vs 4.43 ms -> no gain."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200
from oracle import ppg_oracle as O
sd = O.random_state_dict(0, peaky=True)
engines = [ppgs_b200.Engine(0).load_state_dict(sd) for _ in range(2)]
for e in engines: e.precision = 'f16x2'
audio = O.synthetic_audio(64, 160000, 0).cuda()
halves = [audio[:32].contiguous(), audio[32:].contiguous()]
streams = [torch.cuda.Stream() for _ in range(2)]
def one():
    engines[0].from_audio(audio)
def two():
    for e, h, s in zip(engines, halves, streams):
        with torch.cuda.stream(s):
            e.from_audio(h)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print('one engine, 64 utterances      ', timed(one))
print('two engines x 32 on two streams', timed(two))
quarters = [audio[i*16:(i+1)*16].contiguous() for i in range(4)]
