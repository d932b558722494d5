import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200
from oracle import ppg_oracle as O
from oracle import w2v2_oracle as W
audio = O.synthetic_audio(32, 160000, 0).cuda()
lengths = torch.full((32,), 1000)
ref = None
for setting in (1, 0, 1, 0):
    os.environ['PPGS_B200_SERPENTINE'] = str(setting)
    front = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0))
    front.load_w2v2_state_dict(W.random_state_dict(0))
    head = ppgs_b200.Engine(0, input_channels=768, hidden_channels=512).load_state_dict(
        O.random_state_dict(1, input_channels=768, hidden_channels=512))
    def step():
        return head.transformer(front.w2v2fb(audio), lengths)
    out = step().clone()
    if ref is None: ref = out
    same = bool(torch.equal(out, ref))
    for _ in range(3): step()
    times = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5): step()
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / 5)
    print(json.dumps({'serpentine': setting, 'bitwise_equal': same, 'ms_median': sorted(times)[2], 'ms_min': min(times)}), flush=True)
