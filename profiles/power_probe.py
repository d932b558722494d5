"""Is the config-2 step power-capped?  Runs the device-resident step for a few seconds per
setting while sampling `nvidia-smi` (SM clock, power, throttle reasons) every 50 ms, and
prints ms/step next to the median clock and power.  Settings come from the environment
(PPGS_B200_*), so one process = one setting."""
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0, peaky=True))
engine.precision = os.environ.get('PROBE_PRECISION', 'f16x2')
audio = [O.synthetic_audio(64, 160000, 7 + i).squeeze(1).cuda() for i in range(3)]
out = torch.empty(64, 40, 1000, device='cuda')
for i in range(20):
    engine.from_audio(audio[i % 3], out=out)
torch.cuda.synchronize()
lines = []
proc = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,'
                         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu',
                         '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(line) for line in proc.stdout], daemon=True).start()
steps = int(os.environ.get('PROBE_STEPS', '800'))
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
time.sleep(0.2)
start.record()
for i in range(steps):
    engine.from_audio(audio[i % 3], out=out)
stop.record()
torch.cuda.synchronize()
ms = start.elapsed_time(stop) / steps
time.sleep(0.1)
proc.terminate()
rows = [[p.strip() for p in line.split(',')] for line in lines if line.count(',') >= 5]
busy = [r for r in rows if float(r[1]) > 400]
print(json.dumps({
    'env': {k: v for k, v in os.environ.items() if k.startswith('PPGS_B200_') or k.startswith('PROBE_')},
    'ms_per_step': ms, 'samples_under_load': len(busy),
    'sm_mhz_median': statistics.median(float(r[0]) for r in busy) if busy else None,
    'power_w_median': statistics.median(float(r[1]) for r in busy) if busy else None,
    'power_cap_active_frac': sum(r[2].lower().startswith('active') for r in busy) / max(len(busy), 1),
    'temperature_c': max((float(r[5]) for r in busy), default=None)}), flush=True)
