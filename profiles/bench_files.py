"""BASELINE config 5 in miniature on one GPU: `from_files_to_files` over N synthetic 10 s
16 kHz int16 WAVE files (tmpfs when available), native pipeline (ppgs_files_to_files) vs the
Python reader / writer threads on the same batches.  Prints one JSON line per arm.

    python profiles/bench_files.py [files=2048] [workers=16] [gpus=1]
"""
import json
import os
import shutil
import sys
import tempfile
import time
import wave

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

files = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
device = 0 if gpus == 1 else list(range(gpus))
root = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
try:
    rng = np.random.default_rng(0)
    base = (rng.uniform(-0.5, 0.5, 160000 + files) * 32767).astype(np.int16)
    audio_files, output_files = [], []
    for i in range(files):
        path = os.path.join(root, f'{i:06d}.wav')
        with wave.open(path, 'wb') as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(16000)
            f.writeframes(base[i:i + 160000].tobytes())
        audio_files.append(path)
        output_files.append(os.path.join(root, f'{i:06d}-ppg.pt'))
    checkpoint = os.path.join(root, 'ckpt.pt')
    torch.save({'model': O.random_state_dict(0, peaky=True)}, checkpoint)
    engine = ppgs_b200.load.model(checkpoint, 'mel', 0)
    for arm in ('native', 'python'):
        os.environ['PPGS_B200_NATIVE_FILES'] = '1' if arm == 'native' else '0'
        ppgs_b200.from_files_to_files(audio_files[:128 * gpus], output_files[:128 * gpus], checkpoint=checkpoint,
                                      num_workers=workers, gpu=device, max_frames=64000)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ppgs_b200.from_files_to_files(audio_files, output_files, checkpoint=checkpoint,
                                      num_workers=workers, gpu=device, max_frames=64000)
        torch.cuda.synchronize()
        seconds = time.perf_counter() - t0
        sample = torch.load(output_files[-1])
        assert sample.shape == (40, 1000)
        print(json.dumps({
            'arm': arm, 'gpus': gpus, 'files': files, 'workers': workers, 'seconds': round(seconds, 3),
            'files_per_sec': round(files / seconds, 1),
            'ppg_frames_per_sec': round(files * 1000 / seconds),
            'audio_hours_per_hour': round(files * 10 / seconds),
            'storage': root.split('/')[1], 'host_cores': os.cpu_count(),
            'precision': engine.precision}), flush=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
