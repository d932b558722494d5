"""Round 2: Q and K are kept as ONE fp16 plane for the attention MMAs; how many MMA passes do
their columns of the QKV GEMM need?  PPGS_B200_QK_GEMM_PASSES = 3 (hi.hi + x_hi.W_lo + x_lo.W_hi),
2 (hi.hi + x_lo.W_hi) or 1 (hi.hi): max-abs error of the f16x2 path against the fp32 oracle on
IDENTICAL features and end to end from audio (peaky models, 5 seeds), plus per-kernel times at
BASELINE config 2 (64 x 10 s).  Writes one JSON line per setting.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count())
SEEDS = (0, 1, 2, 3, 4)
cases = []
for seed in SEEDS:
    sd = O.random_state_dict(seed, peaky=True)
    audio = O.synthetic_audio(4, 160000, 100 + seed)
    feats = O.mel_from_audios(audio)
    lengths = torch.tensor([1000, 1000, 777, 401])
    cases.append((sd, audio, feats, lengths, O.from_features(sd, feats, lengths), O.from_audio(sd, audio)))

bench_audio = [O.synthetic_audio(64, 160000, 7 + i).squeeze(1).cuda() for i in range(3)]

for passes in (3, 2, 1, 3, 1):
    os.environ['PPGS_B200_QK_GEMM_PASSES'] = str(passes)
    err_feat, err_audio = [], []
    for sd, audio, feats, lengths, ref_f, ref_a in cases:
        engine = ppgs_b200.Engine(0).load_state_dict(sd)
        engine.precision = 'f16x2'
        out = engine.transformer(feats.cuda(), lengths).cpu()
        err_feat.append((out - ref_f).abs().max().item())
        out = engine.from_audio(audio.cuda()).cpu()
        err_audio.append((out - ref_a).abs().max().item())
    for i in range(6):
        engine.from_audio(bench_audio[i % 3])
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for i in range(20):
        engine.from_audio(bench_audio[i % 3])
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / 20
    engine.set_profiling(True)
    for i in range(10):
        engine.from_audio(bench_audio[i % 3])
    torch.cuda.synchronize()
    stats = {k: round(v[0] / 10, 4) for k, v in engine.kernel_stats().items()}
    engine.set_profiling(False)
    print(json.dumps({'qk_gemm_passes': passes, 'err_from_features': err_feat,
                      'err_from_audio': err_audio, 'ms_per_step': ms, 'kernels_ms': stats}), flush=True)
