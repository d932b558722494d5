"""w2v2fb parity tiers on the GPU: T1 feature flips vs the oracle, T3 end-to-end posterior
error (GPU features -> GPU PPG head vs oracle features -> oracle PPG head), for the
tensor-core encoder (default) and the all-fp32 encoder (PPGS_B200_W2V2_TC=0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppgs_b200  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402
from oracle import w2v2_oracle as W  # noqa: E402

w_sd = W.random_state_dict(3)
audio = O.synthetic_audio(2, 48000, 5)
lengths = torch.tensor([48000, 48000])
ref_feats = W.from_audios(w_sd, audio, lengths)
front = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0)).load_w2v2_state_dict(w_sd)
feats = front.w2v2fb(audio.cuda(), lengths)
f, r = feats.cpu().float().numpy(), ref_feats.float().numpy()
print(f'W2V2_TC={os.environ.get("PPGS_B200_W2V2_TC", "1")}: feature mismatch rate {(f != r).mean():.4f}, '
      f'max abs {np.abs(f - r).max():.2e}')
for peaky in (False, True):
    ppg_sd = O.random_state_dict(4, input_channels=768, hidden_channels=512, peaky=peaky)
    head = ppgs_b200.Engine(0, input_channels=768, hidden_channels=512).load_state_dict(ppg_sd)
    frames = torch.tensor([300, 300])
    out = head.transformer(feats, frames).cpu()
    ref = O.from_features(ppg_sd, ref_feats, frames)
    same = O.from_features(ppg_sd, feats.cpu(), frames)
    print(f'  peaky={peaky}: T3 end-to-end {float((out - ref).abs().max()):.2e}   '
          f'T2 same features {float((out - same).abs().max()):.2e}')
