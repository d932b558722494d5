"""Generate `tests/golden/*.npz` from the UNMODIFIED reference (dev container only).

    python -m oracle.make_golden

Runs the reference's own modules (`ppgs.preprocess.mel.from_audios`,
`ppgs.model.Transformer`, `ppgs.from_audio`) under `oracle/refshim.py` on seeded
synthetic inputs (`oracle.ppg_oracle.synthetic_audio` / `random_state_dict`, both
reproducible from the seed alone) and stores only the OUTPUTS.  Numerics mode O3
(SURVEY.md §8c): the reference modules in fp32 with autocast disabled — the
<=1e-4 target.  One as-shipped (O1, bf16 autocast) `ppgs.from_audio` output is
stored for context.
"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ppg_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402

GOLDEN_DIR = os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# (name, audio kind, batch, samples, audio seed)
MEL_CASES = [
    ('mel_noise_1x64000_s0', 'noise', 1, 64000, 0),
    ('mel_noise_2x16000_s1', 'noise', 2, 16000, 1),
    ('mel_speech_2x32000_s0', 'speech', 2, 32000, 0),
    ('mel_noise_1x160_s2', 'noise', 1, 160 * 3 + 17, 2),
]

# (name, weight seed, peaky, causal, frames, lengths, audio seed)
PPG_CASES = [
    ('ppg_T400_s0', 0, False, False, 400, [400], 0),
    ('ppg_T400_peaky_s1', 1, True, False, 400, [400, 250], 1),
    ('ppg_T500_ragged_s0', 0, False, False, 500, [500, 320, 7], 2),
    ('ppg_T501_s2', 2, False, False, 501, [501, 77], 3),
    ('ppg_T1000_s0', 0, True, False, 1000, [1000, 1000], 4),
    ('ppg_T1234_ragged_s1', 1, False, False, 1234, [1234, 900, 380, 10], 5),
    ('ppg_causal_T160_s0', 0, True, True, 160, [160, 160, 100], 6),
]


def audio_for(kind, batch, samples, seed):
    if kind == 'noise':
        return O.synthetic_audio(batch, samples, seed)
    return O.speechlike_audio(batch, samples, seed)


def main():
    ppgs = refshim.import_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)

    for name, kind, batch, samples, seed in MEL_CASES:
        audio = audio_for(kind, batch, samples, seed)
        lengths = torch.tensor([samples] * batch)
        mel = ppgs.preprocess.mel.from_audios(audio, lengths)   # no autocast
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + '.npz'),
            mel=mel.numpy(), kind=kind, batch=batch, samples=samples, seed=seed)
        print(name, tuple(mel.shape), mel.dtype)

    for name, wseed, peaky, causal, frames, lengths, aseed in PPG_CASES:
        sd = O.random_state_dict(wseed, peaky=peaky)
        model = ppgs.model.Transformer(is_causal=causal)
        model.load_state_dict(sd)
        model.eval()
        audio = O.synthetic_audio(len(lengths), frames * O.HOPSIZE, aseed)
        features = ppgs.preprocess.mel.from_audios(
            audio, torch.tensor([frames * O.HOPSIZE] * len(lengths)))
        with torch.inference_mode():
            logits = model(features.float(), torch.tensor(lengths))
            ppg = torch.softmax(logits, dim=1)
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + '.npz'),
            ppg=ppg.numpy(), logits=logits.numpy(), weight_seed=wseed,
            peaky=peaky, causal=causal, frames=frames, lengths=np.array(lengths),
            audio_seed=aseed)
        print(name, tuple(ppg.shape), float(ppg.max()))

    # As shipped (O1): ppgs.from_audio, gpu=None, bf16 autocast, B=1
    sd = O.random_state_dict(0)
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, 'ckpt.pt')
        torch.save({'model': sd}, ckpt)
        audio = O.synthetic_audio(1, 64000, 0)
        out = ppgs.from_audio(audio[0], 16000, representation='mel', checkpoint=ckpt)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, 'asshipped_from_audio_1x64000_s0.npz'),
        ppg=out.float().numpy(), dtype=str(out.dtype))
    print('asshipped', tuple(out.shape), out.dtype)


if __name__ == '__main__':
    main()
