"""Generate tests/golden/w2v2fb_*.npz by calling the reference's OWN function,
`ppgs.preprocess.w2v2fb.from_audios` (ppgs/preprocess/w2v2fb/core.py:32-75, imported
unmodified under oracle/refshim.py), with the third-party model it would download injected
into its cache: a seeded `transformers.Wav2Vec2Model(Wav2Vec2Config())`
— the wav2vec2-base architecture; pretrained weights are not available offline.

    python -m oracle.make_golden_w2v2

Stores the upsampled fp16 features the reference returns, i.e. exactly the tensor the CUDA
front-end has to reproduce.  `restated_features` (the same lines written out around the HF
module, what round 1 used) must agree bit for bit — asserted here."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ppg_oracle as O  # noqa: E402
from oracle import w2v2_oracle as W  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# (name, weight seed, audio seed, samples, lengths)
CASES = [
    ('w2v2fb_2x16000_s0', 0, 3, 16000, [16000, 9000]),
    ('w2v2fb_1x8333_s1', 1, 4, 8333, [8333]),
]


def hf_model(sd):
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    model = Wav2Vec2Model(Wav2Vec2Config()).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert missing == ['masked_spec_embed'] and not unexpected
    return model


def reference_features(sd, audio, lengths):
    """The reference's own `ppgs.preprocess.w2v2fb.from_audios` with the model injected."""
    from oracle import refshim
    ppgs = refshim.import_reference()
    fn = ppgs.preprocess.w2v2fb.from_audios
    fn.model, fn.device = hf_model(sd), torch.device('cpu')
    return fn(audio, lengths, sample_rate=ppgs.SAMPLE_RATE, gpu=None)


def restated_features(sd, audio, lengths):
    """ppgs/preprocess/w2v2fb/core.py:52-75 written out around the real HF module."""
    model = hf_model(sd)
    pad = W.W2V2_PAD
    padded = torch.nn.functional.pad(audio, (pad, pad)).squeeze(1)
    mask = (torch.arange(padded.shape[-1])[None] < (lengths + 2 * pad)[:, None]).long()
    with torch.no_grad():
        hidden = model(padded, mask).last_hidden_state
        up = torch.nn.functional.interpolate(
            hidden.transpose(1, 2), size=audio.shape[-1] // O.HOPSIZE, mode='nearest')
    return up.to(torch.float16)


def case_inputs(samples, lengths, audio_seed):
    audio = O.synthetic_audio(len(lengths), samples, audio_seed)
    lengths = torch.tensor(lengths)
    for row, n in enumerate(lengths.tolist()):
        audio[row, :, n:] = 0          # collate zero-pads (ppgs/data/collate.py:19-28)
    return audio, lengths


def main():
    for name, wseed, aseed, samples, lengths in CASES:
        sd = W.random_state_dict(wseed)
        audio, lens = case_inputs(samples, lengths, aseed)
        feats = reference_features(sd, audio, lens)
        assert torch.equal(feats, restated_features(sd, audio, lens))
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + '.npz'), features=feats.numpy(),
                            weight_seed=wseed, audio_seed=aseed, samples=samples,
                            lengths=np.array(lengths))
        print(name, tuple(feats.shape), float(feats.float().abs().max()))


if __name__ == '__main__':
    main()
