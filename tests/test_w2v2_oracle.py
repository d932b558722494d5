"""The w2v2fb oracle restatement (oracle/w2v2_oracle.py) against golden outputs of the real
Hugging Face Wav2Vec2Model (oracle/make_golden_w2v2.py) and, when `transformers` is
importable, against the live module on a fresh seed."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import w2v2_oracle as W
from oracle.make_golden_w2v2 import case_inputs

CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, 'w2v2fb_*.npz')))


def close_fp16(a, b):
    """Two fp32 evaluations of the 12-layer encoder differ by ~1e-5; after the fp16 cast that
    is at most one fp16 step of the value (2^-10 relative), or ~1e-5 absolute near zero."""
    a, b = a.astype(np.float32), b.astype(np.float32)
    return np.abs(a - b) <= 2e-5 + 2.0 ** -10 * np.abs(b)


@pytest.mark.parametrize('name', CASES)
def test_oracle_vs_hf_golden(name):
    g = golden(name)
    sd = W.random_state_dict(int(g['weight_seed']))
    audio, lengths = case_inputs(int(g['samples']), g['lengths'].tolist(), int(g['audio_seed']))
    feats = W.from_audios(sd, audio, lengths).numpy()
    assert feats.shape == g['features'].shape and feats.dtype == np.float16
    assert close_fp16(feats, g['features']).all()
    assert (feats != g['features']).mean() <= 5e-2


def test_feature_lengths_and_upsample_index():
    assert W.feature_lengths(torch.tensor([160080, 16080, 400])).tolist() == [500, 50, 1]
    h = torch.arange(499, dtype=torch.float32)[None, None]
    up = W.nearest_upsample(h, 1000)[0, 0]
    ref = torch.nn.functional.interpolate(h, size=1000, mode='nearest')[0, 0]
    assert torch.equal(up, ref)


def test_oracle_vs_live_transformers():
    pytest.importorskip('transformers')
    from oracle.make_golden_w2v2 import reference_features
    sd = W.random_state_dict(5)
    audio, lengths = case_inputs(4800, [4800, 3000, 401], 9)
    ref = reference_features(sd, audio, lengths).numpy()
    mine = W.from_audios(sd, audio, lengths).numpy()
    assert close_fp16(mine, ref).all()
    assert (mine != ref).mean() <= 5e-2
