"""The oracle restatement (oracle/ppg_oracle.py) against the committed golden
vectors, which are OUTPUTS OF THE REFERENCE'S OWN MODULES (oracle/make_golden.py,
generated in the dev container from /root/reference under oracle/refshim.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import ppg_oracle as O

MEL = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, 'mel_*.npz')))
PPG = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, 'ppg_*.npz')))


def mel_case_audio(g):
    kind, batch = str(g['kind']), int(g['batch'])
    samples, seed = int(g['samples']), int(g['seed'])
    if kind == 'noise':
        return O.synthetic_audio(batch, samples, seed)
    return O.speechlike_audio(batch, samples, seed)


def ppg_case_inputs(g):
    sd = O.random_state_dict(int(g['weight_seed']), peaky=bool(g['peaky']))
    lengths = torch.as_tensor(g['lengths']).long()
    audio = O.synthetic_audio(len(lengths), int(g['frames']) * O.HOPSIZE, int(g['audio_seed']))
    return sd, audio, lengths


@pytest.mark.parametrize('name', MEL)
def test_mel_bit_exact(name):
    """T1: fp32-op restatement with the two fp16 casts == reference mel, bit for bit."""
    g = golden(name)
    mel = O.mel_from_audios(mel_case_audio(g)).numpy()
    assert mel.dtype == np.float16 and mel.shape == g['mel'].shape
    assert np.array_equal(mel.view(np.uint16), g['mel'].view(np.uint16))


@pytest.mark.parametrize('name', PPG)
def test_ppg_matches_reference_modules(name):
    """T2/T3: oracle posteriors and logits vs the reference Transformer (fp32,
    autocast off).  Tolerance 2e-5: both sides are fp32 with different summation
    orders (explicit matmuls vs torch's fused kernels)."""
    g = golden(name)
    sd, audio, lengths = ppg_case_inputs(g)
    feats = O.mel_from_audios(audio)
    causal = bool(g['causal'])
    ppg = O.from_features(sd, feats, lengths, is_causal=causal).numpy()
    assert ppg.shape == g['ppg'].shape
    assert np.abs(ppg - g['ppg']).max() <= 2e-5
    logits = O.from_features(sd, feats, lengths, softmax=False, is_causal=causal).numpy()
    assert np.abs(logits - g['logits']).max() <= 2e-4
    # fp64 evaluation of the same restatement is the ground truth the <=1e-4
    # product tolerance is quoted against
    if int(g['frames']) <= 501:
        ppg64 = O.from_features(sd, feats, lengths, is_causal=causal, dtype=torch.float64)
        assert np.abs(ppg64.float().numpy() - g['ppg']).max() <= 2e-5


def test_as_shipped_context():
    """The reference as shipped (bf16 autocast on CPU) is ~1e-2 from fp32: the
    <=1e-4 target is only defined against the autocast-off modules."""
    g = golden('asshipped_from_audio_1x64000_s0')
    assert str(g['dtype']) == 'torch.bfloat16'
    sd = O.random_state_dict(0)
    ppg = O.from_audio(sd, O.synthetic_audio(1, 64000, 0)).numpy()
    err = np.abs(ppg - g['ppg']).max()
    assert 1e-4 < err < 5e-2


def test_chunk_plan_bookkeeping():
    """ppgs/model/transformer.py:49-64 lengths bookkeeping."""
    plan = O.chunk_plan(1234, torch.tensor([1234, 900, 380, 10]))
    assert [(a, b) for a, b, _ in plan] == [(0, 500), (400, 900), (800, 1284), (1200, 1284)]
    assert plan[0][2].tolist() == [500, 500, 430, 60]
    assert plan[1][2].tolist() == [500, 500, 0, 0]
    assert plan[3][2].tolist() == [84, 0, 0, 0]


def test_slaney_basis_vs_torchaudio():
    """librosa.filters.mel restatement vs torchaudio's independent Slaney bank."""
    torchaudio = pytest.importorskip('torchaudio')
    ref = torchaudio.functional.melscale_fbanks(
        513, 0.0, 8000.0, 80, 16000, norm='slaney', mel_scale='slaney').T.numpy()
    assert np.abs(O.mel_basis() - ref).max() < 2e-7


@pytest.mark.skipif(not os.path.isdir('/root/reference/ppgs'),
                    reason='reference tree only exists in the dev container')
def test_oracle_vs_live_reference():
    """Live pin: the reference's own mel + Transformer modules, imported under the
    stub shim, on a fresh seed that is NOT among the golden fixtures."""
    from oracle import refshim
    ppgs = refshim.import_reference()
    audio = O.synthetic_audio(2, 160 * 530, 11)
    lengths = torch.tensor([530, 301])
    mel = ppgs.preprocess.mel.from_audios(audio, lengths * 160)
    assert torch.equal(mel, O.mel_from_audios(audio))
    sd = O.random_state_dict(7, peaky=True)
    model = ppgs.model.Transformer()
    model.load_state_dict(sd)
    model.eval()
    with torch.inference_mode():
        ref = torch.softmax(model(mel.float(), lengths), dim=1)
    mine = O.from_features(sd, mel, lengths)
    assert (ref - mine).abs().max() <= 2e-5
