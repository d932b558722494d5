"""w2v2fb representation (ppgs/preprocess/w2v2fb/core.py:32-94): wav2vec2-base latents at
the PPG frame rate, computed by the CUDA library (split-fp16 tcgen05 GEMMs / attention,
fp32 GroupNorm / LayerNorm; PPGS_B200_W2V2_TC=0 keeps everything in fp32 on the CUDA cores)."""
import os

import torch

from .. import config
from .. import load

_engines = {}


def state_dict():
    """The wav2vec2-base weights (ppgs/preprocess/w2v2fb/core.py:44-47 downloads them with
    `Wav2Vec2Model.from_pretrained`).  Resolution order: `config.W2V2FB_CHECKPOINT` /
    $PPGS_B200_W2V2FB (a torch state-dict file or a Hugging Face model directory), then the
    local Hugging Face cache."""
    source = config.W2V2FB_CHECKPOINT or os.environ.get('PPGS_B200_W2V2FB')
    if source is not None and os.path.isfile(source):
        state = torch.load(source, map_location='cpu')
        return state.get('model', state)
    try:
        import transformers
        model = transformers.Wav2Vec2Model.from_pretrained(
            source or config.W2V2FB_CONFIG, local_files_only=source is None)
    except Exception as error:
        raise RuntimeError(
            'w2v2fb needs the facebook/wav2vec2-base weights: set ppgs_b200.config.'
            f'W2V2FB_CHECKPOINT or $PPGS_B200_W2V2FB to a state-dict file or model directory ({error})'
        ) from error
    return model.state_dict()


def engine(gpu=None, weights=None):
    """Front-end engine of a device (cached); `weights` overrides the resolved state dict."""
    device = load.resolve_device(gpu)
    cached = _engines.get(device.index)
    if cached is None or weights is not None:
        from ..engine import Engine
        from . import mel
        cached = mel._standalone(device) if cached is None else cached
        cached.load_w2v2_state_dict(weights if weights is not None else state_dict())
        _engines[device.index] = cached
    return cached


def from_audios(audio, lengths=None, sample_rate=config.SAMPLE_RATE, gpu=None):
    """(B,1,samples) zero-padded fp32 audio, lengths (B,) samples -> (B,768,samples//160) fp16."""
    from ..core import resample
    audio = resample(audio, sample_rate)
    return engine(gpu).w2v2fb(audio, lengths)


def from_audio(audio, sample_rate=config.SAMPLE_RATE, gpu=None):
    if audio.dim() == 2:
        audio = audio.unsqueeze(dim=0)
    return from_audios(audio, None, sample_rate, gpu)
