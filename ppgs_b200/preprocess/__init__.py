"""`representation=` plugin surface (ppgs/preprocess/__init__.py,
ppgs/preprocess/core.py:194-221): modules exposing `from_audio(audio,
sample_rate=, gpu=)` and `from_audios(audio, lengths, gpu=)`."""
import torch

from .. import config
from . import mel
from . import w2v2fb

REGISTRY = {'mel': mel, 'w2v2fb': w2v2fb}


def get(representation):
    """getattr(ppgs.preprocess, representation) of ppgs/core.py:333-339, as an
    explicit registry; unknown names raise like ppgs/load.py:46-48."""
    if representation is None:
        representation = config.REPRESENTATION
    if representation not in REGISTRY:
        raise ValueError(
            f'Representation {representation} is not supported by ppgs_b200 '
            f'(available: {sorted(REGISTRY)})')
    return REGISTRY[representation]


def from_audio(audio, representation=config.REPRESENTATION,
               sample_rate=config.SAMPLE_RATE, gpu=None):
    """Preprocess audio (ppgs/preprocess/core.py:194-216)."""
    from ..core import resample
    audio = resample(audio, sample_rate)
    features = get(representation).from_audio(
        audio, sample_rate=config.SAMPLE_RATE, gpu=gpu)
    if features.dim() == 2:
        features = features[None]
    return features


def save_masked(tensor, file, length):
    """Save masked tensor (ppgs/preprocess/core.py:219-221)."""
    torch.save(tensor[..., :length].clone(), file)
