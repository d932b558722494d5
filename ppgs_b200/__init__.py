"""ppgs_b200 — B200-native drop-in for the `ppgs.from_audio` forward path of
interactiveaudiolab/ppgs (mel front-end -> Transformer encoder -> 40-way phoneme
posteriors).  Same API names and defaults as the reference (ppgs/__init__.py,
ppgs/core.py); the arithmetic is hand-written sm_100a CUDA behind the C ABI in
include/ppgs_b200.h.  Importing the package loads libppgs_b200.so and raises if it
has not been built — there is no CPU / PyTorch fallback."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .config import *  # noqa: F401,F403
from .config import configure  # noqa: F401
from .phonemes import PHONEMES, PHONEME_TO_INDEX_MAPPING, SILENCE  # noqa: F401
from . import config, data, load, preprocess, parallel  # noqa: F401
from .engine import Engine  # noqa: F401
from .streaming import LongStreamer, Streamer  # noqa: F401
from .core import (  # noqa: F401
    from_audio, from_features, from_file, from_file_to_file, from_files_to_files,
    from_feature_files_to_files, from_dataloader, infer, resample, representation_file_extension,
    distance, interpolate, sparsify)
from . import edit  # noqa: F401

__version__ = '0.1.0'
