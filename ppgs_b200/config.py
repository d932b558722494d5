"""Constants of the reference configuration (ppgs/config/defaults.py:20-32,
127-161, ppgs/config/static.py:22, ppgs/config/w2v2fb.py:7-10).  The reference
freezes these at import time through yapecs; here they are plain module
attributes re-exported from `ppgs_b200`."""

# Audio parameters (defaults.py:20-32)
HOPSIZE = 160
NUM_FFT = 1024
NUM_MELS = 80
SAMPLE_RATE = 16000
WINDOW_SIZE = 1024

# Representations (defaults.py:43-53)
ALL_REPRESENTATIONS = ['mel', 'w2v2fb']
BEST_REPRESENTATION = 'mel'
REPRESENTATION = BEST_REPRESENTATION
REPRESENTATION_KIND = 'ppg'

# Model parameters (defaults.py:127-161)
LOCAL_CHECKPOINT = None
ATTENTION_HEADS = 2
IS_CAUSAL = False
HIDDEN_CHANNELS = 256
INPUT_CHANNELS = 80
KERNEL_SIZE = 5
MODEL = 'transformer'
NUM_HIDDEN_LAYERS = 5
OUTPUT_CHANNELS = 40
CHUNK_OVERLAP = 50
CHUNK_LENGTH = 500
MAX_LEN = 5000                   # ppgs/model/transformer.py:24
FFN_CHANNELS = 2048              # torch.nn.TransformerEncoderLayer default
LAYER_NORM_EPS = 1e-5            # torch.nn.TransformerEncoderLayer default

# Data parameters (defaults.py:170,185,202; static.py:22)
BUCKETS = 1
RANDOM_SEED = 1234
MAX_INFERENCE_FRAMES = float('inf')
MAX_PREPROCESS_FRAMES = 10000    # defaults.py:188

# PPG distance (defaults.py:93,214): the reference ships assets/balanced_similarity.pt; set
# SIMILARITY_MATRIX_PATH to that file (or pass `similarity=` to ppgs_b200.distance)
SIMILARITY_MATRIX_PATH = None
SIMILARITY_EXPONENT = 1.2

# Per-representation model kwargs (ppgs/load.py:35-50, ppgs/config/w2v2fb.py:7-10)
MODEL_KWARGS = {
    'mel': {},
    'w2v2fb': {'hidden_channels': 512, 'input_channels': 768},
}

# Hugging Face checkpoint names (ppgs/load.py:59-67)
HF_REPO = 'CameronChurchwell/ppgs'
HF_CHECKPOINTS = {'mel': 'mel-800k.pt', 'w2v2fb': 'w2v2fb-425k.pt'}

# w2v2fb front-end (ppgs/preprocess/w2v2fb/core.py:17-24): Hugging Face model id the
# reference downloads; a local directory / state-dict file can be given instead
W2V2FB_CONFIG = 'facebook/wav2vec2-base'
W2V2FB_CHECKPOINT = None


class Live:
    """Default argument that is read from the LIVE configuration when the function is called,
    not when it was defined: `def f(representation=config.live('REPRESENTATION'))` +
    `representation = config.resolve(representation)` sees `configure()` / `--config`
    overrides made after import (the reference gets the same effect from yapecs rewriting
    ppgs.config.defaults before the package body runs)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f'config.{self.name}'


def live(name):
    return Live(name)


def resolve(value):
    import sys
    return getattr(sys.modules[__name__], value.name) if isinstance(value, Live) else value


def configure(source):
    """Apply a yapecs-style configuration (the reference reads `--config file.py` at import,
    ppgs/__init__.py:7-15): every UPPER_CASE name of a python file (or dict) overrides the
    attribute of the same name, e.g. IS_CAUSAL = True from config/causal_transformer.py.
    Call before engines are created."""
    import runpy
    import sys
    values = dict(source) if isinstance(source, dict) else runpy.run_path(str(source))
    package = sys.modules.get('ppgs_b200')
    module = sys.modules[__name__]
    applied = {}
    for name, value in values.items():
        if name.isupper() and not name.startswith('_'):
            setattr(module, name, value)
            if package is not None:
                setattr(package, name, value)
            applied[name] = value
    return applied
