"""Import the UNMODIFIED reference (`/root/reference/ppgs`, or its travelled copy
`oracle/_ref/ppgs` made by `oracle/build_ref.py`).

TEST / BENCH INFRASTRUCTURE ONLY.  Used by `oracle/make_golden*.py` (and by
`tests/test_oracle_golden.py::test_oracle_vs_live_reference`) to pin the oracle
restatement against the reference's own modules, and by `oracle/ref_arm.py` for
`bench.py --impl reference` / `cpu_baseline` / `torch_gpu_baseline` (the reference's own
code timed beside the product).  Never imported by the product package, by `-m gpu`
tests or by `smoke()`.

The reference cannot be imported as-is: `yapecs`, `torchutil`,
`pypar`, `librosa`, `matplotlib`, `moviepy`, `espnet`, ... are not installed and
there is no network.  This module registers minimal stand-ins in `sys.modules`
*before* `import ppgs`:

* `yapecs.configure` -> no-op (reference `ppgs/__init__.py:10-11`)
* `torchutil.inference.context(model)` -> eval + inference_mode + autocast of
  the model's device type, then train() (call site `ppgs/core.py:586`);
  a restatement from the public torchutil behaviour — UNPINNED (package absent).
* `librosa.filters.mel` -> the oracle's restatement of librosa's Slaney
  filterbank (call site `ppgs/preprocess/mel.py:61-64`); torchaudio's
  independent implementation agrees to 8e-8.
* everything else the reference merely imports -> auto-stub modules.
"""
import contextlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# The travelled copy made by `oracle/build_ref.py` (byte-identical *.py files of the reference
# package; git-ignored, ships to the GPU box) is used when the reference tree itself is absent.
TRAVELLED_ROOT = os.path.join(_HERE, '_ref')


def _pick_root():
    env = os.environ.get('PPGS_REFERENCE_ROOT')
    for root in ([env] if env else []) + ['/root/reference', TRAVELLED_ROOT]:
        if os.path.isdir(os.path.join(root, 'ppgs')):
            return root
    return env or '/root/reference'


REFERENCE_ROOT = _pick_root()

_AUTO_STUB_ROOTS = (
    'espnet', 'torch_complex', 'nltk', 'gdown', 'humanfriendly', 'dac',
    'encodec', 'g2pM', 'matplotlib', 'mpl_toolkits', 'moviepy', 'cv2',
    'apprise', 'pyfoal', 'pysodic', 'penn', 'promonet', 'tensorboard',
    'accelerate', 'soundfile', 'torbi', 'opencv', 'PIL_stub')


class _Anything:
    """Attribute sink: any attribute / call / subclassing works."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        value = _Anything
        setattr(self, name, value)
        return value


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        module = _StubModule(spec.name)
        module.__path__ = []
        return module

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _AUTO_STUB_ROOTS:
            return importlib.machinery.ModuleSpec(
                fullname, _StubLoader(), is_package=True)
        return None


def _module(name, **attrs):
    module = types.ModuleType(name)
    module.__dict__.update(attrs)
    sys.modules[name] = module
    return module


def slaney_mel_basis(sr=16000, n_fft=1024, n_mels=80, **_):
    """Stand-in for `librosa.filters.mel(sr, n_fft, n_mels)` (htk=False,
    norm='slaney', fmin=0, fmax=sr/2): the oracle's restatement of librosa's
    published construction (librosa itself is absent).  An independent
    implementation (torchaudio.functional.melscale_fbanks) agrees to 8e-8 —
    checked in tests/test_oracle.py."""
    from oracle import ppg_oracle
    return ppg_oracle.mel_basis(sr, n_fft, n_mels)


def install():
    """Register the stand-ins and put the reference on sys.path. Idempotent."""
    if getattr(install, 'done', False):
        return
    import torch
    # transformers first: a `librosa` stub visible earlier breaks its soxr probe
    from transformers import Wav2Vec2Model, Wav2Vec2Config  # noqa: F401
    import argparse

    _module('yapecs', configure=lambda *a, **k: None,
            ArgumentParser=argparse.ArgumentParser)

    @contextlib.contextmanager
    def context(model, autocast=True):
        device_type = next(model.parameters()).device.type
        model.eval()
        with torch.inference_mode(), torch.autocast(device_type, enabled=autocast):
            yield
        model.train()

    def iterator(iterable, message=None, initial=0, total=None):
        return _Anything()

    def notify(*a, **k):
        def decorator(fn):
            return fn
        return decorator

    tu = _module('torchutil', notify=notify, iterator=iterator)
    tu.inference = _module('torchutil.inference', context=context)
    for sub in ('checkpoint', 'tensorboard', 'gradients', 'cuda', 'download',
                'time', 'metrics'):
        stub = _StubModule(f'torchutil.{sub}')
        sys.modules[f'torchutil.{sub}'] = stub
        setattr(tu, sub, stub)
    tu.metrics.Accuracy = type('Accuracy', (), {})
    tu.metrics.Average = type('Average', (), {})

    pypar = _module('pypar', SILENCE='<silent>')
    for name in ('Alignment', 'Word', 'Phoneme'):
        setattr(pypar, name, type(name, (), {'__init__': lambda s, *a, **k: None}))

    librosa = _module('librosa')
    librosa.filters = _module('librosa.filters', mel=slaney_mel_basis)

    sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install.done = True


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'ppgs'))


def import_reference():
    """Return the reference `ppgs` package (imported unchanged)."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    install()
    import ppgs
    return ppgs
