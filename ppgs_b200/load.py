"""Loading utilities (ppgs/load.py:17-81)."""
import ctypes
import os
import threading
from pathlib import Path

import numpy as np
import torch

from . import _lib
from . import config
from .engine import Engine

_engines = {}
_lock = threading.Lock()


def wav_info(file):
    """Header of a RIFF/WAVE file through the native probe (ppgs_wav_info — the
    torchaudio.info call of ppgs/data/dataset.py:187): dict(samples, sample_rate,
    channels, bits, is_float), or None when the file is not a WAVE file the
    library decodes."""
    frames, rate = ctypes.c_int64(), ctypes.c_int()
    channels, bits, is_float = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    code = _lib.lib.ppgs_wav_info(
        os.fsencode(str(file)), ctypes.byref(frames), ctypes.byref(rate), ctypes.byref(channels),
        ctypes.byref(bits), ctypes.byref(is_float))
    if code == _lib.E_UNSUPPORTED:
        return None
    _lib.check(code)
    return {'samples': frames.value, 'sample_rate': rate.value, 'channels': channels.value,
            'bits': bits.value, 'is_float': bool(is_float.value)}


def wav_info_many(files, threads=16):
    """`wav_info` of every file in one native call (header probes on `threads` host threads);
    entries are None for files the library does not decode."""
    count = len(files)
    if count == 0:
        return []
    paths = (ctypes.c_char_p * count)(*[os.fsencode(str(file)) for file in files])
    frames = (ctypes.c_int64 * count)()
    arrays = [(ctypes.c_int32 * count)() for _ in range(5)]   # rate, channels, bits, is_float, status
    _lib.check(_lib.lib.ppgs_wav_info_many(paths, count, int(threads), frames, *arrays))
    rate, channels, bits, is_float, status = arrays
    infos = []
    for i in range(count):
        if status[i] == _lib.E_UNSUPPORTED:
            infos.append(None)
        elif status[i] != 0:
            infos.append(wav_info(files[i]))     # raises with the file's own message
        else:
            infos.append({'samples': frames[i], 'sample_rate': rate[i], 'channels': channels[i],
                          'bits': bits[i], 'is_float': bool(is_float[i])})
    return infos


def flac_info(file):
    """STREAMINFO of a FLAC file (ppgs_flac_info): dict(samples, sample_rate, channels, bits), or
    None when the file is not a FLAC stream."""
    frames, rate, channels, bits = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    code = _lib.lib.ppgs_flac_info(
        os.fsencode(str(file)), ctypes.byref(frames), ctypes.byref(rate), ctypes.byref(channels),
        ctypes.byref(bits))
    if code == _lib.E_UNSUPPORTED:
        return None
    _lib.check(code)
    return {'samples': frames.value, 'sample_rate': rate.value, 'channels': channels.value,
            'bits': bits.value}


def audio(file, device=None):
    """Load audio from disk as (channels, samples) fp32 at 16 kHz
    (ppgs/load.py:17-30).  Mono PCM / float WAVE files and FLAC files are decoded by the
    native readers (ppgs_wav_read_f32 / ppgs_flac_read_f32; torchaudio.load needs
    torchcodec) and resampled on the GPU; multi-channel WAVE files by scipy;
    other containers go through torchaudio when it can decode them.  `device`: return the
    waveform on that CUDA device instead of the CPU (saves a round trip after resampling)."""
    path = Path(file)
    info = wav_info(path) if path.suffix.lower() == '.wav' else None
    flac = flac_info(path) if path.suffix.lower() == '.flac' else None
    if flac is not None:
        waveform = torch.empty(flac['channels'], flac['samples'], dtype=torch.float32)
        frames, rate, channels = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib.ppgs_flac_read_f32(
            os.fsencode(str(path)), ctypes.c_void_p(waveform.data_ptr()), flac['samples'],
            ctypes.byref(frames), ctypes.byref(rate), ctypes.byref(channels)))
        sample_rate = rate.value
    elif info is not None and info['channels'] == 1:
        waveform = torch.empty(1, info['samples'], dtype=torch.float32)
        frames, rate = ctypes.c_int64(), ctypes.c_int()
        _lib.check(_lib.lib.ppgs_wav_read_f32(
            os.fsencode(str(path)), ctypes.c_void_p(waveform.data_ptr()), waveform.numel(),
            ctypes.byref(frames), ctypes.byref(rate)))
        sample_rate = rate.value
    elif path.suffix.lower() == '.wav':
        from scipy.io import wavfile
        sample_rate, data = wavfile.read(path)
        if data.dtype == np.int16:
            data = data.astype(np.float32) / 32768.0
        elif data.dtype == np.int32:
            data = data.astype(np.float32) / 2147483648.0
        elif data.dtype == np.uint8:
            data = (data.astype(np.float32) - 128.0) / 128.0
        else:
            data = data.astype(np.float32)
        data = data[None] if data.ndim == 1 else data.T
        waveform = torch.from_numpy(np.ascontiguousarray(data))
    else:
        import torchaudio
        try:
            if path.suffix.lower() == '.mp3':
                waveform, sample_rate = torchaudio.load(path, format='mp3')
            else:
                waveform, sample_rate = torchaudio.load(file)
        except RuntimeError:
            if path.suffix.lower() == '.mp3':
                raise RuntimeError(
                    'Failed to load mp3 file, make sure ffmpeg<=4.3 is installed')
            raise
    from .core import resample
    if device is not None:
        return resample(waveform.to(device), sample_rate)
    return resample(waveform, sample_rate)


def wav_num_frames(file):
    """(samples, sample_rate) from the header only (torchaudio.info at
    ppgs/data/dataset.py:187)."""
    info = wav_info(file)
    if info is None and str(file).lower().endswith('.flac'):
        info = flac_info(file)
    if info is not None:
        return info['samples'], info['sample_rate']
    waveform = audio(file)
    return waveform.shape[-1], config.SAMPLE_RATE


def tensor_info(file):
    """(shape, dtype) of a `.pt` file holding one whole contiguous fp16 / fp32 CPU tensor, from
    the archive's directory and pickle header only (native: ppgs_pt_info); None when the native
    reader does not understand the file (callers then use torch.load)."""
    import ctypes
    ndim, elem = ctypes.c_int(), ctypes.c_int()
    dims = (ctypes.c_int64 * 3)()
    if _lib.lib.ppgs_pt_info(os.fsencode(str(file)), ctypes.byref(ndim), dims, ctypes.byref(elem)) != 0:
        return None
    shape = tuple(dims[3 - ndim.value:3])
    return shape, torch.float16 if elem.value == 2 else torch.float32


def features(file, out=None):
    """Load cached input features, `torch.load(cache / f'{stem}-{feature}.pt')` of
    ppgs/data/dataset.py:98-101, through the native reader (no unpickling in Python, no GIL
    while the payload is read).  `out`: optional (channels, >= frames) slice of a padded batch
    tensor the rows are read into directly.  Falls back to torch.load for anything the native
    reader does not cover (legacy archives, other dtypes, views)."""
    import ctypes
    info = tensor_info(file)
    if info is None or len(info[0]) < 2:
        tensor = torch.load(file, map_location='cpu')
        if out is not None:
            out[..., :tensor.shape[-1]].copy_(tensor)
            return out[..., :tensor.shape[-1]]
        return tensor
    shape, dtype = info
    rows, cols = 1, shape[-1]
    for n in shape[:-1]:
        rows *= n
    if out is None:
        out = torch.empty(shape, dtype=dtype)
        stride = cols
    else:
        if out.dtype != dtype or out.device.type != 'cpu' or out.stride(-1) != 1 or out.shape[-1] < cols \
                or out.numel() // out.shape[-1] != rows or out.dim() != 2:
            raise ValueError(f'features: `out` must be a CPU {dtype} (rows, >= {cols}) view, got {tuple(out.shape)}')
        stride = out.stride(0)
    _lib.check(_lib.lib.ppgs_pt_read(
        os.fsencode(str(file)), ctypes.c_void_p(out.data_ptr()), rows, cols, 2 if dtype == torch.float16 else 4,
        stride))
    return out[..., :cols] if out.shape[-1] != cols else out


def state_dict(checkpoint=None, representation=None):
    """Resolve + read a checkpoint (ppgs/load.py:59-79).  Unlike the reference,
    `checkpoint=` is honoured for w2v2fb too."""
    if checkpoint is None:
        checkpoint = config.LOCAL_CHECKPOINT
    if checkpoint is None:
        name = representation if representation is not None else config.REPRESENTATION
        if name not in config.HF_CHECKPOINTS:
            raise ValueError(
                f'No default checkpoints exist for representation {name}')
        try:
            import huggingface_hub
            checkpoint = huggingface_hub.hf_hub_download(
                config.HF_REPO, config.HF_CHECKPOINTS[name])
        except Exception as error:
            raise RuntimeError(
                f'No checkpoint given and {config.HF_CHECKPOINTS[name]} could not be '
                f'fetched from {config.HF_REPO}: {error}') from error
    state = torch.load(checkpoint, map_location='cpu')
    if 'model' in state:
        state = state['model']
    return state


def model_kwargs(representation):
    """ppgs/load.py:35-50."""
    if representation is None:
        return {}
    if representation not in config.MODEL_KWARGS:
        raise ValueError(
            'Supplying representation directly only supported for w2v2fb and mel')
    return dict(config.MODEL_KWARGS[representation])


def model(checkpoint=None, representation=None, gpu=None, is_causal=None):
    """Build (or fetch the cached) engine for (representation, checkpoint, gpu)
    — ppgs.load.model + the cache of ppgs/core.py:565-580."""
    device = resolve_device(gpu)
    causal = config.IS_CAUSAL if is_causal is None else bool(is_causal)
    key = cache_key(representation, checkpoint, device.index, causal)
    with _lock:
        engine = _engines.get(key)
        if engine is None:
            kwargs = model_kwargs(representation)
            state = state_dict(checkpoint, representation)
            engine = Engine(device, is_causal=causal, **kwargs)
            engine.load_state_dict(state)
            precision = os.environ.get('PPGS_B200_PRECISION')
            if precision:
                engine.precision = precision
            _engines[key] = engine
    return engine


def utility_engine(gpu=None):
    """A weight-less engine on `gpu` for the ingest kernels (resampler, PCM decode)
    when no model is involved; cached per device."""
    device = resolve_device(gpu)
    key = ('utility', '', device.index, False, ())
    with _lock:
        engine = _engines.get(key)
        if engine is None:
            engine = _engines[key] = Engine(device)
    return engine


def cache_key(representation, checkpoint, gpu, is_causal=None):
    causal = config.IS_CAUSAL if is_causal is None else bool(is_causal)
    # the live model configuration is part of the identity of a cached engine
    shape = (config.INPUT_CHANNELS, config.HIDDEN_CHANNELS, config.NUM_HIDDEN_LAYERS,
             config.ATTENTION_HEADS, config.KERNEL_SIZE, config.OUTPUT_CHANNELS)
    return (str(representation), str(checkpoint), resolve_device(gpu).index, causal, shape)


def clear_cache():
    with _lock:
        _engines.clear()


def resolve_device(gpu=None):
    """`gpu` is a CUDA ordinal as in the reference; None means the current CUDA
    device (the reference would run on the CPU — this engine has no CPU path)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            'ppgs_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if gpu is None:
        gpu = torch.cuda.current_device()
    return torch.device('cuda', int(gpu))
