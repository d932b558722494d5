"""Host-side handle of one packed model on one GPU (ppgs.load.model +
ppgs.Model + the module the reference caches in ppgs/core.py:565-580)."""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from . import config


def mel_basis(sr=config.SAMPLE_RATE, n_fft=config.NUM_FFT, n_mels=config.NUM_MELS):
    """Slaney filterbank, the table `librosa.filters.mel(sr=16000, n_fft=1024,
    n_mels=80)` returns at ppgs/preprocess/mel.py:61-64 (built once, not per
    call — SURVEY.md F12).  Host-side table construction; uploaded to the GPU."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz,
                        min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep,
                        f / f_sp)

    def to_hz(m):
        return np.where(m >= min_log_mel,
                        min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fft_f = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = to_hz(np.linspace(to_mel(0.0), to_mel(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fft_f)
    weights = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float32)
    for i in range(n_mels):
        weights[i] = np.maximum(
            0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One ppgs_engine: packed weights + workspace on `device`."""

    def __init__(self, device, input_channels=config.live('INPUT_CHANNELS'),
                 hidden_channels=config.live('HIDDEN_CHANNELS'),
                 num_hidden_layers=config.live('NUM_HIDDEN_LAYERS'),
                 output_channels=config.live('OUTPUT_CHANNELS'),
                 kernel_size=config.live('KERNEL_SIZE'),
                 attention_heads=config.live('ATTENTION_HEADS'),
                 is_causal=config.live('IS_CAUSAL'), max_len=config.live('MAX_LEN')):
        # defaults come from the configuration as it is NOW (configure() / --config)
        input_channels, hidden_channels = config.resolve(input_channels), config.resolve(hidden_channels)
        num_hidden_layers, output_channels = config.resolve(num_hidden_layers), config.resolve(output_channels)
        kernel_size, attention_heads = config.resolve(kernel_size), config.resolve(attention_heads)
        is_causal, max_len = config.resolve(is_causal), config.resolve(max_len)
        self.device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('ppgs_b200 runs on CUDA devices only (no CPU fallback)')
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', index)
        cfg = _lib.ModelConfig()
        _lib.lib.ppgs_default_config(ctypes.byref(cfg))
        cfg.input_channels = input_channels
        cfg.hidden_channels = hidden_channels
        cfg.num_layers = num_hidden_layers
        cfg.num_heads = attention_heads
        cfg.output_channels = output_channels
        cfg.kernel_size = kernel_size
        cfg.is_causal = int(bool(is_causal))
        cfg.max_len = max_len
        cfg.chunk_length = config.CHUNK_LENGTH
        cfg.chunk_overlap = config.CHUNK_OVERLAP
        cfg.ffn_channels = config.FFN_CHANNELS
        cfg.layer_norm_eps = config.LAYER_NORM_EPS
        self.cfg = cfg
        handle = ctypes.c_void_p()
        torch.cuda.init()
        _lib.check(_lib.lib.ppgs_engine_create(ctypes.byref(cfg), index, ctypes.byref(handle)))
        self._handle = handle
        self._set('frontend.window', torch.hann_window(config.WINDOW_SIZE, dtype=torch.float32))
        self._set('frontend.mel_basis', torch.from_numpy(mel_basis()))

    def __del__(self):
        handle = getattr(self, '_handle', None)
        lib = getattr(_lib, 'lib', None) if _lib is not None else None   # None at interpreter exit
        if handle is not None and handle.value and lib is not None:
            lib.ppgs_engine_destroy(handle)
            self._handle = None

    # -- weights ----------------------------------------------------------
    def _set(self, name, tensor):
        tensor = tensor.detach().to('cpu', torch.float32).contiguous()
        shape = (ctypes.c_int64 * tensor.dim())(*tensor.shape)
        _lib.check(_lib.lib.ppgs_engine_set_weight(
            self._handle, name.encode(), ctypes.c_void_p(tensor.data_ptr()), shape, tensor.dim()))

    def load_state_dict(self, state_dict):
        """Strict load of the reference schema (ppgs/load.py:76-79)."""
        if 'model' in state_dict and not torch.is_tensor(state_dict['model']):
            state_dict = state_dict['model']
        for name, tensor in state_dict.items():
            self._set(name, tensor)
        _lib.check(_lib.lib.ppgs_engine_finalize(self._handle))
        return self

    def load_w2v2_state_dict(self, state_dict):
        """Strict load of a Hugging Face `Wav2Vec2Model` ('facebook/wav2vec2-base'
        architecture) state dict: the front-end of the w2v2fb representation
        (ppgs/preprocess/w2v2fb/core.py:44-47)."""
        for name, tensor in state_dict.items():
            self._set('w2v2.' + name, tensor)
        _lib.check(_lib.lib.ppgs_w2v2_finalize(self._handle))
        self.has_w2v2 = True
        return self

    def blob(self):
        """The packed weight blob as a uint8 CUDA tensor view (for the one-off
        torch.distributed broadcast from rank 0)."""
        nbytes = _lib.lib.ppgs_engine_blob_bytes(self._handle)
        ptr = _lib.lib.ppgs_engine_blob_dev(self._handle)
        if not ptr:
            _lib.check(_lib.E_CUDA)
        return _tensor_from_ptr(ptr, nbytes, self.device)

    def adopt_blob(self):
        _lib.check(_lib.lib.ppgs_engine_adopt_blob(self._handle))
        return self

    # -- knobs --------------------------------------------------------------
    @property
    def precision(self):
        code = _lib.lib.ppgs_engine_get_precision(self._handle)
        return {v: k for k, v in _lib.PRECISIONS.items()}[code]

    @precision.setter
    def precision(self, name):
        _lib.check(_lib.lib.ppgs_engine_set_precision(self._handle, _lib.PRECISIONS[name]))

    @property
    def launches(self):
        return _lib.lib.ppgs_engine_launch_count(self._handle)

    @property
    def graph_replays(self):
        """Forwards replayed as one CUDA graph (steady-state loops of from_audio*)."""
        return _lib.lib.ppgs_engine_graph_replays(self._handle)

    def set_graphs(self, enabled):
        _lib.check(_lib.lib.ppgs_engine_set_graphs(self._handle, int(bool(enabled))))

    def set_profiling(self, enabled):
        """Per-kernel CUDA-event timing (bench.py roofline); clears the stats."""
        _lib.check(_lib.lib.ppgs_engine_set_profiling(self._handle, int(bool(enabled))))

    def kernel_stats(self):
        """{kernel name: (total ms, launches)} since profiling was enabled."""
        stats, index = {}, 0
        name = ctypes.create_string_buffer(128)
        ms, launches = ctypes.c_double(), ctypes.c_int64()
        while _lib.lib.ppgs_engine_kernel_stat(
                self._handle, index, name, 128, ctypes.byref(ms), ctypes.byref(launches)) == 0:
            stats[name.value.decode()] = (ms.value, launches.value)
            index += 1
        return stats

    # -- forward ------------------------------------------------------------
    def mel(self, audio):
        """audio (B,1,samples) or (B,samples) fp32 CUDA -> (B,80,frames) fp16."""
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        audio = self._on_device(audio, torch.float32)
        if audio.stride(-1) != 1:
            audio = audio.contiguous()
        batch, samples = audio.shape
        out = torch.empty(batch, config.NUM_MELS, samples // config.HOPSIZE,
                          dtype=torch.float16, device=self.device)
        stride = audio.stride(0) if batch > 1 else samples
        _lib.check(_lib.lib.ppgs_mel_forward(
            self._handle, ctypes.c_void_p(audio.data_ptr()), batch, samples, stride,
            ctypes.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def w2v2fb(self, audio, lengths=None):
        """audio (B,1,samples) fp32 (zero padded), lengths in samples -> wav2vec2-base
        latents upsampled to the frame rate, (B,768,samples//160) fp16."""
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        audio = self._on_device(audio, torch.float32).contiguous()
        batch, samples = audio.shape
        out = torch.empty(batch, 768, samples // config.HOPSIZE, dtype=torch.float16,
                          device=self.device)
        lengths_host = None if lengths is None else _host_lengths(lengths, batch)
        _lib.check(_lib.lib.ppgs_w2v2fb_forward(
            self._handle, ctypes.c_void_p(audio.data_ptr()), batch, samples, samples, lengths_host,
            ctypes.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def transformer(self, features, lengths, softmax=True, legacy_mode=False):
        """features (B,C,T) fp16 CUDA, lengths (B,) ints -> (B,40,T) fp32."""
        features = self._on_device(features, torch.float16).contiguous()
        batch, _, frames = features.shape
        lengths_host = _host_lengths(lengths, batch)
        out = torch.empty(batch, self.cfg.output_channels, frames,
                          dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.ppgs_transformer_forward(
            self._handle, ctypes.c_void_p(features.data_ptr()), batch, frames, lengths_host,
            int(bool(softmax)), int(bool(legacy_mode)), ctypes.c_void_p(out.data_ptr()),
            _stream_ptr(self.device)))
        return out

    def from_audio(self, audio, lengths=None, softmax=True, legacy_mode=False, out=None):
        """Fused mel + transformer on device buffers. audio (B,1,samples) CUDA;
        lengths in SAMPLES (None = full).  `out`: optional preallocated (B,40,frames) fp32
        CUDA tensor (steady-state loops: no allocation per call)."""
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        audio = self._on_device(audio, torch.float32).contiguous()
        batch, samples = audio.shape
        shape = (batch, self.cfg.output_channels, samples // config.HOPSIZE)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif (tuple(out.shape) != shape or out.dtype != torch.float32 or out.device != self.device
              or not out.is_contiguous()):
            raise ValueError(f'out must be a contiguous fp32 tensor of shape {shape} on {self.device}')
        lengths_host = None if lengths is None else _host_lengths(lengths, batch)
        _lib.check(_lib.lib.ppgs_from_audio(
            self._handle, ctypes.c_void_p(audio.data_ptr()), batch, samples, samples,
            lengths_host, int(bool(softmax)), int(bool(legacy_mode)),
            ctypes.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def from_audio_host(self, audio, out=None, lengths=None, softmax=True, legacy_mode=False,
                        wait=True):
        """Host buffers in, host buffers out (H2D + compute + D2H).  audio: CPU fp32
        (B,1,samples), pinned for full PCIe speed.  `wait=False` only enqueues the
        request (two may be in flight); call `wait()` before reading `out`."""
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        if audio.device.type != 'cpu' or audio.dtype != torch.float32 or not audio.is_contiguous():
            raise ValueError('from_audio_host expects a contiguous fp32 CPU tensor')
        batch, samples = audio.shape
        frames = samples // config.HOPSIZE
        if out is None:
            out = torch.empty(batch, self.cfg.output_channels, frames, dtype=torch.float32,
                              pin_memory=True)
        lengths_host = None if lengths is None else _host_lengths(lengths, batch)
        call = _lib.lib.ppgs_from_audio_host if wait else _lib.lib.ppgs_from_audio_host_submit
        _lib.check(call(
            self._handle, ctypes.c_void_p(audio.data_ptr()), batch, samples, lengths_host,
            int(bool(softmax)), int(bool(legacy_mode)), ctypes.c_void_p(out.data_ptr()),
            _stream_ptr(self.device)))
        if not wait:   # the library reads / writes these buffers until wait()
            self._in_flight = (getattr(self, '_in_flight', None) or [])[-1:] + [(audio, out)]
        return out

    # -- audio ingest / file pipeline ---------------------------------------
    def resample(self, audio, sample_rate, target_rate=config.SAMPLE_RATE):
        """torchaudio.transforms.Resample(sample_rate, target_rate) on the GPU
        (ppgs/core.py:599-608): audio (..., samples) -> (..., ceil(samples * target / orig))
        fp32 CUDA."""
        sample_rate, target_rate = _integer_rate(sample_rate), _integer_rate(target_rate)
        audio = self._on_device(audio, torch.float32)
        shape = audio.shape
        flat = audio.reshape(-1, shape[-1]).contiguous()
        batch, samples = flat.shape
        out_len = _lib.lib.ppgs_resample_length(samples, sample_rate, target_rate)
        out = torch.empty(batch, out_len, dtype=torch.float32, device=self.device)
        if batch and samples:
            _lib.check(_lib.lib.ppgs_resample(
                self._handle, ctypes.c_void_p(flat.data_ptr()), batch, samples, samples,
                sample_rate, target_rate, ctypes.c_void_p(out.data_ptr()), out_len,
                _stream_ptr(self.device)))
        return out.reshape(shape[:-1] + (out_len,))

    def pcm16_to_f32(self, pcm):
        """int16 PCM CUDA tensor -> fp32 / 32768 (torchaudio.load's normalisation)."""
        pcm = self._on_device(pcm, torch.int16).contiguous()
        out = torch.empty(pcm.shape, dtype=torch.float32, device=self.device)
        if pcm.numel():
            _lib.check(_lib.lib.ppgs_pcm16_to_f32(
                self._handle, ctypes.c_void_p(pcm.data_ptr()), pcm.numel(),
                ctypes.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def files_to_files(self, batches, output_files, samples, reader_threads=8, writer_threads=8,
                       legacy_mode=False):
        """The native file pipeline (ppgs_files_to_files): `batches` = list of lists of
        audio file names (16-bit PCM, 16 kHz WAVE), `output_files` = {audio file: output
        file}, `samples` = {audio file: sample count from the header}.  Returns the number
        of posteriorgram frames written."""
        flat = [file for batch in batches for file in batch]
        if not flat:
            return 0
        sizes = (ctypes.c_int32 * len(batches))(*[len(batch) for batch in batches])
        audio = (ctypes.c_char_p * len(flat))(*[os.fsencode(file) for file in flat])
        outputs = (ctypes.c_char_p * len(flat))(*[os.fsencode(output_files[file]) for file in flat])
        counts = (ctypes.c_int64 * len(flat))(*[int(samples[file]) for file in flat])
        frames = ctypes.c_int64()
        _lib.check(_lib.lib.ppgs_files_to_files(
            self._handle, len(batches), sizes, audio, outputs, counts, int(reader_threads),
            int(writer_threads), int(bool(legacy_mode)), _stream_ptr(self.device),
            ctypes.byref(frames)))
        return frames.value

    def check(self):
        """Synchronise and raise if a kernel reported a pipeline time-out (the device
        entry points are asynchronous and cannot report it themselves)."""
        _lib.check(_lib.lib.ppgs_engine_wait(self._handle))

    def wait(self):
        """Block until every request enqueued with `wait=False` has landed."""
        _lib.check(_lib.lib.ppgs_engine_wait(self._handle))
        self._in_flight = None

    def _on_device(self, tensor, dtype):
        if tensor.device != self.device or tensor.dtype != dtype:
            tensor = tensor.to(self.device, dtype)
        return tensor


def _integer_rate(rate):
    """torchaudio refuses non-integer rates (functional._get_sinc_resample_kernel)."""
    if int(rate) != rate:
        raise ValueError(
            'Frequencies must be of integer type to ensure quality resampling computation.')
    return int(rate)


def _host_lengths(lengths, batch):
    if torch.is_tensor(lengths):
        lengths = lengths.detach().cpu().reshape(-1).tolist()
    elif isinstance(lengths, (int, np.integer)):
        lengths = [int(lengths)] * batch
    lengths = [int(x) for x in lengths]
    if len(lengths) != batch:
        # the reference's key-padding mask assertion
        raise ValueError(
            f'Expected lengths to have {batch} entries, but got {len(lengths)}')
    return (ctypes.c_int64 * batch)(*lengths)


def _tensor_from_ptr(ptr, nbytes, device):
    """Zero-copy uint8 CUDA tensor over engine-owned device memory."""

    class _Holder:
        pass

    holder = _Holder()
    holder.__cuda_array_interface__ = {
        'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}
    return torch.as_tensor(holder, device=device)
