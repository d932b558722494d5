"""Stateful streaming inference for causal models (`config/causal_transformer.py`:
IS_CAUSAL = True) — BASELINE config 4 with state.  The reference cannot stream (SURVEY.md
F8); `Streamer` is the incremental form of its un-chunked causal forward: every frame it
emits equals `ppgs.from_features(..., legacy_mode=True)` of the whole utterance, 4 frames
(40 ms) after the frame was pushed."""
import ctypes

import torch

from . import _lib
from . import config
from .engine import _stream_ptr


class Streamer:
    """`streams` utterances decoded in lockstep on one engine (C ABI: ppgs_stream_*)."""

    LOOKAHEAD = 4

    def __init__(self, engine, streams):
        self.engine = engine
        self.streams = int(streams)
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib.ppgs_stream_create(engine._handle, self.streams, ctypes.byref(handle)))
        self._handle = handle
        self._audio = None       # (streams, samples) fp32 tail for push_audio
        self._frames_in = 0      # mel frames already handed to push()
        self._samples_seen = 0

    def __del__(self):
        handle = getattr(self, '_handle', None)
        lib = getattr(_lib, 'lib', None) if _lib is not None else None
        if handle is not None and handle.value and lib is not None:
            lib.ppgs_stream_destroy(handle)
            self._handle = None

    capacity = property(lambda self: _lib.lib.ppgs_stream_capacity())
    length = property(lambda self: _lib.lib.ppgs_stream_length(self._handle))
    emitted = property(lambda self: _lib.lib.ppgs_stream_emitted(self._handle))

    def reset(self, streams=None):
        """Start over: every stream, or only the listed ones (the others keep their state)."""
        if streams is None:
            _lib.check(_lib.lib.ppgs_stream_reset(self._handle, _stream_ptr(self.engine.device)))
            self._audio, self._frames_in, self._samples_seen = None, 0, 0
            return
        flags = (ctypes.c_int32 * self.streams)(*[0] * self.streams)
        for index in streams:
            flags[index] = 1
        _lib.check(_lib.lib.ppgs_stream_reset_streams(
            self._handle, flags, _stream_ptr(self.engine.device)))

    def state(self):
        """(lengths, emitted) per stream."""
        lengths = (ctypes.c_int32 * self.streams)()
        emitted = (ctypes.c_int32 * self.streams)()
        _lib.check(_lib.lib.ppgs_stream_state(self._handle, lengths, emitted))
        return list(lengths), list(emitted)

    def push(self, features=None, final=False, softmax=True, lengths=None):
        """features (streams, channels, n) fp16 (or None / n = 0 with `final` to flush)
        -> posteriorgram frames that became final, (streams, 40, m) fp32 CUDA.

        Independent streams: `lengths` = frames each stream brings (the first lengths[b]
        columns of its row) and `final` may be a per-stream sequence; the result is then
        (out, produced) with produced[b] valid frames in row b."""
        device = self.engine.device
        if features is None:
            features = torch.empty(self.streams, self.engine.cfg.input_channels, 0,
                                   dtype=torch.float16, device=device)
        features = self.engine._on_device(features, torch.float16).contiguous()
        if features.dim() != 3 or features.shape[0] != self.streams:
            raise ValueError(f'expected features of shape ({self.streams}, channels, frames)')
        if features.shape[1] != self.engine.cfg.input_channels:
            raise ValueError(f'expected {self.engine.cfg.input_channels} feature channels')
        frames = features.shape[-1]
        capacity = frames + self.LOOKAHEAD
        out = torch.empty(self.streams, self.engine.cfg.output_channels, capacity,
                          dtype=torch.float32, device=device)
        ragged = lengths is not None or not isinstance(final, (bool, int))
        if not ragged:
            produced = ctypes.c_int()
            _lib.check(_lib.lib.ppgs_stream_push(
                self._handle, ctypes.c_void_p(features.data_ptr()), frames, int(bool(final)),
                int(bool(softmax)), ctypes.c_void_p(out.data_ptr()), capacity, ctypes.byref(produced),
                _stream_ptr(device)))
            return out[..., :produced.value]
        if lengths is None:
            lengths = [frames] * self.streams
        if torch.is_tensor(lengths):
            lengths = lengths.tolist()
        finals = [bool(final)] * self.streams if isinstance(final, (bool, int)) else list(final)
        if len(lengths) != self.streams or len(finals) != self.streams:
            raise ValueError(f'expected {self.streams} lengths / final flags')
        counts = (ctypes.c_int32 * self.streams)(*[int(n) for n in lengths])
        flags = (ctypes.c_int32 * self.streams)(*[int(bool(f)) for f in finals])
        produced = (ctypes.c_int32 * self.streams)()
        _lib.check(_lib.lib.ppgs_stream_push_ragged(
            self._handle, ctypes.c_void_p(features.data_ptr()), frames, counts, flags,
            int(bool(softmax)), ctypes.c_void_p(out.data_ptr()), capacity, produced,
            _stream_ptr(device)))
        produced = list(produced)
        return out[..., :max(produced) if produced else 0], produced

    def push_audio(self, audio, final=False, softmax=True):
        """16 kHz audio (streams, 1, n) appended to the session.  Mel frame t reads samples
        [160 t - 432, 160 t + 592) of the reflect-padded utterance (ppgs/preprocess/
        spectrogram.py:27-43), so a frame is computed once its window is complete (or at
        `final`, with the reference's reflection at the end); the front-end runs on the
        buffered tail only."""
        hop, left = config.HOPSIZE, (config.NUM_FFT - config.HOPSIZE) // 2   # 160, 432
        device = self.engine.device
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        audio = self.engine._on_device(audio, torch.float32)
        if audio.shape[0] != self.streams:
            raise ValueError(f'expected audio of shape ({self.streams}, 1, samples)')
        first = self._frames_in
        # the tail starts 3 hops before the next frame: its first 3 frames absorb the
        # artificial reflection at the cut (3 * 160 >= 432), except at the true start
        skip = min(first, 3)
        start = (first - skip) * hop
        if self._audio is None:
            self._audio, self._offset = audio, 0
        else:
            self._audio = torch.cat((self._audio, audio), dim=-1)
        self._samples_seen += audio.shape[-1]
        self._audio = self._audio[:, start - self._offset:]
        self._offset = start
        total = self._samples_seen
        if final:
            last = total // hop                                   # frames = samples // 160
        else:
            last = max((total - (config.NUM_FFT - left)) // hop + 1, first)   # 160 t + 592 <= total
            last = min(last, total // hop)
        if last > first and self._audio.shape[-1] > left:
            mel = self.engine.mel(self._audio.contiguous())
            features = mel[..., skip:skip + (last - first)]
        else:
            last = first
            features = torch.empty(self.streams, config.NUM_MELS, 0, dtype=torch.float16,
                                   device=device)
        self._frames_in = last
        return self.push(features, final=final, softmax=softmax)


class LongStreamer:
    """Unbounded streaming for causal models: the reference's CHUNKED inference
    (ppgs/model/transformer.py:49-64: 500-frame chunks every 400 frames over the input
    left-padded with 50 replicas of its first frame, the middle 400 frames of every chunk
    kept), computed incrementally.  Chunk i is a `Streamer` session over padded frames
    [400 i, 400 i + 454): by causality its kept outputs (local frames 50..449) never see a
    later frame, so the session stops there and its state is recycled for chunk i + 2; two
    sessions alternate, and the 54 frames they share are pushed to both.  Every emitted
    frame equals `ppgs.from_features(whole_utterance)` (legacy_mode=False) of the causal
    model for utterances longer than one chunk (the reference only chunks when
    frames > 500; shorter utterances are what `Streamer` computes)."""

    def __init__(self, engine, streams):
        self.engine = engine
        self.streams = int(streams)
        self.chunk = config.CHUNK_LENGTH
        self.overlap = config.CHUNK_OVERLAP
        self.stride = self.chunk - 2 * self.overlap
        # a session consumes the kept frames + the look-ahead of the two convolutions
        self.consumed = self.chunk - self.overlap + Streamer.LOOKAHEAD
        self.sessions = [Streamer(engine, streams), Streamer(engine, streams)]
        self.reset()

    def reset(self):
        for session in self.sessions:
            session.reset()
        self.frames = 0          # feature frames pushed so far (un-padded axis)
        self.emitted = 0
        self.first_frame = None  # replicated 50 times in front of the stream
        self.closed = False

    def _session_input(self, index, features, begin, end):
        """Slice of padded frames [begin, end) that chunk `index` still has to see."""
        lo = max(begin, index * self.stride)
        hi = min(end, index * self.stride + self.consumed)
        return lo, hi

    def push(self, features=None, final=False, softmax=True):
        """features (streams, channels, n) fp16 -> (streams, 40, m) fp32: the frames that
        became final, in order."""
        if self.closed:
            raise RuntimeError('the stream was finalised; call reset()')
        device = self.engine.device
        channels = self.engine.cfg.input_channels
        if features is None:
            features = torch.empty(self.streams, channels, 0, dtype=torch.float16, device=device)
        features = self.engine._on_device(features, torch.float16)
        if features.dim() != 3 or features.shape[0] != self.streams or features.shape[1] != channels:
            raise ValueError(f'expected features of shape ({self.streams}, {channels}, frames)')
        n = features.shape[-1]
        if self.first_frame is None and n:
            self.first_frame = features[..., :1]
            features = torch.cat((self.first_frame.expand(-1, -1, self.overlap), features), dim=-1)
            begin = 0                               # padded axis = un-padded + 50
        else:
            begin = self.frames + self.overlap if self.first_frame is not None else 0
        end = begin + features.shape[-1]
        self.frames += n
        if final:
            self.closed = True
        pieces = []
        # chunks that overlap [begin, end) on the padded axis, oldest first
        first = max((begin - self.consumed) // self.stride + 1, 0) if begin >= self.consumed else 0
        last = max((end - 1) // self.stride, 0) if end > 0 else 0
        if final and self.frames:
            # the reference runs ceil(frames / 400) chunks
            last = min(last, -(-self.frames // self.stride) - 1)
        for index in range(first, last + 1):
            session = self.sessions[index % 2]
            lo, hi = self._session_input(index, features, begin, end)
            start = index * self.stride
            if lo == start and hi > lo and session.length:   # recycle the state of chunk index - 2
                session.reset()
            chunk_final = final and index * self.stride + self.consumed > end
            piece = features[..., lo - begin:hi - begin] if hi > lo else None
            if piece is None and not chunk_final:
                continue
            local_before = session.emitted
            out = session.push(piece, final=chunk_final, softmax=softmax)
            # keep local frames [50, 450)
            keep_lo = max(self.overlap - local_before, 0)
            keep_hi = min(self.chunk - self.overlap - local_before, out.shape[-1])
            if keep_hi > keep_lo:
                pieces.append(out[..., keep_lo:keep_hi])
        if pieces:
            result = torch.cat(pieces, dim=-1)
        else:
            result = torch.empty(self.streams, self.engine.cfg.output_channels, 0,
                                 dtype=torch.float32, device=device)
        self.emitted += result.shape[-1]
        return result
