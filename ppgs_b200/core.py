"""Inference API of ppgs_b200 — same names, argument order and defaults as
ppgs/core.py:22-391 of the reference; the arithmetic runs in libppgs_b200.so.

Differences from the reference that a caller can observe:
* every function needs a CUDA device (`gpu=None` means the current CUDA device;
  the reference would use the CPU) — there is no CPU / PyTorch fallback;
* `from_audio` accepts batch > 1 (the reference builds a length tensor of shape
  (1,) and fails for B > 1);
* results are fp32 CUDA tensors; numerics are the reference's modules evaluated
  in fp32 with autocast disabled (oracle mode O3), not bf16/fp16 autocast;
* `gpu` may also be a list of ordinals in `from_files_to_files` to shard the file
  list across GPUs of one box from a single process group (see parallel.py).
"""
import os
import queue
import threading
from typing import Dict, List, Optional, Union

import torch

from . import config
from . import data
from . import load
from . import preprocess

###############################################################################
# Application programming interface
###############################################################################


def from_audio(
    audio: torch.Tensor,
    sample_rate: Union[int, float],
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Optional[Union[str, bytes, os.PathLike]] = None,
    gpu: Optional[int] = None,
    legacy_mode: bool = False
) -> torch.Tensor:
    """Infer ppgs from audio (ppgs/core.py:22-69)

    Arguments
        audio: batched audio, shape=(batch, 1, samples) (or (1, samples))
        sample_rate: audio sampling rate
        representation: 'mel' or 'w2v2fb'
        checkpoint: the checkpoint file
        gpu: CUDA ordinal
        legacy_mode: use legacy (unchunked) inference

    Returns
        ppgs, shape=(batch, len(ppgs.PHONEMES), frames), fp32 on the GPU
    """
    representation = config.resolve(representation)
    if audio.dim() == 2:
        audio = audio.unsqueeze(0)
    engine = load.model(checkpoint, representation, gpu)
    if sample_rate != config.SAMPLE_RATE:
        audio = engine.resample(audio, sample_rate)
    if _fused_frontend(representation):
        # mel front-end + transformer in one C-ABI call; features stay on-chip
        # in the engine workspace (ppgs_from_audio)
        return engine.from_audio(audio, softmax=True, legacy_mode=legacy_mode)
    features = preprocess.from_audio(
        audio, representation=representation,
        sample_rate=config.SAMPLE_RATE, gpu=gpu)
    lengths = torch.full((features.shape[0],), features.shape[-1], dtype=torch.long)
    return from_features(features, lengths, representation, checkpoint, gpu,
                         legacy_mode=legacy_mode)


def from_features(
    features: torch.Tensor,
    lengths: torch.Tensor,
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Optional[Union[str, bytes, os.PathLike]] = None,
    gpu: Optional[int] = None,
    softmax: bool = True,
    legacy_mode: bool = False
) -> torch.Tensor:
    """Infer ppgs from input features (ppgs/core.py:72-128)

    features: shape=(batch, channels, frames); lengths: shape=(batch,)
    """
    representation = config.resolve(representation)
    return infer(features, lengths, representation, checkpoint, softmax,
                 legacy_mode, gpu=gpu)


def from_file(
    file: Union[str, bytes, os.PathLike],
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Optional[Union[str, bytes, os.PathLike]] = None,
    gpu: Optional[int] = None,
    legacy_mode: bool = False
) -> torch.Tensor:
    """Infer ppgs from an audio file (ppgs/core.py:131-168);
    returns shape=(len(ppgs.PHONEMES), frames)"""
    representation = config.resolve(representation)
    audio = load.audio(file, device=load.resolve_device(gpu))
    return from_audio(
        audio, config.SAMPLE_RATE, representation, checkpoint, gpu, legacy_mode
    ).squeeze(0)


def from_file_to_file(
    audio_file: Union[str, bytes, os.PathLike],
    output_file: Union[str, bytes, os.PathLike],
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Optional[Union[str, bytes, os.PathLike]] = None,
    gpu: Optional[int] = None,
    legacy_mode: bool = False
) -> None:
    """Infer ppg from an audio file and save to a torch tensor file
    (ppgs/core.py:171-204)"""
    representation = config.resolve(representation)
    result = from_file(audio_file, representation, checkpoint, gpu, legacy_mode)
    preprocess.save_masked(result.detach().cpu(), output_file, result.shape[-1])


def from_files_to_files(
    audio_files: List[Union[str, bytes, os.PathLike]],
    output_files: List[Union[str, bytes, os.PathLike]],
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Optional[Union[str, bytes, os.PathLike]] = None,
    num_workers: int = 0,
    gpu: Optional[Union[int, List[int]]] = None,
    max_frames: int = config.live('MAX_INFERENCE_FRAMES'),
    legacy_mode: bool = False
) -> None:
    """Infer ppgs from audio files and save to torch tensor files
    (ppgs/core.py:207-272).  `num_workers` == 0: one file per call like the
    reference; > 0: frame-budget batches from `data.loader` (reader threads) and
    `num_workers // 2` writer threads.  `gpu` may be a list of CUDA ordinals: the
    batch list is dealt round-robin to one pipeline per GPU (BASELINE config 5 from
    a single process; `parallel.from_files_to_files` is the torchrun form)."""
    representation = config.resolve(representation)
    max_frames = config.resolve(max_frames)
    if len(audio_files) != len(output_files):
        raise ValueError('audio_files and output_files must have equal lengths')
    if isinstance(gpu, (list, tuple)):
        return _from_files_to_files_sharded(
            audio_files, output_files, representation, checkpoint, num_workers, list(gpu),
            max_frames, legacy_mode)
    if num_workers == 0:
        for audio_file, output_file in zip(audio_files, output_files):
            from_file_to_file(
                audio_file, output_file, representation, checkpoint, gpu, legacy_mode)
        return
    mapping = {
        audio_file: output_file
        for audio_file, output_file in zip(audio_files, output_files)}
    dataloader = data.loader(
        audio_files,
        features=['audio', 'length', 'audio_file'],
        num_workers=num_workers // 2,
        max_frames=max_frames,
        device=gpu)
    if _native_pipeline(dataloader, representation):
        engine = load.model(checkpoint, representation, gpu)
        dataloader.run_native(engine, mapping, num_workers // 2, legacy_mode)
        return
    from_dataloader(
        dataloader, mapping, representation, checkpoint,
        save_workers=num_workers // 2, gpu=gpu, legacy_mode=legacy_mode)


def _from_files_to_files_sharded(audio_files, output_files, representation, checkpoint,
                                 num_workers, gpus, max_frames, legacy_mode):
    """One pipeline thread per GPU over batches i::len(gpus) of the SAME deterministic
    batch list (so every posterior equals the single-GPU run).  The checkpoint is read
    once; the other GPUs adopt the first engine's packed weight blob through one
    device-to-device copy each (the single-process form of the NCCL broadcast)."""
    if not gpus:
        raise ValueError('gpu: empty list of devices')
    workers = max(num_workers // 2, 1)
    mapping = dict(zip(audio_files, output_files))
    first = load.model(checkpoint, representation, gpus[0])
    engines = [first]
    for gpu in gpus[1:]:
        key = load.cache_key(representation, checkpoint, gpu)
        with load._lock:
            engine = load._engines.get(key)
        if engine is None or engine is first:
            from .engine import Engine
            engine = Engine(gpu, is_causal=bool(first.cfg.is_causal),
                            **load.model_kwargs(representation))
            engine.blob().copy_(first.blob())
            torch.cuda.synchronize(engine.device)
            engine.adopt_blob()
            engine.precision = first.precision
            if gpu != gpus[0]:
                with load._lock:
                    load._engines[key] = engine
        engines.append(engine)
    loaders = []
    for rank in range(len(gpus)):
        loaders.append(data.loader(
            audio_files, num_workers=workers, max_frames=max_frames, shard=(rank, len(gpus)),
            dataset=loaders[0].dataset if loaders else None, device=gpus[rank]))
    errors = []

    def run(engine, dataloader):
        try:
            with torch.cuda.device(engine.device):
                if _native_pipeline(dataloader, representation):
                    dataloader.run_native(engine, mapping, workers, legacy_mode)
                else:
                    from_dataloader(dataloader, mapping, representation, checkpoint,
                                    save_workers=workers, gpu=engine.device.index,
                                    legacy_mode=legacy_mode, engine=engine)
        except Exception as error:      # surfaced after every shard has stopped
            errors.append(error)

    threads = [threading.Thread(target=run, args=pair) for pair in zip(engines, loaders)]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    if errors:
        raise errors[0]


###############################################################################
# Batched file inference
###############################################################################


def from_dataloader(
    dataloader,
    output_files: Dict[
        Union[str, bytes, os.PathLike],
        Union[str, bytes, os.PathLike]],
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Union[str, bytes, os.PathLike] = None,
    save_workers: int = 1,
    gpu: Optional[int] = None,
    legacy_mode: bool = False,
    engine=None
) -> None:
    """Infer ppgs from a dataloader yielding (audio, length, audio_filename)
    batches (ppgs/core.py:280-391).  The reference pickles every result to a
    spawn Pool; here writer *threads* take (pinned host tensor, filename,
    frames) items from a bounded queue, so the D2H copy of batch i overlaps the
    kernels of batch i+1."""
    representation = config.resolve(representation)
    if engine is None:
        engine = load.model(checkpoint, representation, gpu)
    writer = _Writer(save_workers)
    try:
        for audios, lengths, audio_files in dataloader:
            frame_lengths = lengths // config.HOPSIZE
            if _fused_frontend(representation):
                result = engine.from_audio(
                    audios, lengths=lengths, softmax=True, legacy_mode=legacy_mode)
            else:
                features = preprocess.get(representation).from_audios(
                    audios, lengths, gpu=gpu)
                if features.requires_grad:
                    raise ValueError('All representations should be detached')
                result = engine.transformer(
                    features, frame_lengths, softmax=True, legacy_mode=legacy_mode)
            host = torch.empty(result.shape, dtype=result.dtype, pin_memory=True)
            host.copy_(result, non_blocking=True)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(engine.device))
            filenames = [output_files[file] for file in audio_files]
            writer.put(done, host, filenames, frame_lengths.tolist())
    finally:
        writer.close()


def from_feature_files_to_files(
    feature_files: List[Union[str, bytes, os.PathLike]],
    output_files: List[Union[str, bytes, os.PathLike]],
    representation: str = config.live('REPRESENTATION'),
    checkpoint: Optional[Union[str, bytes, os.PathLike]] = None,
    num_workers: int = 0,
    gpu: Optional[int] = None,
    max_frames: int = config.live('MAX_INFERENCE_FRAMES'),
    legacy_mode: bool = False,
    container: Optional[Union[str, bytes, os.PathLike]] = None
) -> None:
    """Infer ppgs from cached input features (the `<stem>-mel.pt` / `<stem>-w2v2fb.pt` files
    `python -m ppgs.preprocess` writes and ppgs/data/dataset.py:98-101 reads back with
    torch.load): each file holds a (channels, frames) fp16 tensor.  Files are read by the native
    `.pt` reader straight into pinned padded batches (frame-budget batches like the audio path),
    run through `ppgs_transformer_forward`, and written cropped, one `.pt` per input.

    `container`: write ONE file per batch instead, `<container>.<batch:05d>.pt` =
    {'files': [output names], 'lengths': int64 (B,), 'ppgs': float32 (B, 40, max_frames)} — the
    sharded / batched output option for corpora where 10^5 small files are the bottleneck."""
    representation = config.resolve(representation)
    max_frames = config.resolve(max_frames)
    if len(feature_files) != len(output_files):
        raise ValueError('feature_files and output_files must have equal lengths')
    engine = load.model(checkpoint, representation, gpu)
    channels = engine.cfg.input_channels
    shapes = []
    for file in feature_files:
        info = load.tensor_info(file)
        shape = info[0] if info is not None else tuple(torch.load(file, map_location='cpu').shape)
        if len(shape) != 2 or shape[0] != channels:
            raise ValueError(f'{file}: expected features of shape ({channels}, frames), got {shape}')
        shapes.append(shape)
    lengths = [shape[-1] for shape in shapes]
    budget = data.bounded_max_frames(max_frames, engine.device)
    batches = data.frame_budget_batches(lengths, budget)
    writer = _Writer(max(num_workers // 2, 1) if num_workers else 0)
    readers = max(num_workers - num_workers // 2, 1)
    import concurrent.futures
    pool = concurrent.futures.ThreadPoolExecutor(readers)
    try:
        for index, batch in enumerate(batches):
            frames = max(lengths[i] for i in batch)
            host = torch.zeros(len(batch), channels, frames, dtype=torch.float16, pin_memory=True)
            list(pool.map(lambda item: load.features(feature_files[item[1]], out=host[item[0]]),
                          enumerate(batch)))
            batch_lengths = torch.tensor([lengths[i] for i in batch], dtype=torch.long)
            result = engine.transformer(host.to(engine.device, non_blocking=True), batch_lengths,
                                        softmax=True, legacy_mode=legacy_mode)
            out_host = torch.empty(result.shape, dtype=result.dtype, pin_memory=True)
            out_host.copy_(result, non_blocking=True)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(engine.device))
            names = [output_files[i] for i in batch]
            if container is None:
                writer.put(done, out_host, names, batch_lengths.tolist())
            else:
                writer.put_container(done, out_host, [str(name) for name in names], batch_lengths,
                                     f'{os.fspath(container)}.{index:05d}.pt')
    finally:
        pool.shutdown()
        writer.close()


class _Writer:
    """Bounded queue + `workers` saver threads (0 = save synchronously).
    Replaces the spawn Pool + qsize back-pressure of ppgs/core.py:311-314,358-365."""

    def __init__(self, workers):
        self.workers = max(int(workers), 0)
        self.errors = []
        if self.workers:
            self.queue = queue.Queue(maxsize=4 * self.workers)
            self.threads = [
                threading.Thread(target=self._run, daemon=True)
                for _ in range(self.workers)]
            for thread in self.threads:
                thread.start()

    @staticmethod
    def _save(done, host, filenames, frames):
        done.synchronize()
        for ppg, filename, length in zip(host, filenames, frames):
            preprocess.save_masked(ppg, filename, int(length))

    def _run(self):
        while True:
            item = self.queue.get()
            if item is None:
                return
            try:
                (self._save_container if len(item) == 5 else self._save)(*item)
            except Exception as error:   # surfaced by close()
                self.errors.append(error)

    def put(self, done, host, filenames, frames):
        if self.workers:
            self.queue.put((done, host, filenames, frames))
        else:
            self._save(done, host, filenames, frames)

    @staticmethod
    def _save_container(done, host, filenames, lengths, file):
        done.synchronize()
        torch.save({'files': filenames, 'lengths': lengths, 'ppgs': host.clone()}, file)

    def put_container(self, done, host, filenames, lengths, file):
        """One file for the whole batch (padded tensor + lengths + names)."""
        item = (done, host, filenames, lengths, file)
        if self.workers:
            self.queue.put(item)
        else:
            self._save_container(*item)

    def close(self):
        if self.workers:
            for _ in self.threads:
                self.queue.put(None)
            for thread in self.threads:
                thread.join()
        if self.errors:
            raise self.errors[0]


###############################################################################
# PPG distance, interpolation, sparsification (device kernels, csrc/postops.cu)
###############################################################################


def _post_engine(tensor):
    return load.utility_engine(tensor.device.index if tensor.is_cuda else None)


def _stream(engine):
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream(engine.device).cuda_stream)


def _f32(engine, tensor):
    return tensor.to(engine.device, torch.float32).contiguous()


def _back(result, like):
    """Results live where the (first) input lives, like the reference's."""
    return result if like.is_cuda else result.cpu()


def distance(
    ppgX: torch.Tensor,
    ppgY: torch.Tensor,
    reduction: str = 'mean',
    normalize: bool = True,
    exponent: float = None,
    similarity: Optional[torch.Tensor] = None
) -> torch.Tensor:
    """Compute the pronunciation distance between two aligned PPGs
    (ppgs/core.py:399-469): similarity-weighted Jensen-Shannon distance, one CUDA kernel.

    ppgX, ppgY: shape=(len(ppgs.PHONEMES), frames); reduction in ['mean', 'none', 'sum'];
    `similarity`: the (phonemes, phonemes) similarity matrix — default: the tensor stored
    at config.SIMILARITY_MATRIX_PATH (the reference's assets/balanced_similarity.pt)."""
    import ctypes
    from . import _lib
    if reduction is None:
        reduction = 'none'
    if reduction not in ('mean', 'none', 'sum'):
        raise ValueError(f'Reduction method {reduction} not defined')
    if exponent is None:
        exponent = config.SIMILARITY_EXPONENT
    engine = _post_engine(ppgX)
    x, y = _f32(engine, ppgX), _f32(engine, ppgY)
    if x.dim() != 2 or x.shape != y.shape:
        raise ValueError('ppgX and ppgY must both have shape (phonemes, frames)')
    weights = None
    if normalize:
        if similarity is None:
            if config.SIMILARITY_MATRIX_PATH is None:
                raise ValueError(
                    'distance(normalize=True) needs the phoneme similarity matrix: pass '
                    'similarity= or set ppgs_b200.config.SIMILARITY_MATRIX_PATH')
            key = (str(config.SIMILARITY_MATRIX_PATH), engine.device.index)
            if getattr(distance, 'key', None) != key:    # the cache of ppgs/core.py:436-442
                distance.similarity_matrix = torch.load(config.SIMILARITY_MATRIX_PATH)
                distance.key = key
            similarity = distance.similarity_matrix
        weights = _f32(engine, similarity)
        if weights.shape != (x.shape[0], x.shape[0]):
            raise ValueError('similarity must have shape (phonemes, phonemes)')
    phonemes, frames = x.shape
    out = torch.empty(frames if reduction == 'none' else 1, dtype=torch.float32, device=engine.device)
    _lib.check(_lib.lib.ppgs_ppg_distance(
        engine._handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), phonemes,
        frames, x.stride(0), y.stride(0),
        ctypes.c_void_p(weights.data_ptr()) if weights is not None else None, float(exponent),
        {'none': 0, 'mean': 1, 'sum': 2}[reduction], ctypes.c_void_p(out.data_ptr()),
        _stream(engine)))
    return _back(out if reduction == 'none' else out[0], ppgX)


def interpolate(
    ppgX: torch.Tensor,
    ppgY: torch.Tensor,
    interp: Union[float, torch.Tensor]
) -> torch.Tensor:
    """Linear interpolation (ppgs/core.py:475-496): (1 - interp) * ppgX + interp * ppgY,
    interp a scalar or shape=(frames,)."""
    import ctypes
    from . import _lib
    engine = _post_engine(ppgX)
    x, y = _f32(engine, ppgX), _f32(engine, ppgY)
    if x.shape != y.shape:
        raise ValueError('ppgX and ppgY must have the same shape')
    frames = x.shape[-1]
    weights, scalar = None, 0.0
    if torch.is_tensor(interp) and interp.dim() > 0:
        if interp.shape != (frames,):
            raise ValueError('interp must be a scalar or have shape (frames,)')
        weights = _f32(engine, interp)
    else:
        scalar = float(interp)
    out = torch.empty_like(x)
    _lib.check(_lib.lib.ppgs_ppg_interpolate(
        engine._handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()),
        ctypes.c_void_p(weights.data_ptr()) if weights is not None else None, scalar,
        x.numel() // max(frames, 1), frames, ctypes.c_void_p(out.data_ptr()), _stream(engine)))
    return _back(out, ppgX)


def sparsify(
    ppg: torch.Tensor,
    method: str = 'percentile',
    threshold: Union[float, int, torch.Tensor] = 0.85
) -> torch.Tensor:
    """Make phonetic posteriorgrams sparse (ppgs/core.py:504-543).

    ppg: shape=(batch, len(ppgs.PHONEMES), frames); method in ['constant', 'percentile',
    'topk'].  Like the reference, a 1-element *tensor* threshold with 'percentile' adds a
    leading axis to the result (torch.quantile keeps the q axis); 'topk' treats every batch
    row like the reference treats a single-row batch.

    Deviations from the reference (also listed in INTEGRATION.md): the default threshold is the
    float 0.85, so a default call returns (batch, phonemes, frames), where the reference's
    default `torch.Tensor([0.85])` returns (1, batch, phonemes, frames) — pass
    `torch.tensor([0.85])` for that shape; and the result is a new tensor (the reference's
    'topk' zeroes its input in place)."""
    import ctypes
    from . import _lib
    methods = {'constant': 0, 'percentile': 1, 'topk': 2}
    if method not in methods:
        raise ValueError(f'Sparsification method {method} not defined')
    engine = _post_engine(ppg)
    x = _f32(engine, ppg)
    squeeze = x.dim() == 2
    if squeeze:
        x = x[None]
    if x.dim() != 3:
        raise ValueError('ppg must have shape (batch, phonemes, frames)')
    extra_axis = method == 'percentile' and torch.is_tensor(threshold) and threshold.dim() > 0
    value = float(threshold.reshape(-1)[0]) if torch.is_tensor(threshold) else float(threshold)
    batch, phonemes, frames = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib.ppgs_ppg_sparsify(
        engine._handle, ctypes.c_void_p(x.data_ptr()), batch, phonemes, frames, methods[method],
        value, ctypes.c_void_p(out.data_ptr()), _stream(engine)))
    if squeeze:
        out = out[0]
    if extra_axis:
        out = out[None]
    return _back(out, ppg)


###############################################################################
# Inference
###############################################################################


def infer(
    features,
    lengths,
    representation=config.live('REPRESENTATION'),
    checkpoint=None,
    softmax=True,
    legacy_mode=False,
    gpu=None
):
    """Perform model inference (ppgs/core.py:551-596): cached engine per
    (representation, checkpoint, device); logits -> softmax(dim=1) in-kernel."""
    representation = config.resolve(representation)
    if gpu is None and features.is_cuda:
        gpu = features.device.index
    engine = load.model(checkpoint, representation, gpu)
    return engine.transformer(features, lengths, softmax=softmax, legacy_mode=legacy_mode)


def resample(audio, sample_rate, target_rate=config.SAMPLE_RATE):
    """Perform audio resampling (ppgs/core.py:599-608): the arithmetic of
    torchaudio.transforms.Resample(sample_rate, target_rate) as a CUDA kernel
    (ppgs_resample).  The result lives where `audio` lives, like the reference's."""
    if sample_rate == target_rate:
        return audio
    gpu = audio.device.index if audio.is_cuda else None
    result = load.utility_engine(gpu).resample(audio, sample_rate, target_rate)
    return result if audio.is_cuda else result.cpu()


def representation_file_extension():
    """ppgs/core.py:611-621"""
    if (config.REPRESENTATION == config.BEST_REPRESENTATION and
            config.REPRESENTATION_KIND == 'ppg'):
        return '-ppg.pt'
    if config.REPRESENTATION_KIND == 'ppg':
        return f'-{config.REPRESENTATION}-ppg.pt'
    return f'-{config.REPRESENTATION}.pt'


def _native_pipeline(dataloader, representation):
    """The whole loop runs inside libppgs_b200 (ppgs_files_to_files) when the
    representation is mel and every file is 16-bit PCM WAVE or <= 16-bit FLAC at 16 kHz; otherwise the
    batches come from the Python reader threads (same batches, same kernels).
    PPGS_B200_NATIVE_FILES=0 forces the latter."""
    if os.environ.get('PPGS_B200_NATIVE_FILES', '1') == '0':
        return False
    return _fused_frontend(representation) and dataloader.dataset.native


def _fused_frontend(representation):
    return (representation if representation is not None else config.REPRESENTATION) == 'mel'
