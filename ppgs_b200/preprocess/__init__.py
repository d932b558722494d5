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


def from_audio(audio, representation=config.live('REPRESENTATION'),
               sample_rate=config.SAMPLE_RATE, gpu=None):
    """Preprocess audio (ppgs/preprocess/core.py:194-216)."""
    representation = config.resolve(representation)
    from ..core import resample
    audio = resample(audio, sample_rate)
    features = get(representation).from_audio(
        audio, sample_rate=config.SAMPLE_RATE, gpu=gpu)
    if features.dim() == 2:
        features = features[None]
    return features


def save_masked(tensor, file, length):
    """Save masked tensor (ppgs/preprocess/core.py:219-221): the native writer for 2-D
    fp16 / fp32 CPU tensors (no pickling of the padded row, no GIL while writing), torch.save
    otherwise."""
    import ctypes
    import os
    from .. import _lib
    length = int(length)
    if (tensor.dim() == 2 and tensor.device.type == 'cpu' and tensor.stride(-1) == 1 and
            tensor.dtype in (torch.float16, torch.float32) and 0 <= length <= tensor.shape[-1] and
            tensor.shape[0] * length < 2 ** 30):   # the native writer's zip fields; larger -> torch.save
        write = (_lib.lib.ppgs_pt_write_f16 if tensor.dtype == torch.float16
                 else _lib.lib.ppgs_pt_write_f32)
        _lib.check(write(os.fsencode(str(file)), ctypes.c_void_p(tensor.data_ptr()),
                         tensor.shape[0], length, tensor.stride(0)))
        return
    torch.save(tensor[..., :length].clone(), file)


def from_files_to_files(audio_files, output_files, representations=None, num_workers=0, gpu=None):
    """Preprocess from files (ppgs/preprocess/core.py:63-97): every representation of every
    audio file -> `output_file` (a `{}` in the name is replaced by the representation), fp16
    tensors of shape (channels, samples // 160)."""
    from .. import data
    if len(audio_files) != len(output_files):
        raise ValueError('audio_files and output_files must have equal lengths')
    dataloader = data.loader(
        audio_files,
        features=['audio', 'length', 'audio_file'],
        num_workers=num_workers // 2,
        max_frames=config.MAX_PREPROCESS_FRAMES)
    from_dataloader(
        dataloader,
        representations,
        dict(zip(audio_files, output_files)),
        num_workers=(num_workers + 1) // 2,
        gpu=gpu)


def from_dataloader(loader, representations, output, num_workers=0, gpu=None):
    """Preprocess from a dataloader yielding (audio, length, filename) batches
    (ppgs/preprocess/core.py:105-190).  The reference pickles every batch to a spawn pool;
    here `num_workers` writer threads crop and write the pinned host copy of the features
    while the next batch is on the GPU."""
    import concurrent.futures
    from .. import load
    if representations is None:
        representations = [config.REPRESENTATION]
    if isinstance(representations, str):
        representations = [representations]
    modules = [get(representation) for representation in representations]   # ValueError if unknown
    device = load.resolve_device(gpu)
    pool = concurrent.futures.ThreadPoolExecutor(num_workers) if num_workers else None
    pending = []

    def save(done, host, filenames, lengths):
        done.synchronize()
        for row, filename, length in zip(host, filenames, lengths):
            save_masked(row, filename, length)

    try:
        for audios, lengths, audio_files in loader:
            audios = audios.to(device, non_blocking=True)
            frame_lengths = (lengths // config.HOPSIZE).tolist()
            for representation, module in zip(representations, modules):
                features = module.from_audios(audios, lengths, gpu=device.index)
                if features.requires_grad:
                    raise ValueError('All representations should be detached')
                host = torch.empty(features.shape, dtype=features.dtype, pin_memory=True)
                host.copy_(features, non_blocking=True)
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(device))
                filenames = [
                    str(output[file]).format(representation) if '{}' in str(output[file])
                    else output[file] for file in audio_files]
                if pool is None:
                    save(done, host, filenames, frame_lengths)
                else:
                    pending.append(pool.submit(save, done, host, filenames, frame_lengths))
                    while len(pending) > 4 * num_workers:      # back-pressure (core.py:178-180)
                        pending.pop(0).result()
        for future in pending:
            future.result()
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
