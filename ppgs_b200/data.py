"""Batch scheduler for the file API — the host-side replacement of
ppgs/data/{loader,dataset,sampler,collate}.py on the inference path
(SURVEY.md §8 a11): file list -> frame lengths from the WAV headers ->
frame-budget batches -> zero-padded (B,1,max_samples) pinned fp32 batches.

Batch composition is part of the reference's semantics (a padded row's result
depends on the batch's padded length, SURVEY.md §3.2 ii-iii), so the packing
below reproduces the reference sampler's batches exactly (same seed, same
greedy rule); only the machinery differs: reader *threads* decoding straight
into a pinned batch buffer instead of DataLoader worker processes + collate.
"""
import concurrent.futures
import warnings

import numpy as np
import torch

from . import config
from . import load


class Metadata:
    """Frame lengths of a list of audio files (ppgs/data/dataset.py:136-216,
    list-of-files branch).  Files longer than `max_frames` are skipped with the
    reference's warning."""

    def __init__(self, audio_files, max_frames=config.live('MAX_INFERENCE_FRAMES')):
        max_frames = config.resolve(max_frames)
        self.audio_files, self.lengths, self.samples = [], [], {}
        # every file is 16-bit PCM WAVE or <= 16-bit FLAC at 16 kHz: eligible for ppgs_files_to_files
        self.native = True
        audio_files = list(audio_files)
        for audio_file, info in zip(audio_files, load.wav_info_many(audio_files)):
            if info is None:
                flac = load.flac_info(audio_file) if str(audio_file).lower().endswith('.flac') else None
                if flac is not None:     # the native pipeline decodes <= 16-bit FLAC at 16 kHz too
                    samples, sample_rate = flac['samples'], flac['sample_rate']
                    if flac['bits'] > 16 or sample_rate != config.SAMPLE_RATE:
                        self.native = False
                else:
                    samples, sample_rate = load.wav_num_frames(audio_file)
                    self.native = False
            else:
                samples, sample_rate = info['samples'], info['sample_rate']
                if info['bits'] != 16 or info['is_float'] or sample_rate != config.SAMPLE_RATE:
                    self.native = False
            length = int(samples * (config.SAMPLE_RATE / sample_rate)) // config.HOPSIZE
            if length <= max_frames:
                self.audio_files.append(audio_file)
                self.lengths.append(length)
                self.samples[audio_file] = samples
            else:
                warnings.warn(
                    f'File {audio_file} of length {length} '
                    f'exceeds max_frames of {max_frames}. Skipping.')

    def __len__(self):
        return len(self.audio_files)


def frame_budget_batches(lengths, max_frames, seed=config.RANDOM_SEED, epoch=0,
                         buckets=config.BUCKETS):
    """Batches of file indices with (n+1)*max_len <= max_frames
    (ppgs/data/sampler.py:46-82 over the buckets of ppgs/data/dataset.py:107-127).
    Deterministic in (seed + epoch)."""
    generator = torch.Generator()
    generator.manual_seed(seed + epoch)
    lengths = np.asarray(lengths)
    order = np.argsort(lengths)
    pairs = np.stack((order, lengths[order])).T if len(order) else np.zeros((0, 2), int)
    size = max(len(pairs) // buckets, 1)
    groups = [pairs[i * size:(i + 1) * size] for i in range(buckets)]
    if len(pairs) > buckets * size:   # remainder joins the last bucket
        groups[-1] = np.concatenate((groups[-1], pairs[buckets * size:]))
    batches = []
    for group in groups:
        if not len(group):
            continue
        group = group[torch.randperm(len(group), generator=generator).tolist()]
        batch, longest = [], 0
        for index, length in group:
            longest = max(longest, length)
            if batch and (len(batch) + 1) * longest > max_frames:
                batches.append(batch)
                longest = length
                batch = [int(index)]
            else:
                batch.append(int(index))
        if batch:
            batches.append(batch)
    return [batches[i] for i in torch.randperm(len(batches), generator=generator).tolist()]


def collate(audios, pin_memory=False):
    """Zero-pad (1,samples_i) audios into (B,1,max_samples) fp32 + int64 sample
    lengths (ppgs/data/collate.py:19-28)."""
    lengths = torch.tensor([audio.shape[-1] for audio in audios], dtype=torch.long)
    padded = torch.zeros(
        len(audios), 1, int(lengths.max()), dtype=torch.float32, pin_memory=pin_memory)
    for row, audio in zip(padded, audios):
        row[:, :audio.shape[-1]] = audio[:1]
    return padded, lengths


# Engine workspace per output frame at the default model: activations of 1.25 computed frames
# (x, QKV, attention output, FFN hidden as split planes) + features + posteriors
WORKSPACE_BYTES_PER_FRAME = 16 * 1024


def bounded_max_frames(max_frames, device=None):
    """The reference's default `max_frames` is infinite (ppgs/config/static.py:22): every file
    lands in ONE batch, which cannot work for a corpus.  An infinite budget is replaced by
    what a quarter of the free memory of `device` (the GPU that will run the batches; None =
    the current one) holds (warned once); finite budgets are kept."""
    if max_frames != float('inf') or not torch.cuda.is_available():
        return max_frames
    free, _ = torch.cuda.mem_get_info(device)
    return max(int(free // 4 // WORKSPACE_BYTES_PER_FRAME), 1000)


class Loader:
    """Iterable of (audio (B,1,max_samples) pinned fp32, lengths (B,) int64
    samples, [audio_file]) — what `ppgs.data.loader(files, ['audio','length',
    'audio_file'], num_workers, max_frames)` yields (ppgs/data/loader.py:20-43).
    `num_workers` reader threads decode files; up to `prefetch` batches ahead."""

    def __init__(self, audio_files, num_workers=0, max_frames=config.live('MAX_INFERENCE_FRAMES'),
                 prefetch=2, shard=None, dataset=None, device=None):
        # `dataset`: reuse the header probe of another shard's loader; `device`: the GPU the
        # batches will run on (sizes an unbounded frame budget from ITS free memory)
        max_frames = config.resolve(max_frames)
        self.dataset = dataset if dataset is not None else Metadata(audio_files, max_frames)
        budget = bounded_max_frames(max_frames, device)
        if budget != max_frames and sum(self.dataset.lengths) > budget:
            warnings.warn(
                f'max_frames is unbounded and the files hold {sum(self.dataset.lengths)} frames: '
                f'batching with max_frames={budget} (a quarter of the free GPU memory)')
            max_frames = budget
        self.batches = frame_budget_batches(self.dataset.lengths, max_frames)
        if shard is not None:
            rank, world = shard
            self.batches = self.batches[rank::world]
        self.num_workers = max(int(num_workers), 0)
        self.prefetch = prefetch
        self.pin = torch.cuda.is_available()

    def __len__(self):
        return len(self.batches)

    def run_native(self, engine, output_files, workers=1, legacy_mode=False):
        """This loader's batches through the native pipeline (reader / writer threads
        and the GPU loop all inside ppgs_files_to_files).  Returns frames written."""
        batches = [[self.dataset.audio_files[i] for i in batch] for batch in self.batches]
        return engine.files_to_files(
            batches, output_files, self.dataset.samples,
            reader_threads=max(workers, 1), writer_threads=max(workers, 1),
            legacy_mode=legacy_mode)

    def _load(self, batch):
        files = [self.dataset.audio_files[i] for i in batch]
        audios = [load.audio(file) for file in files]
        padded, lengths = collate(audios, self.pin)
        return padded, lengths, files

    def __iter__(self):
        if self.num_workers == 0:
            for batch in self.batches:
                yield self._load(batch)
            return
        with concurrent.futures.ThreadPoolExecutor(self.num_workers) as pool:
            pending = []
            batches = iter(self.batches)
            for batch in batches:
                pending.append(pool.submit(self._load, batch))
                if len(pending) > self.prefetch:
                    yield pending.pop(0).result()
            for future in pending:
                yield future.result()


def loader(audio_files, features=('audio', 'length', 'audio_file'), num_workers=0,
           max_frames=config.live('MAX_INFERENCE_FRAMES'), shard=None, dataset=None, device=None):
    """ppgs.data.loader for the inference feature set."""
    if list(features) != ['audio', 'length', 'audio_file']:
        raise ValueError(
            "ppgs_b200.data.loader serves the inference path only: "
            "features must be ['audio', 'length', 'audio_file']")
    return Loader(audio_files, num_workers, max_frames, shard=shard, dataset=dataset, device=device)
