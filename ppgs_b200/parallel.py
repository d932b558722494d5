"""Multi-GPU plumbing: one process per GPU (torchrun), utterances / files
sharded on the batch axis, ONE broadcast of the packed weight blob at load and no
per-step collective (SURVEY.md §8e — the reference has no distributed code at
all, ppgs/train/core.py:24-26 is commented out)."""
import os

import torch
import torch.distributed as dist

from . import config
from . import load
from .engine import Engine


def init(backend=None):
    """Join the torchrun process group (RANK / WORLD_SIZE / MASTER_* from the
    environment); no-op for a single process.  Returns (rank, world_size)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        if backend is None:
            backend = os.environ.get('PPGS_B200_DIST_BACKEND') or (
                'nccl' if torch.cuda.is_available() else 'gloo')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def local_device():
    """CUDA ordinal of this rank: LOCAL_RANK, folded onto the visible devices (several ranks
    may share one GPU in tests)."""
    count = max(torch.cuda.device_count(), 1)
    return int(os.environ.get('LOCAL_RANK', '0')) % count


def broadcast_engine(state_dict=None, representation=None, gpu=None, src=0, **kwargs):
    """Build this rank's engine with rank `src`'s weights: only `src` needs the
    state dict; its packed blob (one device allocation whose layout is a pure
    function of the model config) goes to every other rank in ONE NCCL broadcast
    over NVLink, and those ranks adopt it without touching the checkpoint."""
    gpu = local_device() if gpu is None else gpu
    # the live configuration (configure() / --config), not import-time defaults: every rank
    # must build the engine load.cache_key() will describe
    kwargs = {'is_causal': config.IS_CAUSAL, **load.model_kwargs(representation), **kwargs}
    engine = Engine(gpu, **kwargs)
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == src:
        if state_dict is None:
            raise ValueError('the source rank needs the state dict')
        engine.load_state_dict(state_dict)
    if dist.is_initialized() and dist.get_world_size() > 1:
        blob = engine.blob()
        dist.broadcast(blob, src=src)
        torch.cuda.synchronize(engine.device)
        if rank != src:
            engine.adopt_blob()
    return engine


def shard(items, rank=None, world=None):
    """Contiguous-by-stride shard of a list of work items (utterances, files or
    batches) for this rank; no data-path collective."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return list(items)[rank::world]


def shard_by_frames(frame_lengths, world):
    """Longest-processing-time assignment of utterances to ranks so every GPU
    gets about the same number of frames; returns one index list per rank."""
    order = sorted(range(len(frame_lengths)), key=lambda i: -frame_lengths[i])
    loads = [0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=loads.__getitem__)
        parts[r].append(i)
        loads[r] += frame_lengths[i]
    return [sorted(p) for p in parts]


def from_files_to_files(audio_files, output_files, representation=config.live('REPRESENTATION'),
                        checkpoint=None, num_workers=2, max_frames=64000, legacy_mode=False):
    """Sharded `from_files_to_files` under torchrun: every rank builds the same
    deterministic batch list (data.frame_budget_batches) and takes batches
    rank::world, so batch composition — and therefore every posterior — is
    identical to a single-GPU run."""
    from . import core, data
    representation = config.resolve(representation)
    rank, world = init()
    gpu = local_device()
    key = load.cache_key(representation, checkpoint, gpu)
    with load._lock:
        cached = key in load._engines
    if world > 1 and not cached:   # one broadcast per (representation, checkpoint, device), like load.model
        state = load.state_dict(checkpoint, representation) if rank == 0 else None
        engine = broadcast_engine(state, representation, gpu)
        with load._lock:
            load._engines[key] = engine
    dataloader = data.loader(
        audio_files, num_workers=max(num_workers // 2, 1), max_frames=max_frames,
        shard=(rank, world), device=gpu)
    mapping = dict(zip(audio_files, output_files))
    if core._native_pipeline(dataloader, representation):
        engine = load.model(checkpoint, representation, gpu)
        dataloader.run_native(engine, mapping, max(num_workers // 2, 1), legacy_mode)
    else:
        core.from_dataloader(dataloader, mapping, representation, checkpoint,
                             save_workers=max(num_workers // 2, 1), gpu=gpu,
                             legacy_mode=legacy_mode)
    if world > 1:
        dist.barrier()
