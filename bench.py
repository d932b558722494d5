#!/usr/bin/env python
"""bench.py — PPG frames/s of the `ppgs.from_audio` hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): mel + default Transformer, batch = 64 x 10 s
synthetic 16 kHz audio per GPU (weak scaling: utterances shard on the batch axis,
no data-path collective; one weight-blob broadcast at load).  A "step" is one
pass of the hot path (fused STFT+mel -> Transformer -> softmax) over one batch.

  value  device-resident: the batch is already in HBM when the timed region
         starts (CUDA events, max over ranks);
  e2e    the same metric through the C-ABI `ppgs_from_audio_host` with pinned
         HOST buffers: H2D + compute + D2H inside the timed region;
  roofline    the dominant kernel (per-kernel CUDA events, a second timed pass
         over the same steps with the library's launch profiling enabled);
  cpu_baseline / --impl reference    the reference's OWN code (oracle/_ref, the
         travelled byte-identical copy made by oracle/build_ref.py, imported under
         oracle/refshim.py) as shipped: gpu=None, bf16 autocast, all host cores
         (`kind: "reference"`; falls back to the oracle port when the copy is absent);
  torch_gpu_baseline   the same reference code on the same B200 (gpu=0: PyTorch eager,
         fp16 autocast = oracle mode O2; and its modules in fp32 with TF32 off = O3),
         timed outside the headline region — the library baseline the kernels must beat;
  spread    the K-step window is repeated and the median reported (min / max beside it).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH, SECONDS, SAMPLE_RATE, HOP = 64, 10, 16000, 160
SAMPLES = SECONDS * SAMPLE_RATE
FRAMES = SAMPLES // HOP
METRIC, UNIT = 'ppg_frames_per_sec', 'frames/s'
WORKLOAD = 'mel + default Transformer (5x256, 2 heads, FFN 2048), batch=64x10s synthetic 16 kHz'

# Algorithmic FLOPs of the reference's chunked algorithm:
# 1000 frames -> chunks of 500/500/250 computed frames.
H, C_IN, F_FFN, O_OUT, KSIZE, LAYERS = 256, 80, 2048, 40, 5, 5
CHUNKS = (500, 500, 250)
COMPUTED_FRAMES = sum(CHUNKS)
DENSE_PER_FRAME = 2 * (KSIZE * C_IN * H + LAYERS * (3 * H * H + H * H + 2 * H * F_FFN)
                       + KSIZE * H * O_OUT)
ATTN_PER_UTT = sum(LAYERS * 4 * H * s * s for s in CHUNKS)
FLOPS_PER_UTT = DENSE_PER_FRAME * COMPUTED_FRAMES + ATTN_PER_UTT
FFN_FLOPS_PER_LAUNCH_HALF = 2 * H * F_FFN     # per computed frame, one of linear1 / linear2


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        peaks = json.load(open(path))
        return {'hbm_gbs': peaks['hbm_gbs'], 'tflops_burst': peaks['bf16_tflops'],
                'tflops_sustained': peaks['bf16_tflops_sustained'], 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0,
            'source': 'fallback'}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this
    workload, from the committed `ncu --set full` capture (profiles/ncu_traffic.json);
    None when the kernel has no capture."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None, None, None
    entry = json.load(open(path)).get('kernels', {}).get(kernel)
    return (None, None, None) if entry is None else (
        entry['dram_bytes'], entry['source'], entry.get('tensor_pipe_active_pct'))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, sm_max, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                sm_max = float(parts[1])
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': sm_max,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_rate(steps, warmup, utterances=None, budget_s=20.0):
    """The reference's CPU path as shipped (oracle.AsShipped: stock torch layers,
    bf16 autocast, inference_mode) on all host cores, on a bounded sample of the
    workload: `utterances` x 10 s per step through mel.from_audios +
    from_features (ppgs/core.py:333-352; from_audio itself fails for B>1)."""
    import torch
    from oracle import ppg_oracle as O
    from oracle import ref_arm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.random_state_dict(0)
    if ref_arm.available():
        # the reference's own code: mel.from_audios + ppgs.from_features(gpu=None)
        ckpt = ref_arm.checkpoint_for(sd)
        run = lambda audio: ref_arm.from_audios(audio, ckpt, gpu=None)   # noqa: E731
        what = ("reference's own ppgs.preprocess.mel.from_audios + ppgs.from_features(gpu=None) "
                '(oracle/_ref, bf16 autocast as shipped)')
    else:
        run = O.AsShipped(sd).from_audios
        what = 'oracle port (as-shipped bf16-autocast torch modules; oracle/_ref absent)'
    probe = O.synthetic_audio(2, SAMPLES, 0)
    run(probe)
    t0 = time.perf_counter()
    run(probe)
    per_utt = (time.perf_counter() - t0) / 2
    if utterances is None:
        utterances = int(max(2, min(BATCH, budget_s / max(steps + warmup, 1) / per_utt)))
    audio = O.synthetic_audio(utterances, SAMPLES, 1)
    for _ in range(warmup):
        run(audio)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = run(audio)
        times.append(time.perf_counter() - t0)
    assert out.shape == (utterances, O_OUT, FRAMES)
    mean = sum(times) / len(times)
    return {'value': utterances * FRAMES / mean, 'unit': UNIT, 'cores': torch.get_num_threads(),
            'kind': ref_arm.kind(),
            'sample': f'{utterances}x10s utterances per step, {steps} steps (+{warmup} warm-up), '
                      f'{what}, output dtype {out.dtype}'.replace('torch.', '')}, mean


def torch_gpu_baseline(device, audio_dev, audio_host, steps=5, warmup=3):
    """The bar on the same box: the reference's own code on
    this GPU.  O2 = as shipped with gpu=0 (PyTorch eager, fp16 autocast: cuBLAS / cuDNN /
    SDPA kernels); O3 = its modules in fp32 with autocast and TF32 off (the numerics class of
    the 1e-4 target).  Same 64 x 10 s batch, inputs resident on the device, CUDA events,
    outside the headline timed region."""
    import torch
    from oracle import ppg_oracle as O
    from oracle import ref_arm
    if not ref_arm.available():
        return {'unavailable': 'oracle/_ref absent (run python -m oracle.build_ref in the dev container)'}
    sd = O.random_state_dict(0, peaky=True)
    ckpt = ref_arm.checkpoint_for(sd)
    ref = O.from_audio(sd, audio_host[:2].unsqueeze(1))
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    result = {'batch': f'{BATCH}x10s, device-resident audio', 'steps': steps, 'warmup': warmup,
              'torch': torch.__version__}
    try:
        fp32 = ref_arm.modules_fp32(ckpt, device)
        arms = {
            'fp16_autocast_as_shipped': lambda a: ref_arm.from_audios(a, ckpt, gpu=device.index),
            'fp32_tf32_off': fp32,
        }
        audio3 = audio_dev.unsqueeze(1)
        for name, fn in arms.items():
            for _ in range(warmup):
                out = fn(audio3)
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(device)
            start.record()
            for _ in range(steps):
                out = fn(audio3)
            stop.record()
            torch.cuda.synchronize(device)
            ms = start.elapsed_time(stop) / steps
            result[name] = {'ms_per_step': ms, 'frames_per_sec': BATCH * FRAMES / (ms / 1e3),
                            'max_abs_vs_oracle': (out[:2].float().cpu() - ref).abs().max().item(),
                            'output_dtype': str(out.dtype).replace('torch.', '')}
            del out
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return result


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(args.steps, 1)
    warmup = max(args.warmup, 1)
    base, mean = cpu_reference_rate(steps, warmup, budget_s=90.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': mean * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16 (cpu autocast)', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': 'CPU reference arm: bounded sample of the workload'},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_other_workload(args):
    print(json.dumps(measure_other_workload(args.workload, max(args.steps, 1), max(args.warmup, 3))),
          flush=True)


def measure_other_workload(workload_name, steps, warmup):
    """BASELINE configs[2] / configs[3] on one GPU, same JSON keys (not the headline line)."""
    import torch
    import ppgs_b200
    from oracle import ppg_oracle as O   # synthetic inputs only
    device = torch.device('cuda', 0)
    torch.cuda.set_device(device)
    peaks = load_peaks()

    def timed(fn, n):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        start.record()
        for i in range(n):
            fn(i)
        stop.record()
        torch.cuda.synchronize(device)
        return start.elapsed_time(stop)

    if workload_name == 'w2v2fb':
        from oracle import w2v2_oracle as W
        batch = 32
        front = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0))
        front.load_w2v2_state_dict(W.random_state_dict(0))
        head = ppgs_b200.Engine(0, input_channels=768, hidden_channels=512).load_state_dict(
            O.random_state_dict(1, input_channels=768, hidden_channels=512, peaky=True))
        rotate = 4
        host = [O.synthetic_audio(batch, SAMPLES, 200 + i).squeeze(1).pin_memory() for i in range(rotate)]
        dev = [h.to(device) for h in host]
        lengths = torch.full((batch,), FRAMES)
        out_host = torch.empty(batch, O_OUT, FRAMES, dtype=torch.float32).pin_memory()

        def device_step(i):
            return head.transformer(front.w2v2fb(dev[i % rotate]), lengths)

        def host_step(i):
            audio = host[i % rotate].to(device, non_blocking=True)
            out_host.copy_(head.transformer(front.w2v2fb(audio), lengths), non_blocking=True)

        frames_per_step, flops_per_frame = batch * FRAMES, 198.6e6
        workload = 'w2v2fb (wav2vec2-base front-end + hidden-512 PPG Transformer), batch=32x10s synthetic 16 kHz'
        engines = (front, head)
    else:
        streams, chunk, pushes = 256, 160, 3
        engine = ppgs_b200.Engine(0, is_causal=True).load_state_dict(O.random_state_dict(2, peaky=True))
        engine.precision = 'f16x2'
        features = O.mel_from_audios(O.synthetic_audio(8, chunk * pushes * 160, 0))
        features = features.repeat(streams // 8, 1, 1).contiguous()
        host = [features[..., i * chunk:(i + 1) * chunk].contiguous().pin_memory() for i in range(pushes)]
        dev = [h.to(device) for h in host]
        streamer = ppgs_b200.Streamer(engine, streams)
        out_host = torch.empty(streams, O_OUT, chunk + 4, dtype=torch.float32).pin_memory()

        def device_step(i):   # one step = one session: three 160-frame pushes with state
            streamer.reset()
            for j in range(pushes):
                streamer.push(dev[j], final=j + 1 == pushes)

        def host_step(i):
            streamer.reset()
            for j in range(pushes):
                out = streamer.push(host[j].to(device, non_blocking=True), final=j + 1 == pushes)
                out_host[..., :out.shape[-1]].copy_(out, non_blocking=True)

        frames_per_step = streams * chunk * pushes
        # dense 13.414 MFLOP per computed frame + causal attention over the growing context
        attention = sum(LAYERS * 4 * H * (t + 1) for t in range(chunk * pushes)) / (chunk * pushes)
        flops_per_frame = DENSE_PER_FRAME + attention
        workload = ('config/causal_transformer.py, stateful streaming: 256 sessions x three 160-frame '
                    'pushes (480-frame context), mel features in')
        engines = (engine,)

    for i in range(warmup):
        device_step(i)
        host_step(i)
    sampler = ClockSampler(0)
    sampler.start()
    before = sum(e.launches for e in engines)
    device_ms = timed(device_step, steps)
    launches = sum(e.launches for e in engines) - before
    e2e_ms = timed(host_step, steps)
    clocks = sampler.stop()
    for e in engines:
        e.set_profiling(True)
    timed(device_step, steps)
    stats = {}
    for e in engines:
        for k, (ms, n) in e.kernel_stats().items():
            stats[k] = round(stats.get(k, 0.0) + ms / steps, 4)
        e.set_profiling(False)
    value = frames_per_step / (device_ms / steps / 1e3)
    tflops = value * flops_per_frame / 1e12
    in_bytes = sum(h.numel() * h.element_size() for h in host) if workload_name != 'w2v2fb' \
        else host[0].numel() * 4
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': steps, 'warmup': warmup,
        'ms_per_step': device_ms / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16x2-split tcgen05 MMA, f32 accumulate', 'data': 'synthetic',
        'config': {'workload': workload, 'l2': 'rotating inputs; activation traffic per step exceeds L2'},
        'e2e': {'value': frames_per_step / (e2e_ms / steps / 1e3), 'unit': UNIT,
                'ms_per_step': e2e_ms / steps, 'h2d_bytes_per_step': in_bytes,
                'd2h_bytes_per_step': frames_per_step * O_OUT * 4,
                'api': 'Engine / Streamer public API, pinned host tensors in, pinned host tensor out'},
        'gpu_launches': launches, 'clocks': clocks,
        'roofline': {'kernel': 'whole step', 'bound': 'tensor', 'achieved': tflops,
                     'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s',
                     'frac': tflops / peaks['tflops_sustained'], 'traffic': None,
                     'algorithmic_flops_per_frame': flops_per_frame, 'kernels_ms_per_step': stats},
        'cpu_baseline': None,
    }
    return line


def run_files_workload(args, rank, world, local_rank):
    """BASELINE configs[4]: `from_files_to_files` over a synthetic corpus of 10 s 16 kHz int16
    WAVE files, batch-sharded over the ranks (ppgs_b200.parallel.from_files_to_files: every
    rank takes batches rank::world of the same deterministic batch list, no collective).  Weak
    scaling: `--files-per-gpu` files per rank (4500 x 8 ranks = 36 000 files = the 100 h corpus).
    Files live on tmpfs (/dev/shm) spread over 64 directories; outputs likewise.  Wall clock
    between barriers (a host pipeline: file reads, PCIe, kernels and .pt writes all count), max
    over ranks.  PPGS_B200_FILES_NULL_GPU=1 gives the host ceiling (kernels skipped)."""
    import shutil
    import wave
    import numpy as np
    import torch
    import torch.distributed as dist
    import ppgs_b200
    from ppgs_b200 import parallel
    from oracle import ppg_oracle as O   # synthetic weights only
    torch.cuda.set_device(local_rank)
    if world > 1:
        parallel.init('nccl')
    per_gpu = args.files_per_gpu
    total = per_gpu * world
    root = os.path.join('/dev/shm' if os.path.isdir('/dev/shm') else '/tmp', f'ppgs_b200_files_{args.run_id}')
    dirs = 64
    audio_files = [os.path.join(root, f'in{i % dirs:02d}', f'{i:06d}.wav') for i in range(total)]
    output_files = [os.path.join(root, f'out{i % dirs:02d}', f'{i:06d}-ppg.pt') for i in range(total)]
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
        for d in range(dirs):
            os.makedirs(os.path.join(root, f'in{d:02d}'))
            os.makedirs(os.path.join(root, f'out{d:02d}'))
        torch.save({'model': O.random_state_dict(0, peaky=True)}, os.path.join(root, 'ckpt.pt'))
    if world > 1:
        dist.barrier()
    rng = np.random.default_rng(rank)
    base = (rng.uniform(-0.5, 0.5, SAMPLES + per_gpu) * 32767).astype(np.int16)
    for j, i in enumerate(range(rank, total, world)):        # every rank writes its share of the corpus
        with wave.open(audio_files[i], 'wb') as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(SAMPLE_RATE)
            f.writeframes(base[j:j + SAMPLES].tobytes())
    if world > 1:
        dist.barrier()
    checkpoint = os.path.join(root, 'ckpt.pt')
    workers = args.file_workers
    try:
        warm = 256 * world
        parallel.from_files_to_files(audio_files[:warm], output_files[:warm], 'mel', checkpoint,
                                     num_workers=workers, max_frames=BATCH * FRAMES)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        passes = []
        for _ in range(max(args.file_repeats, 1)):   # whole-corpus passes; sub-second each, so repeated
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            parallel.from_files_to_files(audio_files, output_files, 'mel', checkpoint, num_workers=workers,
                                         max_frames=BATCH * FRAMES)
            torch.cuda.synchronize()
            elapsed = torch.tensor([time.perf_counter() - t0], device='cuda')
            if world > 1:
                dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
            passes.append(elapsed.item())
        seconds = statistics.median(passes)
        if rank == 0:
            sample = torch.load(output_files[-1])
            assert sample.shape == (O_OUT, FRAMES) and bool(torch.isfinite(sample).all())
            engine = ppgs_b200.load.model(checkpoint, 'mel', local_rank)
            null_gpu = os.environ.get('PPGS_B200_FILES_NULL_GPU', '0') not in ('', '0')
            line = {
                'metric': METRIC, 'value': total * FRAMES / seconds, 'unit': UNIT, 'n_gpus': world,
                'steps': 1, 'warmup': 1, 'ms_per_step': seconds * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f16x2-split tcgen05 MMA, f32 accumulate', 'data': 'synthetic',
                'config': {'workload': 'from_files_to_files, synthetic corpus of 10 s 16 kHz int16 WAVE files '
                                       f'({total} files = {total * SECONDS / 3600:.1f} h), max_frames=64000, '
                                       'batch-sharded over the ranks', 'files_per_gpu': per_gpu,
                           'files': total, 'storage': root.split('/')[1] + ' (tmpfs), 64 input + 64 output directories',
                           'host_cores': os.cpu_count(), 'reader_writer_threads_per_rank': workers,
                           'null_gpu_host_ceiling': null_gpu, 'precision': engine.precision},
                'files_per_sec': total / seconds, 'audio_hours_per_hour': total * SECONDS / seconds,
                'seconds': seconds, 'passes_seconds': [round(x, 4) for x in passes], 'statistic': 'median pass',
                'e2e': {'value': total * FRAMES / seconds, 'unit': UNIT,
                        'h2d_bytes_per_step': total * SAMPLES * 2, 'd2h_bytes_per_step': total * O_OUT * FRAMES * 4,
                        'api': 'ppgs_b200.parallel.from_files_to_files (files on tmpfs in, .pt files out)'},
                'gpu_launches': int(engine.launches),
            }
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(root, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=20)
    parser.add_argument('--warmup', type=int, default=5)
    parser.add_argument('--impl', default='ppgs_b200', choices=['ppgs_b200', 'reference'])
    parser.add_argument('--precision', default=os.environ.get('PPGS_B200_PRECISION', 'auto'))
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-torch-baseline', action='store_true')
    parser.add_argument('--no-other-configs', action='store_true')
    parser.add_argument('--windows', type=int, default=10,
                        help='repeats of the K-step timed window (median reported, min / max in "spread")')
    parser.add_argument('--workload', default='mel', choices=['mel', 'w2v2fb', 'causal-stream', 'files'],
                        help="mel = BASELINE configs[1] (the headline; default); w2v2fb = configs[2]; "
                             "causal-stream = configs[3] with state (extra lines, N=1 only); "
                             "files = configs[4], from_files_to_files over a synthetic corpus (any N)")
    parser.add_argument('--files-per-gpu', type=int, default=4500,
                        help='files workload: 10 s files per rank (4500 x 8 = the 100 h corpus)')
    parser.add_argument('--file-workers', type=int, default=16,
                        help='files workload: num_workers per rank (half readers, half writers)')
    parser.add_argument('--file-repeats', type=int, default=5, help='files workload: timed passes over the corpus')
    parser.add_argument('--run-id', default=os.environ.get('MASTER_PORT', 'single'))
    args = parser.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.workload == 'files':
        run_files_workload(args, rank, world, local_rank)
        return
    if args.workload != 'mel':
        if rank == 0:
            run_other_workload(args)
        return

    import torch
    import torch.distributed as dist
    import ppgs_b200
    from ppgs_b200 import parallel
    from oracle import ppg_oracle as O   # synthetic inputs + CPU baseline only

    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        parallel.init('nccl')

    # weights: rank 0 owns the state dict; one blob broadcast, no per-step collective
    state = O.random_state_dict(0, peaky=True) if rank == 0 else None
    engine = parallel.broadcast_engine(state, 'mel', local_rank)
    if args.precision == 'auto':
        for name in ('f16x2', 'fp32'):
            try:
                engine.precision = name
                break
            except RuntimeError:
                continue
    else:
        engine.precision = args.precision
    precision = engine.precision

    # inputs larger than L2: 4 rotating batches (4 x 41 MB of audio; the per-step
    # activation traffic in the workspace is > 1 GB on top)
    rotate = 4
    host = [O.synthetic_audio(BATCH, SAMPLES, 100 + rank * rotate + i).squeeze(1).pin_memory()
            for i in range(rotate)]
    dev = [h.to(device) for h in host]
    out_host = [torch.empty(BATCH, O_OUT, FRAMES, dtype=torch.float32).pin_memory()
                for _ in range(2)]

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, n):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for i in range(n):
            fn(i)
        engine.wait()      # pipelined host requests: the last D2H has landed
        stop.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(stop)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def timed_windows(fn, n, windows):
        """`windows` back-to-back timed regions of exactly `n` steps each (every one bracketed
        by barrier + synchronize, max over ranks); the headline is the median window."""
        return [timed(fn, n) for _ in range(windows)]

    # device-resident step through the C ABI (ppgs_from_audio) into a fixed output buffer:
    # no allocation and no host work besides the call itself inside the timed region
    out_dev = torch.empty(BATCH, O_OUT, FRAMES, dtype=torch.float32, device=device)
    device_step = lambda i: engine.from_audio(dev[i % rotate], out=out_dev)         # noqa: E731
    # reference-facing call with HOST buffers; requests are pipelined two deep (the
    # H2D of step i+1 and the D2H of step i-1 overlap the kernels of step i), every
    # step still copies its own input in and its own result out
    host_step = lambda i: engine.from_audio_host(                                   # noqa: E731
        host[i % rotate], out=out_host[i % 2], wait=False)

    for i in range(warmup):
        device_step(i)
        host_step(i)
    engine.wait()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_before = engine.launches
    device_windows = timed_windows(device_step, steps, args.windows)
    launches = (engine.launches - launches_before) // args.windows
    e2e_windows = timed_windows(host_step, steps, max(1, args.windows // 2))
    clocks = sampler.stop() if rank == 0 else None
    device_ms = statistics.median(device_windows)
    e2e_ms = statistics.median(e2e_windows)

    # roofline pass: same steps with per-kernel CUDA events (library profiling)
    engine.set_profiling(True)
    timed(device_step, steps)
    stats = engine.kernel_stats()
    engine.set_profiling(False)

    # parity of what was just timed (2 rows of the last batch vs the oracle, rank 0)
    parity = None
    if rank == 0:
        # rows 0-1 of the FULL batch (the kernels that were timed: small batches take other paths)
        last = (steps - 1) % rotate
        got = engine.from_audio(dev[last], out=out_dev)[:2].cpu()
        ref = O.from_audio(O.random_state_dict(0, peaky=True), host[last][:2].unsqueeze(1))
        parity = (got - ref).abs().max().item()

    # context only (never the headline): the same kernels with ONE fp16 MMA pass, i.e. the
    # reference's own CUDA-autocast numerics class, which misses the 1e-4 parity bar
    single_pass = None
    if precision == 'f16x2' and rank == 0 and world == 1:
        engine.precision = 'f16'
        for i in range(3):
            device_step(i)
        ms = timed(device_step, steps) / steps
        got = engine.from_audio(dev[0], out=out_dev)[:2].cpu()
        ref = O.from_audio(O.random_state_dict(0, peaky=True), host[0][:2].unsqueeze(1))
        single_pass = {'ms_per_step': ms, 'frames_per_sec': BATCH * FRAMES / (ms / 1e3),
                       'max_abs_vs_oracle': (got - ref).abs().max().item(),
                       'note': 'PPGS_PRECISION_F16, not parity-valid; for context'}
        engine.precision = precision

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    frames_total = BATCH * FRAMES * world
    value = frames_total / (device_ms / steps / 1e3)
    e2e_value = frames_total / (e2e_ms / steps / 1e3)

    # dominant kernel = largest share of device time among the profiled kernels
    total_kernel_ms = sum(ms for ms, _ in stats.values()) or 1.0
    name, (kernel_ms, kernel_launches) = max(stats.items(), key=lambda kv: kv[1][0])
    per_launch_ms = kernel_ms / max(kernel_launches, 1)
    computed = BATCH * COMPUTED_FRAMES
    flops_by_kernel = {
        'ffn1': FFN_FLOPS_PER_LAUNCH_HALF * computed, 'ffn2': FFN_FLOPS_PER_LAUNCH_HALF * computed,
        'ffn': 2 * FFN_FLOPS_PER_LAUNCH_HALF * computed, 'qkv': 2 * 3 * H * H * computed,
        'out_proj': 2 * H * H * computed, 'conv_in': 2 * KSIZE * C_IN * H * computed,
        'conv_out': 2 * KSIZE * H * O_OUT * computed, 'attention': BATCH * ATTN_PER_UTT / LAYERS}
    traffic, traffic_source, tensor_pipe = ncu_traffic(name)
    roofline = {'kernel': name, 'share_of_step': kernel_ms / total_kernel_ms,
                'ms_per_launch': per_launch_ms, 'traffic': traffic, 'traffic_unit': 'bytes per launch',
                'traffic_source': traffic_source,
                # committed ncu capture of this kernel (not measured in this run)
                'ncu_tensor_pipe_active_pct': tensor_pipe}
    key = next((k for k in sorted(flops_by_kernel, key=len, reverse=True) if k in name), None)
    if key is not None:
        achieved = flops_by_kernel[key] / (per_launch_ms * 1e-3) / 1e12
        peak = peaks['tflops_sustained']
        roofline.update({'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak,
                         'peak_source': f"{peaks['source']} bf16 cuBLAS sustained (kernel timed "
                                        'inside a long step)',
                         'algorithmic_flops_per_launch': flops_by_kernel[key],
                         'mma_passes_per_algorithmic_mma': 3 if precision == 'f16x2' else 1,
                         'executed_frac': achieved / peak * (3 if precision == 'f16x2' else 1)})
    else:   # mel front-end
        bytes_per_launch = BATCH * FRAMES * 800
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        roofline.update({'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                         'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                         'peak_source': peaks['source']})
    whole = FLOPS_PER_UTT * BATCH / (device_ms / steps * 1e-3) / 1e12
    roofline['whole_step'] = {'achieved_tflops': whole, 'frac': whole / peaks['tflops_sustained'],
                              'algorithmic_flops_per_step': FLOPS_PER_UTT * BATCH}
    mel_stat = stats.get('mel_stft_fbank_rows') or stats.get('mel_stft_fbank')
    if mel_stat:
        mel_ms = mel_stat[0] / max(mel_stat[1], 1)
        gbs = BATCH * FRAMES * 800 / (mel_ms * 1e-3) / 1e9
        roofline['mel_frontend'] = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm_gbs'],
                                    'unit': 'GB/s', 'frac': gbs / peaks['hbm_gbs'],
                                    'ms_per_launch': mel_ms}
    roofline['kernels_ms_per_step'] = {k: round(ms / steps, 4) for k, (ms, _) in sorted(stats.items())}

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # reported baseline: rank 0 at N = 1 only
        cpu, _ = cpu_reference_rate(steps=3, warmup=1, budget_s=20.0)

    # the same-box library baseline (outside every timed region of the product)
    torch_gpu = None
    if not args.no_torch_baseline and world == 1:
        try:
            torch_gpu = torch_gpu_baseline(device, dev[0], host[0])
            for arm in ('fp16_autocast_as_shipped', 'fp32_tf32_off'):
                if arm in torch_gpu:
                    torch_gpu[arm]['ppgs_b200_speedup'] = torch_gpu[arm]['ms_per_step'] / (device_ms / steps)
            if single_pass and 'fp16_autocast_as_shipped' in torch_gpu:
                torch_gpu['fp16_autocast_as_shipped']['ppgs_b200_f16_single_pass_speedup'] = (
                    torch_gpu['fp16_autocast_as_shipped']['ms_per_step'] / single_pass['ms_per_step'])
        except Exception as error:   # a baseline must never take the headline down
            torch_gpu = {'unavailable': f'{type(error).__name__}: {error}'[:300]}

    # BASELINE configs[2] / configs[3] in short form (their own full lines: --workload ...)
    others = None
    if not args.no_other_configs and world == 1:
        others = {}
        torch.cuda.empty_cache()
        for name in ('w2v2fb', 'causal-stream'):
            try:
                full = measure_other_workload(name, steps=5, warmup=3)
                others[name] = {'value': full['value'], 'unit': UNIT, 'ms_per_step': full['ms_per_step'],
                                'e2e_value': full['e2e']['value'], 'workload': full['config']['workload'],
                                'roofline_frac_whole_step': full['roofline']['frac'],
                                'gpu_launches': full['gpu_launches']}
            except Exception as error:
                others[name] = {'unavailable': f'{type(error).__name__}: {error}'[:300]}

    dtype = {'fp32': 'f32 (CUDA-core FFMA)', 'f16x2': 'f16x2-split tcgen05 MMA, f32 accumulate',
             'f16': 'f16 tcgen05 MMA, f32 accumulate'}.get(precision, precision)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
        'warmup': warmup, 'ms_per_step': device_ms / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': dtype, 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'per_gpu_batch': BATCH, 'samples': SAMPLES,
                   'frames_per_utterance': FRAMES, 'chunking': '500/400/50 (reference semantics)',
                   'precision': precision, 'parallelism': f'dp{world} (utterance shards, no collective)',
                   'l2': f'inputs larger than L2: {rotate} rotating batches + >1 GB activation traffic per step',
                   'max_abs_vs_oracle': parity},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms / steps,
                'h2d_bytes_per_step': BATCH * SAMPLES * 4 * world,
                'd2h_bytes_per_step': BATCH * O_OUT * FRAMES * 4 * world,
                'api': 'ppgs_from_audio_host_submit + ppgs_engine_wait (C ABI), pinned host '
                       'buffers, 2 requests in flight'},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': roofline,
        'cpu_baseline': cpu,
        'spread': {'windows': len(device_windows), 'steps_per_window': steps, 'statistic': 'median',
                   'ms_per_step_min': min(device_windows) / steps,
                   'ms_per_step_median': device_ms / steps,
                   'ms_per_step_max': max(device_windows) / steps,
                   'e2e_windows': len(e2e_windows),
                   'e2e_ms_per_step_min': min(e2e_windows) / steps,
                   'e2e_ms_per_step_max': max(e2e_windows) / steps},
        'torch_gpu_baseline': torch_gpu,
        'other_configs': others,
        'single_pass_f16_context': single_pass,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def stdout_is_the_json_line_only():
    """The driver reads ONE JSON line from stdout: send everything else written to file descriptor 1
    (NCCL's version banner, library chatter from C code) to stderr and keep the real stdout for
    print()."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, 'w', buffering=1)


if __name__ == '__main__':
    stdout_is_the_json_line_only()
    main()
