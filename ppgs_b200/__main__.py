"""`python -m ppgs_b200` — the flags of `python -m ppgs` (ppgs/__main__.py:12-59)."""
import argparse
from pathlib import Path

import ppgs_b200


def parse_args(argv=None):
    # --config first: it decides the defaults of the other flags (the reference's yapecs reads
    # it at import time, before `ppgs.REPRESENTATION` is looked at)
    early = argparse.ArgumentParser(add_help=False)
    early.add_argument('--config', type=Path, nargs='*')
    for file in early.parse_known_args(argv)[0].config or []:
        ppgs_b200.configure(file)
    parser = argparse.ArgumentParser(description='Phonetic posteriorgram inference')
    parser.add_argument('--audio_files', nargs='+', type=Path, required=True,
                        help='Paths to audio files')
    parser.add_argument('--output_files', nargs='+', type=Path, required=True,
                        help='The one-to-one corresponding output files')
    parser.add_argument('--representation', type=str, default=ppgs_b200.REPRESENTATION,
                        help='Representation to use for inference')
    parser.add_argument('--checkpoint', type=Path, help='The checkpoint file')
    parser.add_argument('--num-workers', type=int, default=0,
                        help='Number of CPU threads for reading / saving')
    parser.add_argument('--gpu', type=int, nargs='+',
                        help='The index of the GPU to use for inference (default: current); '
                             'several indices shard the file list over those GPUs')
    parser.add_argument('--max-frames', type=float, default=ppgs_b200.MAX_INFERENCE_FRAMES,
                        help='Maximum number of frames in a batch')
    parser.add_argument('--legacy-mode', action='store_true',
                        help='Use legacy (unchunked) inference')
    parser.add_argument('--config', type=Path, nargs='*',
                        help='yapecs-style configuration files (UPPER_CASE overrides)')
    args = parser.parse_args(argv)
    if args.gpu is not None and len(args.gpu) == 1:
        args.gpu = args.gpu[0]          # one index: the reference's `--gpu N`
    return args


def main(argv=None):
    args = vars(parse_args(argv))
    args.pop('config', None)   # applied by parse_args, before the defaults were read
    if isinstance(args['gpu'], list) and args['num_workers'] == 0:
        args['num_workers'] = 2 * len(args['gpu'])     # the sharded path is the batched one
    ppgs_b200.from_files_to_files(**args)


if __name__ == '__main__':
    main()
